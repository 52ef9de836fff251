"""MAPPO (separated nets) with the reference's class API (``MAPPO_file/MAPPO.py:106-507``) on the fused B200 kernels.

``MAPPO(dim_info: dict, is_continue, actor_lr, critic_lr, horizon, device, trick=None)``; ``select_action(obs: dict)
-> (actions, log_probs)``, ``evaluate_action``, ``add`` (dicts), ``all``, ``learn(minibatch_size, gamma, lmbda,
clip_param, K_epochs, entropy_coefficient, huber_delta=None)``, ``save``/``load``.

Reference behaviour kept: per-agent actor on its own obs and a centralised critic on the joint obs;
ONE Adam(eps 1e-5, lr = actor_lr) over actor+critic per agent and NO gradient clipping; GAE per agent column with its
own critic; joint advantage normalisation over ``[T, N]`` with the unbiased std; the broadcast quirk
``surr = ratio[mb,1] * adv[index][mb,N]`` (each agent's ratio multiplies ALL agents' advantages) and
``v_s.repeat(1,N)`` against ``v_target[mb,N]``; value clip + huber with ``max(original, clipped)`` — the clipped term
never exceeds the original elementwise, so loss and gradient are the original term's; LayerNorm/feature_norm without
affine; orthogonal init (gain sqrt2, 0.01 for the action head).  Continuous actions (the reference default).
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer_for_PPO
from .PPO import Agent as _PPOAgent


def net_init(m, gain=None, use_relu=True):
    """``MAPPO.py:106-125``: orthogonal weight (gain = calculate_gain('relu') unless given), zero bias."""
    gain = gain if gain is not None else nn.init.calculate_gain(['tanh', 'relu'][use_relu])
    nn.init.orthogonal_(m.weight, gain=gain)
    nn.init.constant_(m.bias, 0)


class Agent(_PPOAgent):
    def __init__(self, obs_dim, action_dim, dim_info, actor_lr, critic_lr, is_continue, device, trick):
        joint = sum(val[0] for val in dim_info.values())

        def hook(module, names, kind):
            if trick['orthogonal_init']:
                net_init(getattr(module, names[0]))
                net_init(getattr(module, names[1]))
                net_init(getattr(module, names[2]), gain=0.01) if kind == "actor" else net_init(getattr(module, names[2]))
        super().__init__(obs_dim, action_dim, actor_lr, critic_lr, is_continue, device, critic_in=joint, init_hook=hook)


class MAPPO:
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, mode=None):
        self.device = _lib.require_device(device)
        if bool(trick['LayerNorm']) != bool(trick['feature_norm']):
            raise NotImplementedError("LayerNorm and feature_norm must be switched together")
        self.agents, self.buffers = {}, {}
        for agent_id, (obs_dim, action_dim) in dim_info.items():
            self.agents[agent_id] = Agent(obs_dim, action_dim, dim_info, actor_lr, critic_lr, is_continue, self.device, trick)
            self.buffers[agent_id] = Buffer_for_PPO(horizon, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.dim_info = dim_info
        self.is_continue = is_continue
        print('actor_type:continue') if self.is_continue else print('actor_type:discrete')
        self.horizon = int(horizon)
        self.trick = trick
        self.num_agents = len(self.agents)
        self.layer_norm = bool(trick['LayerNorm'])
        self.mode = _common.resolve_mode(mode)
        self._seed = _common.default_seed()
        self._n_act = 0
        sm = _lib.sm_count()
        n_p = max(a._net.n_p for a in self.agents.values())
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self._gpart, self._sumsq, self._segcnt, self._stats = z(sm, n_p), z(sm, 2), z(sm, _lib.NSEG), z(sm, 8)
        self.last_metrics = None

    # ---- acting ------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        if not self.is_continue:
            return self._select_action_discrete(obs, noise)
        actions, action_log_pis = {}, {}
        self._n_act += 1
        for i, (agent_id, o) in enumerate(obs.items()):
            od, ad = self.dim_info[agent_id]
            x, single = _common.as_obs_batch(o, od)
            n = x.shape[0]
            nz = None
            if noise is not None:
                nz = torch.as_tensor(noise[agent_id], dtype=torch.float32).to(self.device).reshape(n, ad).contiguous()
            elif self.mode == "parity":
                nz = _common.reference_randn((n, ad), self.device)                 # dist.sample() per agent, in agent order
            out = _common.infer(self.agents[agent_id]._net, x, _lib.INFER_PPO_GAUSS, self.device, 2 * ad, noise=nz, seed=self._seed,
                                counter=self._n_act * 16 + i, l0=0, nl=3, layer_norm=self.layer_norm).cpu().numpy()
            a, lp = out[:, :ad], out[:, ad:]
            actions[agent_id] = a[0] if single else a
            action_log_pis[agent_id] = lp[0] if single else lp
        return actions, action_log_pis

    def evaluate_action(self, obs):
        actions = {}
        for agent_id, o in obs.items():
            od, ad = self.dim_info[agent_id]
            x, single = _common.as_obs_batch(o, od)
            if self.is_continue:
                a = _common.infer(self.agents[agent_id]._net, x, _lib.INFER_TANH, self.device, ad, l0=0, nl=3, layer_norm=self.layer_norm).cpu().numpy()
            else:                                     # np.argmax(a_prob)  (MAPPO.py:335)
                a = _common.infer(self.agents[agent_id]._net, x, _lib.INFER_ARGMAX, self.device, 1, l0=0, nl=3,
                                  layer_norm=self.layer_norm).reshape(-1).to(torch.int64).cpu().numpy()
            actions[agent_id] = a[0] if single else a
        return actions

    def _select_action_discrete(self, obs, noise):
        """``Categorical(probs=actor(obs)).sample()`` + ``log_prob`` per agent (MAPPO.py:316-318): torch.multinomial draws
        q ~ Exp(1) per class and takes argmax(p / q)"""
        actions, action_log_pis = {}, {}
        self._n_act += 1
        for i, (agent_id, o) in enumerate(obs.items()):
            od, ad = self.dim_info[agent_id]
            x, single = _common.as_obs_batch(o, od)
            n = x.shape[0]
            nz = None
            if noise is not None:
                nz = torch.as_tensor(noise[agent_id], dtype=torch.float32).to(self.device).reshape(n, ad).contiguous()
            elif self.mode == "parity":
                nz = torch.empty((n, ad), dtype=torch.float32, device=self.device).exponential_(1)
            out = _common.infer(self.agents[agent_id]._net, x, _lib.INFER_PPO_CAT, self.device, 2, noise=nz, seed=self._seed,
                                counter=self._n_act * 16 + i, l0=0, nl=3, layer_norm=self.layer_norm).cpu().numpy()
            a, lp = out[:, 0].astype(np.int64), out[:, 1]
            actions[agent_id] = a[0] if single else a
            action_log_pis[agent_id] = lp[0] if single else lp
        return actions, action_log_pis

    # ---- buffer ------------------------------------------------------------------------------------
    def add(self, obs, action, reward, next_obs, done, action_log_pi, adv_dones):
        for agent_id, buffer in self.buffers.items():
            buffer.add(obs[agent_id], action[agent_id], reward[agent_id], next_obs[agent_id], done[agent_id], action_log_pi[agent_id], adv_dones[agent_id])

    def all(self):
        keys = ("obs", "action", "reward", "next_obs", "done", "action_log_pi", "adv_dones")
        out = tuple({} for _ in keys)
        for agent_id, buffer in self.buffers.items():
            for d, v in zip(out, buffer.all()):
                d[agent_id] = v
        return out

    # ---- learning ----------------------------------------------------------------------------------
    def compute_advantages(self, gamma, lmbda):
        ids = list(self.agents.keys())
        b0 = self.buffers[ids[0]]
        M = b0.capacity
        E = b0.n_envs if (M % max(b0.n_envs, 1) == 0) else 1
        T = M // E
        joint = torch.cat([self.buffers[k].obs for k in ids], dim=1).contiguous()
        joint_n = torch.cat([self.buffers[k].next_obs for k in ids], dim=1).contiguous()
        advs, vts = [], []
        for k in ids:
            net, b = self.agents[k]._net, self.buffers[k]
            vs = _common.infer(net, joint, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, layer_norm=self.layer_norm)
            vs_ = _common.infer(net, joint_n, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, layer_norm=self.layer_norm)
            adv = torch.empty(M, dtype=torch.float32, device=self.device)
            vt = torch.empty(M, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().frl_gae(_lib.ptr(b.rewards), _lib.ptr(b.dones), _lib.ptr(b.adv_dones), _lib.ptr(vs), _lib.ptr(vs_), T, E,
                                          float(gamma), float(lmbda), _lib.ptr(adv), _lib.ptr(vt), _lib.stream_ptr(self.device)), "frl_gae")
            advs.append(adv)
            vts.append(vt)
        adv = torch.stack(advs, dim=1).contiguous()           # [M, N]
        v_target = torch.stack(vts, dim=1).contiguous()
        if self.trick['adv_norm']:
            dp = getattr(self, "_dp", None)
            if dp is not None and dp[2] > 1:
                # data parallel: the reference normalises over the WHOLE rollout (MAPPO.py:386-388), i.e. over the union of the ranks'
                # shards — all-reduce (sum, sum of squares, count) in float64, then normalise with the global mean / unbiased std
                dist, group, _ = dp
                a64 = adv.double()
                st = torch.stack([a64.sum(), (a64 * a64).sum(), torch.tensor(float(adv.numel()), dtype=torch.float64, device=self.device)])
                dist.all_reduce(st, group=group)
                n, mean = st[2], st[0] / st[2]
                std = ((st[1] - n * mean * mean) / (n - 1)).clamp_min(0.0).sqrt()
                adv = ((a64 - mean) / (std + 1e-8)).float().contiguous()
            else:
                _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(adv), adv.numel(), 1e-8, _lib.ptr(adv), _lib.stream_ptr(self.device)), "frl_adv_norm")
        return adv, v_target, joint

    # optimiser of MAPPO.update_ac (MAPPO.py:230-247): ONE Adam(eps 1e-5) over actor + critic, lr = actor_lr, no clip_grad_norm_
    max_norm = 0.0
    adam_eps = 1e-5

    def _permutations(self, ag, K_epochs, given):
        H = self.horizon
        if given is not None:
            return given
        if self.mode == "parity":
            return [np.random.permutation(H) for _ in range(K_epochs)]                  # MAPPO.py:395
        return None                              # fast mode: _agent_update builds the plan on the device

    def _agent_update(self, agent_id, adv, v_target, joint, minibatch_size, K_epochs, clip_param, entropy_coefficient, huber_delta, perms):
        """All K_epochs x minibatches of ONE agent in one persistent launch (adv / v_target: [M, n_adv] device tensors)."""
        ag, b = self.agents[agent_id], self.buffers[agent_id]
        H = self.horizon
        nmb = (H + minibatch_size - 1) // minibatch_size
        if perms is None:                       # fast mode: permutations and minibatch slicing on the device
            idx_d, rows_d, n_updates = _common.device_minibatch_plan(H, minibatch_size, K_epochs, self.device, self._seed + ag.step)
        else:
            idx = np.zeros((K_epochs * nmb, minibatch_size), np.int64)
            rows = np.zeros(K_epochs * nmb, np.int32)
            for e, perm in enumerate(perms):
                for j in range(nmb):
                    sl = np.asarray(perm[j * minibatch_size:(j + 1) * minibatch_size])
                    idx[e * nmb + j, :sl.size] = sl
                    rows[e * nmb + j] = sl.size
            idx_d, rows_d = torch.from_numpy(idx).to(self.device), torch.from_numpy(rows).to(self.device)
            n_updates = idx.shape[0]
        out = torch.zeros((n_updates, 8), dtype=torch.float32, device=self.device)
        a = _lib.PpoArgs()
        a.net, a.continuous = ag._net.c_struct(), int(self.is_continue)
        a.obs, a.action, a.logp_old = b.obs.data_ptr(), b.actions.data_ptr(), b.action_log_probs.data_ptr()
        a.adv, a.v_target = adv.data_ptr(), v_target.data_ptr()
        a.M, a.obs_dim, a.act_cols, a.logp_cols, a.n_adv = b.capacity, b.obs_dim, b.act_dim, b.logp_dim, adv.shape[1]
        a.indices, a.mb_rows, a.mb, a.n_updates = idx_d.data_ptr(), rows_d.data_ptr(), minibatch_size, n_updates
        a.clip_param, a.entropy_coef = clip_param, entropy_coefficient
        a.max_norm_actor = a.max_norm_critic = self.max_norm
        a.optimizer = _lib.OPT_ADAM
        a.lr, a.beta1, a.beta2, a.eps = ag.lr, 0.9, 0.999, self.adam_eps
        a.lr_critic = float(getattr(ag, "lr_critic", 0.0))
        a.step0 = ag.step
        a.layer_norm = int(self.layer_norm)
        a.critic_obs, a.critic_obs_dim = joint.data_ptr(), joint.shape[1]
        a.value_loss = 1 if self.trick['huber_loss'] else 0
        a.huber_delta = float(huber_delta) if huber_delta is not None else 0.0
        a.gpart, a.sumsq, a.segcnt = self._gpart.data_ptr(), self._sumsq.data_ptr(), self._segcnt.data_ptr()
        a.umma_ws = _common.umma_ws_ptr(self.device, int(a.mb))
        a.stats, a.out = self._stats.data_ptr(), out.data_ptr()
        self._launch_update(a, ag._net, n_updates)
        ag.step += n_updates
        self._keep = (idx_d, rows_d, joint, adv, v_target)
        return out

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None, *, permutations=None):
        adv, v_target, joint = self.compute_advantages(gamma, lmbda)
        self.last_adv, self.last_v_target = adv, v_target
        outs = []
        for agent_id, ag in self.agents.items():
            perms = self._permutations(ag, K_epochs, None if permutations is None else permutations[agent_id])
            outs.append(self._agent_update(agent_id, adv, v_target, joint, minibatch_size, K_epochs, clip_param, entropy_coefficient,
                                           huber_delta, perms))
        self.last_metrics = torch.cat(outs)
        for buffer in self.buffers.values():
            buffer.clear()

    enable_data_parallel = __import__("freerl_b200.PPO", fromlist=["PPO"]).PPO.enable_data_parallel      # one exchange block for all agents

    _launch_update = __import__("freerl_b200.PPO", fromlist=["PPO"]).PPO._launch_update

    def lr_decay(self, episode_num, max_episodes):
        """MAPPO.py:484-491 walks ``agent.actor_optimizer`` / ``agent.critic_optimizer``, which its merged-optimiser Agent does not
        have: upstream this raises AttributeError as soon as the ``lr_decay`` trick is on.  Kept as an error (not papered over);
        IPPO / HAPPO, whose agents do have the two optimisers, override it."""
        raise AttributeError("'Agent' object has no attribute 'actor_optimizer' (reference defect, MAPPO.py:488: lr_decay is not "
                             "usable with MAPPO's merged optimiser)")

    def _lr_decay_two_optimisers(self, episode_num, max_episodes):
        """IPPO.py:324-331 / HAPPO.py:460-467: both learning rates decay linearly with the episode count."""
        for ag in self.agents.values():
            ag.lr = self.actor_lr * (1 - episode_num / max_episodes)
            ag.lr_critic = self.critic_lr * (1 - episode_num / max_episodes)

    def save(self, model_dir):
        torch.save({name: {k: v.detach().clone().cpu() for k, v in agent.actor.state_dict().items()} for name, agent in self.agents.items()},
                   os.path.join(model_dir, 'MAPPO.pth'))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = MAPPO(dim_info, is_continue=is_continue, actor_lr=0, critic_lr=0, horizon=0, device=device, trick=trick)
        data = torch.load(os.path.join(model_dir, 'MAPPO.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
