"""IPPO with the reference's class API (``MAPPO_file/IPPO.py:100-340``) on the fused PPO kernels.

``IPPO(dim_info: dict, is_continue, actor_lr, critic_lr, horizon, device, trick)``: every agent is an independent PPO on its
OWN observation — decentralised critic, per-agent GAE (``frl_gae``) and per-agent ``adv_norm`` (``frl_adv_norm``), the MAPPO
network body (``feature_norm`` / ``LayerNorm``, orthogonal init), separate Adams for actor / critic (``eps 1e-5`` with
``adam_eps``; one per-network-lr sweep in the kernel, ``frl_ppo_args_t.lr_critic``) with ``clip_grad_norm_(0.5)`` each,
huber value loss.  ``ValueClip`` needs no kernel work: ``|clamp(v_target, V±c) - V| <= |v_target - V|`` element-wise and both
losses are monotone in ``|e|``, so ``max(loss_original, loss_clip)`` (``IPPO.py:299-307``) is always ``loss_original``.
Continuous (Gaussian, ``log_std`` parameter) and discrete (``Categorical(probs=softmax)``) actors are both fused.
"""
import os

import numpy as np
import torch

from . import _common, _lib
from .Buffer import Buffer_for_PPO
from .MAPPO import MAPPO as _MAPPO, net_init
from .PPO import Agent as _PPOAgent


class Agent(_PPOAgent):
    def __init__(self, obs_dim, action_dim, actor_lr, critic_lr, is_continue, device, trick):
        def hook(module, names, kind):
            if trick['orthogonal_init']:
                net_init(getattr(module, names[0]))
                net_init(getattr(module, names[1]))
                net_init(getattr(module, names[2]), gain=0.01) if kind == "actor" else net_init(getattr(module, names[2]))
        super().__init__(obs_dim, action_dim, actor_lr, critic_lr, is_continue, device, critic_in=obs_dim, init_hook=hook)
        self.lr_critic = critic_lr


class IPPO(_MAPPO):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, mode=None):
        self.device = _lib.require_device(device)
        if bool(trick['LayerNorm']) != bool(trick['feature_norm']):
            raise NotImplementedError("LayerNorm and feature_norm must be switched together")
        self.agents, self.buffers = {}, {}
        for agent_id, (obs_dim, action_dim) in dim_info.items():
            self.agents[agent_id] = Agent(obs_dim, action_dim, actor_lr, critic_lr, is_continue, self.device, trick)
            self.buffers[agent_id] = Buffer_for_PPO(horizon, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
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

    # select_action / evaluate_action: inherited from MAPPO (Gaussian and Categorical heads, same network body)

    def lr_decay(self, episode_num, max_episodes):
        self._lr_decay_two_optimisers(episode_num, max_episodes)

    # ---- learning ----------------------------------------------------------------------------------
    def compute_advantages(self, gamma, lmbda):
        """dict agent -> (adv [M, 1], v_target [M, 1])"""
        return {k: self.compute_advantages_one(k, gamma, lmbda) for k in self.agents}

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, huber_delta=None, *, permutations=None):
        H = self.horizon
        nmb = (H + minibatch_size - 1) // minibatch_size
        outs, keep = [], []
        self.last_adv, self.last_v_target = {}, {}
        for agent_id, ag in self.agents.items():
            b = self.buffers[agent_id]
            # the reference computes agent i's advantages inside the agent loop, i.e. after agents < i were updated; the
            # agents share nothing, so the order is immaterial
            adv, v_target = self.compute_advantages_one(agent_id, gamma, lmbda)
            self.last_adv[agent_id], self.last_v_target[agent_id] = adv, v_target
            if permutations is None and self.mode != "parity":          # fast mode: plan built on the device
                idx_d, rows_d, n_updates = _common.device_minibatch_plan(H, minibatch_size, K_epochs, self.device, self._seed + ag.step)
            else:
                perms = permutations[agent_id] if permutations is not None else [np.random.permutation(H) for _ in range(K_epochs)]  # IPPO.py:274
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
            a.M, a.obs_dim, a.act_cols, a.logp_cols, a.n_adv = b.capacity, b.obs_dim, b.act_dim, b.logp_dim, 1
            a.indices, a.mb_rows, a.mb, a.n_updates = idx_d.data_ptr(), rows_d.data_ptr(), minibatch_size, n_updates
            a.clip_param, a.entropy_coef = clip_param, entropy_coefficient
            a.max_norm_actor = a.max_norm_critic = 0.5                                # IPPO.py:175,182
            a.optimizer = _lib.OPT_ADAM
            a.lr, a.lr_critic = ag.lr, ag.lr_critic
            a.beta1, a.beta2, a.eps = 0.9, 0.999, (1e-5 if self.trick['adam_eps'] else 1e-8)
            a.step0 = ag.step
            a.layer_norm = int(self.layer_norm)
            a.value_loss = 1 if self.trick['huber_loss'] else 0
            a.huber_delta = float(huber_delta) if huber_delta is not None else 0.0
            a.gpart, a.sumsq, a.segcnt = self._gpart.data_ptr(), self._sumsq.data_ptr(), self._segcnt.data_ptr()
            a.umma_ws = _common.umma_ws_ptr(self.device, int(a.mb))
            a.stats, a.out = self._stats.data_ptr(), out.data_ptr()
            self._launch_update(a, ag._net, n_updates)
            ag.step += n_updates
            outs.append(out)
            keep.append((idx_d, rows_d, adv, v_target))
        self._keep = keep
        self.last_metrics = torch.cat(outs)
        for buffer in self.buffers.values():
            buffer.clear()

    def compute_advantages_one(self, agent_id, gamma, lmbda):
        """critic(own obs), critic(own next_obs) -> GAE scan -> optional adv_norm over this agent's [M, 1] (IPPO.py:255-270)"""
        ag, b = self.agents[agent_id], self.buffers[agent_id]
        M = b.capacity
        E = b.n_envs if (M % max(b.n_envs, 1) == 0) else 1
        vs = _common.infer(ag._net, b.obs, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, layer_norm=self.layer_norm)
        vs_ = _common.infer(ag._net, b.next_obs, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, layer_norm=self.layer_norm)
        adv = torch.empty((M, 1), dtype=torch.float32, device=self.device)
        vt = torch.empty((M, 1), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().frl_gae(_lib.ptr(b.rewards), _lib.ptr(b.dones), _lib.ptr(b.adv_dones), _lib.ptr(vs), _lib.ptr(vs_), M // E, E,
                                      float(gamma), float(lmbda), _lib.ptr(adv), _lib.ptr(vt), _lib.stream_ptr(self.device)), "frl_gae")
        if self.trick['adv_norm']:
            _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(adv), adv.numel(), 1e-8, _lib.ptr(adv), _lib.stream_ptr(self.device)), "frl_adv_norm")
        return adv, vt

    def save(self, model_dir):
        torch.save({name: {k: v.detach().clone().cpu() for k, v in agent.actor.state_dict().items()} for name, agent in self.agents.items()},
                   os.path.join(model_dir, 'IPPO.pth'))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = IPPO(dim_info, is_continue=is_continue, actor_lr=0, critic_lr=0, horizon=0, device=device, trick=trick)
        data = torch.load(os.path.join(model_dir, 'IPPO.pth'), map_location=device)
        for agent_id, agent in policy.agents.items():
            agent.actor.load_state_dict(data[agent_id])
        return policy
