"""PPO with the reference's class API (``PPO_file/PPO.py:58-297``) on the fused B200 kernels.

``PPO(dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None)``; ``select_action -> (action,
log_prob)``, ``evaluate_action``, ``add(obs, action, reward, next_obs, done, action_log_pi, adv_dones)``,
``learn(minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient)``, ``save``/``load``.
Reference behaviour kept: ONE cautious-AdamW (eps 1e-6) over ``actor.parameters() + critic.parameters()`` with
lr = ``actor_lr`` (``critic_lr`` unused, PPO.py:121); separate ``clip_grad_norm_(…, 0.5)`` for actor and critic;
discrete head returns logits; GAE reverse scan with ``adv_done`` masking and ``v_target = adv + V(s)``; learn()
always consumes the full ``horizon`` arrays and clears the buffer.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer_for_PPO
from .nets import DeviceNet, _Shim


class _ActorInit(nn.Module):
    def __init__(self, obs_dim, action_dim, head, hidden=128):
        super().__init__()
        self.l1 = nn.Linear(obs_dim, hidden)
        self.l2 = nn.Linear(hidden, hidden)
        setattr(self, head, nn.Linear(hidden, action_dim))


class _CriticInit(nn.Module):
    def __init__(self, obs_dim, hidden=128):
        super().__init__()
        self.l1 = nn.Linear(obs_dim, hidden)
        self.l2 = nn.Linear(hidden, hidden)
        self.l3 = nn.Linear(hidden, 1)


def _shim(net, layer_ids, names, extra_name=None):
    shim = _Shim()
    if extra_name:
        shim.register_parameter(extra_name, nn.Parameter(net.extra().view(1, -1), requires_grad=False))
    shim._net = net
    for li, name in zip(layer_ids, names):
        lin = nn.Module()
        lin.weight = nn.Parameter(net.weight(li), requires_grad=False)
        lin.bias = nn.Parameter(net.bias(li), requires_grad=False)
        shim.add_module(name, lin)
    shim.register_load_state_dict_post_hook(lambda module, incompatible: net.sync_mirror())
    return shim


class Agent:
    """``PPO.py:109-152``: actor + critic with ONE merged optimiser -> ONE device parameter block (layers 0-2 actor,
    3-5 critic, extra = log_std)."""

    def __init__(self, obs_dim, action_dim, actor_lr, critic_lr, is_continue, device, critic_in=None, init_hook=None, beta=False):
        head = "mean_layer" if is_continue else "l3"
        critic_in = obs_dim if critic_in is None else critic_in
        self.beta = bool(beta and is_continue)
        if self.beta:
            self._init_beta(obs_dim, action_dim, critic_in, device, init_hook)
            self.lr, self.step = actor_lr, 0
            return
        dims = [(obs_dim, 128), (128, 128), (128, action_dim), (critic_in, 128), (128, 128), (128, 1)]
        self._net = DeviceNet(dims, device, True, x_len=action_dim if is_continue else 0)
        a_init = _ActorInit(obs_dim, action_dim, head)          # same RNG consumption as Actor / Actor_discrete
        if init_hook:
            init_hook(a_init, ("l1", "l2", head), "actor")          # re-initialisation inside the module constructor
        c_init = _CriticInit(critic_in)
        if init_hook:
            init_hook(c_init, ("l1", "l2", "l3"), "critic")
        with torch.no_grad():
            for li, lay in enumerate((a_init.l1, a_init.l2, getattr(a_init, head), c_init.l1, c_init.l2, c_init.l3)):
                self._net.weight(li).copy_(lay.weight.to(device))
                self._net.bias(li).copy_(lay.bias.to(device))
        self._net.sync_mirror()
        self.actor_names = ("l1", "l2", head)
        self.actor = _shim(self._net, (0, 1, 2), self.actor_names, "log_std" if is_continue else None)
        self.critic = _shim(self._net, (3, 4, 5), ("l1", "l2", "l3"))
        self.lr = actor_lr                                       # AdamW(actor+critic params, lr=actor_lr)  PPO.py:121
        self.step = 0

    def _init_beta(self, obs_dim, action_dim, critic_in, device, init_hook):
        """``Actor_Beta`` (``PPO_with_tricks.py:120-150``): l1, l2, alpha_layer, beta_layer.  The two heads are ONE device layer of width
        2 * action_dim (rows [alpha | beta]); the shim exposes them as ``alpha_layer`` / ``beta_layer`` row slices, in the reference's
        state-dict order.  No ``log_std``."""
        A = action_dim
        dims = [(obs_dim, 128), (128, 128), (128, 2 * A), (critic_in, 128), (128, 128), (128, 1)]
        self._net = net = DeviceNet(dims, device, True, x_len=0)
        a_init = nn.Module()
        a_init.l1, a_init.l2 = nn.Linear(obs_dim, 128), nn.Linear(128, 128)
        a_init.alpha_layer, a_init.beta_layer = nn.Linear(128, A), nn.Linear(128, A)
        if init_hook:
            init_hook(a_init, ("l1", "l2", "alpha_layer", "beta_layer"), "actor")
        c_init = _CriticInit(critic_in)
        if init_hook:
            init_hook(c_init, ("l1", "l2", "l3"), "critic")
        with torch.no_grad():
            for li, lay in ((0, a_init.l1), (1, a_init.l2), (3, c_init.l1), (4, c_init.l2), (5, c_init.l3)):
                net.weight(li).copy_(lay.weight.to(device))
                net.bias(li).copy_(lay.bias.to(device))
            net.weight(2)[:A].copy_(a_init.alpha_layer.weight.to(device)); net.weight(2)[A:].copy_(a_init.beta_layer.weight.to(device))
            net.bias(2)[:A].copy_(a_init.alpha_layer.bias.to(device)); net.bias(2)[A:].copy_(a_init.beta_layer.bias.to(device))
        net.sync_mirror()
        self.actor_names = ("l1", "l2", "alpha_layer", "beta_layer")
        shim = _shim(net, (0, 1), ("l1", "l2"))
        for name, sl in (("alpha_layer", slice(0, A)), ("beta_layer", slice(A, 2 * A))):
            lin = nn.Module()
            lin.weight = nn.Parameter(net.weight(2)[sl], requires_grad=False)
            lin.bias = nn.Parameter(net.bias(2)[sl], requires_grad=False)
            shim.add_module(name, lin)
        self.actor = shim
        self.critic = _shim(net, (3, 4, 5), ("l1", "l2", "l3"))


class PPO:
    optimizer = _lib.OPT_CAUTIOUS_ADAMW
    adam_eps = 1e-6
    max_norm = 0.5
    hidden_tanh = 0              # PPO_with_tricks' `tanh` switch, bit 0: actor, bit 1: critic use tanh instead of ReLU hidden activations

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, horizon, device, trick=None, mode=None):
        obs_dim, action_dim = dim_info
        self.device = _lib.require_device(device)
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.agent = Agent(obs_dim, action_dim, actor_lr, critic_lr, is_continue, self.device, init_hook=getattr(self, "_init_hook", None),
                           beta=getattr(self, "_beta", False))
        self.buffer = Buffer_for_PPO(horizon, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.is_continue = is_continue
        print('actor_type:continue') if self.is_continue else print('actor_type:discrete')
        self.horizon = int(horizon)
        self.trick = trick
        self.mode = _common.resolve_mode(mode)
        self._seed = _common.default_seed()
        self._n_act = 0
        sm = _lib.sm_count()
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self._gpart, self._sumsq = z(sm, self.agent._net.n_p), z(sm, 2)
        self._segcnt, self._stats = z(sm, _lib.NSEG), z(sm, 8)
        self.last_metrics = None

    # ---- acting ------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        n = x.shape[0]
        net = self.agent._net
        self._n_act += 1
        if self.is_continue and getattr(self.agent, "beta", False):
            # Beta(alpha, beta).sample() (PPO_with_tricks.py:238-240, 246-247): the kernel evaluates the network, the draw and its
            # log-prob are torch's own (Dirichlet / gamma sampler — on the CPU generator in parity mode, like the reference)
            z = _common.infer(net, x, _lib.INFER_RAW, self.device, 2 * self.action_dim, l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1),
                              obs_norm=getattr(self, "batch_size_obs_norm", None))
            if self.mode == "parity":
                z = z.cpu()
            al = torch.nn.functional.softplus(z[:, :self.action_dim]) + 1.0
            be = torch.nn.functional.softplus(z[:, self.action_dim:]) + 1.0
            dist = torch.distributions.Beta(al, be)
            action = dist.sample() if noise is None else torch.as_tensor(noise, dtype=torch.float32).reshape(n, self.action_dim).to(al.device)
            logp = dist.log_prob(action)
            action, logp = action.cpu().numpy(), logp.cpu().numpy()
            return (action[0], logp[0]) if single else (action, logp)
        if self.is_continue:
            if noise is None and self.mode == "parity":
                noise = _common.reference_randn((n, self.action_dim), self.device)    # dist.sample() (PPO.py:173)
            elif noise is not None:
                noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).reshape(n, self.action_dim).contiguous()
            out = _common.infer(net, x, _lib.INFER_PPO_GAUSS, self.device, 2 * self.action_dim, noise=noise, seed=self._seed,
                                counter=self._n_act, l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1),
                                obs_norm=getattr(self, "batch_size_obs_norm", None)).cpu().numpy()
            action, logp = out[:, :self.action_dim], out[:, self.action_dim:]
            return (action[0], logp[0]) if single else (action, logp)
        if noise is None and self.mode == "parity":
            # Categorical.sample() -> torch.multinomial(probs, 1): q = empty_like(probs).exponential_(1); argmax(probs / q)
            noise = torch.empty((n, self.action_dim), dtype=torch.float32, device=self.device).exponential_(1)
        elif noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).reshape(n, self.action_dim).contiguous()
        out = _common.infer(net, x, _lib.INFER_PPO_CAT, self.device, 2, noise=noise, seed=self._seed, counter=self._n_act,
                            l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1), obs_norm=getattr(self, "batch_size_obs_norm", None)).cpu().numpy()
        action, logp = out[:, 0].astype(np.int64), out[:, 1]
        return (action[0], logp[0]) if single else (action, logp)

    def evaluate_action(self, obs):
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        net = self.agent._net
        if self.is_continue and getattr(self.agent, "beta", False):          # 2 (alpha / (alpha + beta) - 0.5)   PPO_with_tricks.py:259-262
            z = _common.infer(net, x, _lib.INFER_RAW, self.device, 2 * self.action_dim, l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1))
            al = torch.nn.functional.softplus(z[:, :self.action_dim]) + 1.0
            be = torch.nn.functional.softplus(z[:, self.action_dim:]) + 1.0
            a = (2.0 * (al / (al + be) - 0.5)).cpu().numpy()
            return a[0] if single else a
        if self.is_continue:
            a = _common.infer(net, x, _lib.INFER_TANH, self.device, self.action_dim, l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1)).cpu().numpy()
            return a[0] if single else a
        a = _common.infer(net, x, _lib.INFER_ARGMAX, self.device, 1, l0=0, nl=3, hidden_tanh=bool(self.hidden_tanh & 1)).reshape(-1).to(torch.int64).cpu().numpy()
        return a[0] if single else a

    def add(self, obs, action, reward, next_obs, done, action_log_pi, adv_dones):
        self.buffer.add(obs, action, reward, next_obs, done, action_log_pi, adv_dones)

    # ---- learning ----------------------------------------------------------------------------------
    def _values(self, obs):
        return _common.infer(self.agent._net, obs, _lib.INFER_RAW, self.device, 1, l0=3, nl=3, hidden_tanh=bool(self.hidden_tanh & 2))

    def compute_gae(self, gamma, lmbda):
        """critic(obs), critic(next_obs) -> fused TD-delta + GAE scan (PPO.py:222-233) -> (adv, v_target) [M,1]."""
        b = self.buffer
        M = b.capacity
        N = b.n_envs if (M % max(b.n_envs, 1) == 0) else 1
        T = M // N
        vs, vs_ = self._values(b.obs), self._values(b.next_obs)
        adv = torch.empty((M, 1), dtype=torch.float32, device=self.device)
        vt = torch.empty((M, 1), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().frl_gae(_lib.ptr(b.rewards), _lib.ptr(b.dones), _lib.ptr(b.adv_dones), _lib.ptr(vs), _lib.ptr(vs_),
                                      T, N, float(gamma), float(lmbda), _lib.ptr(adv), _lib.ptr(vt), _lib.stream_ptr(self.device)),
                   "frl_gae")
        return adv, vt

    def _minibatch_plan(self, minibatch_size, K_epochs, permutations):
        """Reference: per epoch ``np.random.permutation(horizon)`` sliced into minibatches (PPO.py:247-248)."""
        H = self.horizon
        if permutations is None and self.mode != "parity":
            return _common.device_minibatch_plan(H, minibatch_size, K_epochs, self.device, self._seed + self.agent.step)
        if permutations is None:
            if self.mode == "parity":
                permutations = [np.random.permutation(H) for _ in range(K_epochs)]
            else:
                g = torch.Generator(device="cpu")
                g.manual_seed((self._seed + self.agent.step) & 0x7FFFFFFF)
                permutations = [torch.randperm(H, generator=g).numpy() for _ in range(K_epochs)]
        nmb = (H + minibatch_size - 1) // minibatch_size
        idx = np.zeros((K_epochs * nmb, minibatch_size), np.int64)
        rows = np.zeros(K_epochs * nmb, np.int32)
        for k, perm in enumerate(permutations):
            for j in range(nmb):
                sl = np.asarray(perm[j * minibatch_size:(j + 1) * minibatch_size])
                idx[k * nmb + j, :sl.size] = sl
                rows[k * nmb + j] = sl.size
        return torch.from_numpy(idx).to(self.device), torch.from_numpy(rows).to(self.device), idx.shape[0]

    def _update(self, adv, v_target, minibatch_size, K_epochs, clip_param, entropy_coefficient, permutations, n_adv=1):
        b, ag = self.buffer, self.agent
        idx, rows, n_updates = self._minibatch_plan(minibatch_size, K_epochs, permutations)
        out = torch.zeros((n_updates, 8), dtype=torch.float32, device=self.device)
        a = _lib.PpoArgs()
        a.net = ag._net.c_struct()
        a.continuous = 2 if getattr(self.agent, "beta", False) else int(self.is_continue)
        a.obs, a.action, a.logp_old = b.obs.data_ptr(), b.actions.data_ptr(), b.action_log_probs.data_ptr()
        a.adv, a.v_target = adv.data_ptr(), v_target.data_ptr()
        a.M, a.obs_dim, a.act_cols, a.logp_cols, a.n_adv = b.capacity, self.obs_dim, b.act_dim, b.logp_dim, n_adv
        a.indices, a.mb_rows, a.mb, a.n_updates = idx.data_ptr(), rows.data_ptr(), minibatch_size, n_updates
        a.clip_param, a.entropy_coef = clip_param, entropy_coefficient
        a.max_norm_actor = a.max_norm_critic = self.max_norm
        a.optimizer = self.optimizer
        a.lr, a.beta1, a.beta2, a.eps = ag.lr, 0.9, 0.999, self.adam_eps
        a.hidden_tanh = int(self.hidden_tanh)
        a.lr_critic = float(getattr(ag, "lr_critic", 0.0))        # separate actor / critic Adams (PPO_advance); 0 = merged
        a.step0 = ag.step
        a.gpart, a.sumsq, a.segcnt = self._gpart.data_ptr(), self._sumsq.data_ptr(), self._segcnt.data_ptr()
        a.umma_ws = _common.umma_ws_ptr(self.device, int(a.mb))
        a.stats, a.out = self._stats.data_ptr(), out.data_ptr()
        self._launch_update(a, ag._net, n_updates)
        ag.step += n_updates
        self.last_metrics = out
        self._keep = (idx, rows, adv, v_target)

    # ---- data parallel (one process per GPU) ---------------------------------------------------------
    def enable_data_parallel(self, group=None, peer=None):
        """Synchronous data-parallel training over ``torch.distributed`` (one process per GPU): every rank holds its own env shard /
        rollout and an identical replica of the networks; each minibatch is the union of the ranks' equal-size sub-minibatches.
        Per optimiser step the flat gradient (net.g, 36 k floats) is summed over the ranks between the in-kernel cross-CTA
        reduction and the clip + (cautious-)AdamW stages and scaled by 1/world, so all replicas apply bit-identical updates
        (SURVEY §8e).  ``peer=True`` (default on CUDA): the sum runs INSIDE the one persistent launch over peer-mapped NVLink
        memory (``_common.DpPeers`` / the exchange stage of csrc/algo_ppo.cuh) — no NCCL call per step; ``peer=False`` (and the
        test-only host emulation): the host splits every step into launch [0,2) -> ``dist.all_reduce(net.g)`` -> launch [2,5)."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        self._dp = (dist, group, world)
        nets = [ag._net for ag in self.agents.values()] if hasattr(self, "agents") else [self.agent._net]
        for net in nets:
            dist.broadcast(net.p, src=0, group=group)
            net.sync_mirror()
        if peer is None:
            peer = (self.device.type == "cuda" and world > 1 and not _lib.lib().frl_is_emulation()
                    and os.environ.get("FREERL_B200_DP_PEER", "1") != "0")
        self._dp_peers = _common.DpPeers(dist, group, self.device, max(n.n_p for n in nets)) if (peer and world > 1) else None
        self.dp_collective = ("in-kernel peer-memory sum (NVLink P2P loads, flag hand-off), one launch per learn" if self._dp_peers
                              else "dist.all_reduce(net.g) between two launches per optimiser step")

    def _launch_update(self, a, net, n_updates):
        dp = getattr(self, "_dp", None)
        if dp is None or dp[2] == 1:
            _lib.check(_lib.lib().frl_ppo_update(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ppo_update")
            return
        dist, group, world = dp
        a.grad_scale = 1.0 / world
        peers = getattr(self, "_dp_peers", None)
        if peers is not None:                       # every optimiser step of the learn in ONE launch, gradients summed over NVLink
            peers.fill(a, net.n_p, n_updates)
            _lib.check(_lib.lib().frl_ppo_update(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ppo_update")
            return
        idx0, rows0, out0, step0 = a.indices, a.mb_rows, a.out, a.step0
        a.n_updates = 1
        for u in range(n_updates):
            a.indices = idx0 + u * a.mb * 8
            a.mb_rows = rows0 + u * 4
            a.out = out0 + u * 8 * 4
            a.step0 = step0 + u
            a.stage_lo, a.stage_hi = 0, 2
            _lib.check(_lib.lib().frl_ppo_update(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ppo_update")
            dist.all_reduce(net.g, group=group)                       # flat gradient buffer, sum over ranks
            a.stage_lo, a.stage_hi = 2, 5
            _lib.check(_lib.lib().frl_ppo_update(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ppo_update")

    def learn(self, minibatch_size, gamma, lmbda, clip_param, K_epochs, entropy_coefficient, *, permutations=None):
        adv, v_target = self.compute_gae(gamma, lmbda)
        self.last_adv, self.last_v_target = adv, v_target
        self._update(adv, v_target, minibatch_size, K_epochs, clip_param, entropy_coefficient, permutations)
        self.buffer.clear()

    # ---- checkpoint ---------------------------------------------------------------------------------
    def save(self, model_dir):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.actor.state_dict().items()}, os.path.join(model_dir, "PPO.pt"))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = PPO(dim_info, is_continue, 0, 0, 0, device=device, trick=trick)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "PPO.pt"), map_location=device))
        return policy
