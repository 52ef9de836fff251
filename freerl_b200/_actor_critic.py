"""Shared host logic of the off-policy actor-critic classes (SAC / TD3 / DDPG) over ``frl_ac_learn``."""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer
from .nets import DeviceNet, alias_module, bind_module
from .normalization import Normalization_batch_size


class _ActorInit(nn.Module):
    """nn.Linear construction in reference order (l1, l2, head) — ``SAC.py:61-66``, ``TD3.py:53-57``, ``DDPG.py:71-75``."""

    def __init__(self, obs_dim, action_dim, head_name, hidden=128):
        super().__init__()
        self.l1 = nn.Linear(obs_dim, hidden)
        self.l2 = nn.Linear(hidden, hidden)
        setattr(self, head_name, nn.Linear(hidden, action_dim))


class _CriticInit(nn.Module):
    """``SAC.py:104-114`` / ``TD3.py:87-101`` (twin: l1-l3, l4-l6) or ``DDPG.py:90-96`` (single: l1-l3)."""

    def __init__(self, in_dim, n_heads, hidden=128):
        super().__init__()
        for h in range(n_heads):
            setattr(self, "l%d" % (3 * h + 1), nn.Linear(in_dim, hidden))
            setattr(self, "l%d" % (3 * h + 2), nn.Linear(hidden, hidden))
            setattr(self, "l%d" % (3 * h + 3), nn.Linear(hidden, 1))


class ACAgent:
    """Reference ``Agent`` (``SAC.py:129-151``): actor, critic, deep-copied targets, two Adam optimisers."""

    def __init__(self, obs_dim, action_dim, actor_lr, critic_lr, device, n_heads, sac, post_init=None):
        head = "mean_layer" if sac else "l3"
        a_dims = [(obs_dim, 128), (128, 128), (128, action_dim)]
        c_dims = [(obs_dim + action_dim, 128), (128, 128), (128, 1)] * n_heads
        self.actor_names = ("l1", "l2", head)
        self.critic_names = tuple("l%d" % (i + 1) for i in range(3 * n_heads))
        self.extra = "log_std" if sac else None
        self._actor = DeviceNet(a_dims, device, True, x_len=action_dim if sac else 0)
        self._critic = DeviceNet(c_dims, device, True)
        self._actor_t = DeviceNet(a_dims, device, False, x_len=action_dim if sac else 0)
        self._critic_t = DeviceNet(c_dims, device, False)
        a_init = _ActorInit(obs_dim, action_dim, head)
        if sac:
            a_init.log_std = nn.Parameter(torch.zeros(1, action_dim))          # SAC.py:65
        if post_init:
            post_init(a_init, self.actor_names)          # re-init happens inside the module constructor in the reference
        c_init = _CriticInit(obs_dim + action_dim, n_heads)
        if post_init:
            post_init(c_init, self.critic_names)
        self.actor = bind_module(self._actor, a_init, self.actor_names, self.extra)
        self.critic = bind_module(self._critic, c_init, self.critic_names)
        self._actor_t.copy_from(self._actor)
        self._critic_t.copy_from(self._critic)
        self.actor_target = alias_module(self._actor_t, self.actor_names, self.extra)
        self.critic_target = alias_module(self._critic_t, self.critic_names)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.actor_step = 0
        self.critic_step = 0


class ACBase:
    """Common ``add / sample / learn`` plumbing; subclasses fill the algorithm-specific kernel arguments."""

    n_heads = 2
    sac = False

    def _setup(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, mode, post_init=None,
               batch_obs_norm=False):
        obs_dim, action_dim = dim_info
        self.device = _lib.require_device(device)
        self._bon = bool(batch_obs_norm)
        if self._bon:
            self.batch_size_obs_norm = Normalization_batch_size(shape=obs_dim, device=self.device)
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.is_continue = is_continue
        self.agent = ACAgent(obs_dim, action_dim, actor_lr, critic_lr, self.device, self.n_heads, self.sac, post_init)
        self.buffer = Buffer(buffer_size, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.mode = _common.resolve_mode(mode)
        self._scratch = _common.DeviceScratch(self.device, max(self.agent._actor.n_p, self.agent._critic.n_p))
        self._seed = _common.default_seed()
        self._n_learn = 0
        self._n_act = 0
        self.last_metrics = None

    # ---- buffer ------------------------------------------------------------------------------------
    def add(self, obs, action, reward, next_obs, done):
        self.buffer.add(obs, action, reward, next_obs, done)

    def sample(self, batch_size):
        total_size = len(self.buffer)
        batch_size = min(total_size, batch_size)
        indices = np.random.choice(total_size, batch_size, replace=False)
        obs, actions, rewards, next_obs, dones = self.buffer.sample(indices)
        if self._bon:
            obs = self.batch_size_obs_norm(obs)
            next_obs = self.batch_size_obs_norm(next_obs, update=False)
        return obs, actions, rewards, next_obs, dones

    def _obs_norm(self):
        return self.batch_size_obs_norm if self._bon else None

    # ---- kernel call -------------------------------------------------------------------------------
    def _base_args(self, batch_size, gamma, tau, n_updates, indices):
        total = len(self.buffer)
        if total == 0:
            raise RuntimeError("learn() called on an empty replay buffer")
        B = min(total, batch_size)
        if indices is None:
            idx = _common.make_indices(self.mode, total, B, n_updates, self.device, self._seed, self._n_learn)
        else:
            idx = self.buffer._indices_to_device(indices).reshape(n_updates, -1)
            B = idx.shape[1]
        ag = self.agent
        a = _lib.AcArgs()
        a.actor, a.actor_target = ag._actor.c_struct(), ag._actor_t.c_struct()
        a.critic, a.critic_target = ag._critic.c_struct(), ag._critic_t.c_struct()
        a.n_heads = self.n_heads
        a.actor_kind = _lib.ACTOR_SAC if self.sac else _lib.ACTOR_TANH
        a.replay = self.buffer.c_struct()
        a.indices, a.B, a.n_updates = idx.data_ptr(), B, n_updates
        a.seed = self._seed
        a.gamma, a.tau = gamma, tau
        a.lr_actor, a.lr_critic = ag.actor_lr, ag.critic_lr
        a.beta1, a.beta2, a.eps, a.wd_critic = 0.9, 0.999, 1e-8, 0.0
        a.max_norm = 0.5                                   # clip_grad_norm_(…, 0.5): SAC.py:144,150 etc.
        a.step_actor0, a.step_critic0, a.total_it0 = ag.actor_step, ag.critic_step, self._n_learn
        a.policy_freq, a.target_smoothing = 1, 0
        a.max_action, a.policy_noise_scale = 1.0, 1.0
        out = self._scratch.out(n_updates, self.device)
        a.gpart, a.sumsq = self._scratch.gpart.data_ptr(), self._scratch.sumsq.data_ptr()
        a.stats, a.out = self._scratch.stats.data_ptr(), out.data_ptr()
        a.xchg = self._scratch.xchg(B, self.device).data_ptr()
        if self._bon:                                      # running statistics folded in by the kernel, once per learn
            a.obs_norm[0] = self.batch_size_obs_norm.data_ptr()
            a.obs_norm_n0 = self.batch_size_obs_norm.running_ms.n
        return a, idx, B, out

    def _fx_scratch(self, a):
        """Scratch of the small-batch schedule (``csrc/algo_acfx.cuh``): exchange blocks sized by the library for these shapes
        (0 floats = not eligible -> the generic kernel runs) and the barrier / flag words.  Call once every argument is set."""
        need = int(_lib.lib().frl_ac_ws_floats(ctypes.byref(a)))
        if need <= 0:
            return
        ws = getattr(self, "_fx_ws", None)
        if ws is None or ws.numel() < need:
            self._fx_ws = ws = torch.zeros(need, dtype=torch.float32, device=self.device)
            self._fx_sync = torch.zeros(4096, dtype=torch.int32, device=self.device)
        a.ws, a.sync = ws.data_ptr(), self._fx_sync.data_ptr()

    def _launch(self, a, keep, n_updates, out):
        if a.n_agents <= 1:
            self._fx_scratch(a)
        self.last_path = int(_lib.lib().frl_ac_path(ctypes.byref(a)))       # 1: small-batch schedule, 0: generic kernel
        _lib.check(_lib.lib().frl_ac_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_ac_learn")
        self.last_metrics = out[:n_updates]
        self._keepalive = keep
        if self._bon:
            self.batch_size_obs_norm.running_ms.n += n_updates

    def _noise(self, given, n_updates, B):
        """parity mode: draw [n_updates, B, act] like the reference would, one tensor per learn, in order."""
        if given is not None:
            return torch.as_tensor(given, dtype=torch.float32).to(self.device).reshape(n_updates, B, self.action_dim).contiguous()
        return None

    # ---- checkpoint ---------------------------------------------------------------------------------
    def _save_actor(self, path):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.actor.state_dict().items()}, path)

    def update_target(self, tau):
        for t, s in ((self.agent._critic_t, self.agent._critic), (self.agent._actor_t, self.agent._actor)):
            t.p.mul_(1.0 - tau).add_(s.p * tau)
            t.sync_mirror()

    # ---- multi-GPU: replicas with sharded env workers / replay (off-policy) ---------------------------
    def enable_replica_sync(self, group=None):
        """One process per GPU, each with its own env shard and replay shard (no data-path collective).  The replicas stay ONE
        policy by averaging parameters over ``torch.distributed`` (NCCL over NVLink on the GPU box): rank 0's parameters are
        broadcast now, ``sync_replicas()`` averages them afterwards (call it every K learns; the bench uses K = one vector step).
        Adam moments stay local — each replica's optimiser sees its own shard's gradients, the standard local-SGD scheme."""
        import torch.distributed as dist
        self._rs = (dist, group, dist.get_world_size(group))
        for n_ in self._replica_nets():
            dist.broadcast(n_.p, src=0, group=group)
            n_.sync_mirror()
        if self.sac:
            dist.broadcast(self.alphas.state, src=0, group=group)
        self._rs_peers = _common.replica_peers(dist, group, self.device, self._replica_tensors())
        self.replica_collective = ("in-kernel peer-memory average (frl_replica_average), one launch per sync" if self._rs_peers
                                   else "dist.all_reduce + divide per parameter block")

    def _replica_tensors(self):
        ts = [n_.p for n_ in self._replica_nets()]
        if self.sac:
            ts.append(self.alphas.state[:1])          # log_alpha (the Adam moments of alpha stay local like the others)
        return ts

    def _replica_nets(self):
        ag = self.agent
        return (ag._actor, ag._critic, ag._actor_t, ag._critic_t)

    def sync_replicas(self):
        """parameter average over the replicas (all-reduce sum, scale by 1/world, refresh the transposed mirrors)"""
        rs = getattr(self, "_rs", None)
        if rs is None or rs[2] == 1:
            return
        dist, group, world = rs
        if getattr(self, "_rs_peers", None) is not None:        # one cooperative launch over NVLink peer memory, then the mirrors
            _common.replica_average(self._rs_peers, self._replica_tensors(), self.device)
            for n_ in self._replica_nets():
                n_.sync_mirror()
            return
        for n_ in self._replica_nets():
            dist.all_reduce(n_.p, group=group)
            n_.p.div_(world)
            n_.sync_mirror()
        if self.sac:                      # log_alpha (the Adam moments of alpha stay local like the others)
            la = self.alphas.state[:1]
            dist.all_reduce(la, group=group)
            la.div_(world)
