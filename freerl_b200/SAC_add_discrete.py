"""``SAC_file/SAC_add_discrete.py`` — SAC with an added hand-made discrete variant.

For continuous action spaces the file's class is ``SAC_file/SAC.py``'s, statement for statement: running the unmodified reference
class on the seeds of ``oracle/make_golden.py::gen_sac`` reproduces ``tests/golden/sac.npz`` bit for bit
(``tests/test_launcher.py::test_sac_add_discrete_continuous_is_sac``), so the continuous case IS :class:`freerl_b200.SAC.SAC`
on the fused actor-critic kernel.

The discrete ``hands_on`` variant (``SAC_add_discrete.py:137-177, 226-360``) runs on its own fused kernel (``frl_sacd_learn``,
``csrc/algo_sacd.cuh``): softmax actor ``obs -> 128 -> 128 -> n_actions``, twin critic heads ``obs -> 128 -> 128 -> n_actions``
(``Critic_discrete_hands_on``: l1-l3, l4-l6), target ``r + gamma (1 - d) (sum_a p'_a min(Q1t, Q2t)_a + alpha H(p'))`` with ``p'`` from the
ONLINE actor like upstream, actor loss ``mean(-sum_a p_a min(Q1, Q2)_a - alpha H(p))``, ``log(p + 1e-8)``, target entropy
``0.6 log(n_actions)``, both optimisers clip the global norm at 0.5, Polyak on critic and actor targets.  The replay stores the action
index in one column; ``select_action`` samples ``Categorical(probs)`` (torch.multinomial's exponential trick, like the PPO discrete head),
``evaluate_action`` is the arg-max probability.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer
from .nets import DeviceNet, alias_module, bind_module
from .SAC import SAC as _SAC, Alpha


class _DiscreteNetInit(nn.Module):
    """``Actor_discrete_hands_on`` (n_heads = 1: l1-l3) / ``Critic_discrete_hands_on`` (n_heads = 2: l1-l6), nn.Linear in reference order"""

    def __init__(self, obs_dim, action_dim, n_heads, hidden=128):
        super().__init__()
        for h in range(n_heads):
            setattr(self, "l%d" % (3 * h + 1), nn.Linear(obs_dim, hidden))
            setattr(self, "l%d" % (3 * h + 2), nn.Linear(hidden, hidden))
            setattr(self, "l%d" % (3 * h + 3), nn.Linear(hidden, action_dim))


class _DiscreteAgent:
    def __init__(self, obs_dim, action_dim, actor_lr, critic_lr, device):
        a_dims = [(obs_dim, 128), (128, 128), (128, action_dim)]
        c_dims = a_dims * 2
        self.actor_names, self.critic_names = ("l1", "l2", "l3"), tuple("l%d" % (i + 1) for i in range(6))
        self._actor, self._critic = DeviceNet(a_dims, device, True), DeviceNet(c_dims, device, True)
        self._actor_t, self._critic_t = DeviceNet(a_dims, device, False), DeviceNet(c_dims, device, False)
        self.actor = bind_module(self._actor, _DiscreteNetInit(obs_dim, action_dim, 1), self.actor_names)      # SAC_add_discrete.py:187-188
        self.critic = bind_module(self._critic, _DiscreteNetInit(obs_dim, action_dim, 2), self.critic_names)
        self._actor_t.copy_from(self._actor)
        self._critic_t.copy_from(self._critic)
        self.actor_target = alias_module(self._actor_t, self.actor_names)
        self.critic_target = alias_module(self._critic_t, self.critic_names)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.actor_step = self.critic_step = 0


class _DiscreteSAC:
    """The ``is_continue=False`` branch of ``SAC_add_discrete.SAC``."""

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        obs_dim, action_dim = dim_info
        self.discrete_type = {'hands_on': True, 'other': False}
        print(self.discrete_type)
        self.trick = trick if trick is not None else {}
        self.device = _lib.require_device(device)
        self._bon = bool(self.trick.get("Batch_ObsNorm", False))
        if self._bon:           # SAC_add_discrete.py:240-241 (the script's default trick set turns it on)
            from .normalization import Normalization_batch_size
            self.batch_size_obs_norm = Normalization_batch_size(shape=obs_dim, device=self.device)
        self.obs_dim, self.action_dim, self.is_continue = obs_dim, action_dim, False
        self.agent = _DiscreteAgent(obs_dim, action_dim, actor_lr, critic_lr, self.device)
        self.buffer = Buffer(buffer_size, obs_dim, act_dim=1, device=self.device)
        self.mode = _common.resolve_mode(mode)
        self._scratch = _common.DeviceScratch(self.device, max(self.agent._actor.n_p, self.agent._critic.n_p))
        self._seed = _common.default_seed()
        self._n_learn = self._n_act = 0
        self.adaptive_alpha = True
        self.alphas = Alpha(action_dim, self.device, alpha=0.01, requires_grad=True, is_continue=False)
        self.alphas.target_entropy = float(0.6 * (-torch.log(torch.tensor(1.0 / action_dim))))          # float32 like upstream (:214)
        self.last_metrics = None

    # ---- acting ------------------------------------------------------------------------------------
    def select_action(self, obs, *, noise=None):
        """``Categorical(probs=actor(obs)).sample()`` (:253-255): int64 scalar, or [N] for a batch of observations."""
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        n = x.shape[0]
        if noise is None and self.mode == "parity":
            noise = torch.empty((n, self.action_dim), dtype=torch.float32, device=self.device).exponential_(1)
        elif noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32).to(self.device).reshape(n, self.action_dim).contiguous()
        self._n_act += 1
        out = _common.infer(self.agent._actor, x, _lib.INFER_PPO_CAT, self.device, 2, noise=noise, seed=self._seed, counter=self._n_act,
                            obs_norm=self.batch_size_obs_norm if self._bon else None).cpu().numpy()
        a = out[:, 0].astype(np.int64)
        return a[0] if single else a

    def evaluate_action(self, obs):
        """arg-max probability; like upstream (:264-266) WITHOUT the Batch_ObsNorm normalisation select_action applies"""
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._actor, x, _lib.INFER_ARGMAX, self.device, 1).reshape(-1).to(torch.int64).cpu().numpy()
        return a[0] if single else a

    # ---- buffer ------------------------------------------------------------------------------------
    def add(self, obs, action, reward, next_obs, done):
        self.buffer.add(obs, action, reward, next_obs, done)

    def sample(self, batch_size):
        total_size = len(self.buffer)
        indices = np.random.choice(total_size, min(total_size, batch_size), replace=False)
        obs, actions, rewards, next_obs, dones = self.buffer.sample(indices)
        if self._bon:
            obs = self.batch_size_obs_norm(obs)
            next_obs = self.batch_size_obs_norm(next_obs, update=False)
        return obs, actions, rewards, next_obs, dones

    # ---- learning ----------------------------------------------------------------------------------
    def learn(self, batch_size, gamma, tau, *, n_updates=1, indices=None):
        total = len(self.buffer)
        if total == 0:
            raise RuntimeError("learn() called on an empty replay buffer")
        B = min(total, batch_size)
        if indices is None:
            idx = _common.make_indices(self.mode, total, B, n_updates, self.device, self._seed, self._n_learn)
        else:
            idx = self.buffer._indices_to_device(indices).reshape(n_updates, -1)
            B = idx.shape[1]
        if self._bon:
            # Batch_ObsNorm (:284-286): the running statistics are folded with the mean of EVERY sampled batch and the batch is normalised
            # before it reaches the networks.  The kernel gathers its rows itself, so each learn is staged: rows gathered, statistics
            # updated, normalised rows packed into a B-row staging replay, one learn on that.
            if getattr(self, "_stage_buf", None) is None or self._stage_buf.capacity < B:
                self._stage_buf = Buffer(B, self.obs_dim, act_dim=1, device=self.device)
                self._stage_idx = torch.arange(B, dtype=torch.int64, device=self.device)
            for u in range(n_updates):
                obs, act, rew, nobs, done = self.buffer.sample(idx[u])
                obs = self.batch_size_obs_norm(obs)
                nobs = self.batch_size_obs_norm(nobs, update=False)
                st = self._stage_buf
                st._index, st._size = 0, 0
                st.add_device(obs.contiguous(), act, rew.reshape(-1), nobs.contiguous(), done.reshape(-1))
                self._learn_on(st, self._stage_idx[:B].reshape(1, B), B, 1, gamma, tau)
            return
        self._learn_on(self.buffer, idx, B, n_updates, gamma, tau)

    def _learn_on(self, buffer, idx, B, n_updates, gamma, tau):
        ag, al = self.agent, self.alphas
        a = _lib.SacdArgs()
        a.actor, a.actor_target = ag._actor.c_struct(), ag._actor_t.c_struct()
        a.critic, a.critic_target = ag._critic.c_struct(), ag._critic_t.c_struct()
        a.replay = buffer.c_struct()
        a.indices, a.B, a.n_updates = idx.data_ptr(), B, n_updates
        a.gamma, a.tau = gamma, tau
        a.lr_actor, a.lr_critic, a.beta1, a.beta2, a.eps = ag.actor_lr, ag.critic_lr, 0.9, 0.999, 1e-8
        a.max_norm = 0.5
        a.step_actor0, a.step_critic0 = ag.actor_step, ag.critic_step
        a.alpha_state, a.adaptive_alpha = al.state.data_ptr(), int(self.adaptive_alpha)
        a.alpha_lr, a.target_entropy, a.step_alpha0 = al.alpha_lr, float(al.target_entropy), al.step
        out = self._scratch.out(n_updates, self.device)
        a.gpart, a.sumsq = self._scratch.gpart.data_ptr(), self._scratch.sumsq.data_ptr()
        a.stats, a.out = self._scratch.stats.data_ptr(), out.data_ptr()
        _lib.check(_lib.lib().frl_sacd_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_sacd_learn")
        self.last_metrics = out[:n_updates]
        self._keepalive = (idx,)
        ag.actor_step += n_updates
        ag.critic_step += n_updates
        al.step += n_updates
        self._n_learn += n_updates

    def update_target(self, tau):
        for t, s in ((self.agent._critic_t, self.agent._critic), (self.agent._actor_t, self.agent._actor)):
            t.p.mul_(1.0 - tau).add_(s.p * tau)
            t.sync_mirror()

    # ---- checkpoint ---------------------------------------------------------------------------------
    def save(self, model_dir):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.actor.state_dict().items()}, os.path.join(model_dir, "SAC.pt"))


class SAC(_SAC):
    def __new__(cls, dim_info=None, is_continue=True, *args, **kw):
        if cls is SAC and not is_continue:
            return _DiscreteSAC(dim_info, is_continue, *args, **kw)
        return object.__new__(cls)

    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        self.discrete_type = {'hands_on': True, 'other': False}
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=trick, mode=mode)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        if is_continue:
            return _SAC.load(dim_info, is_continue, model_dir, trick=trick, device=device)
        device = device if device is not None else torch.device("cuda")
        policy = _DiscreteSAC(dim_info, is_continue, 0, 0, 0, device=device, trick=trick)
        policy.agent.actor.load_state_dict(torch.load(os.path.join(model_dir, "SAC.pt"), map_location=device))
        return policy
