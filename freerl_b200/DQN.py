"""DQN with the reference's class API (``DQN_file/DQN.py:32-138``) on the fused B200 kernel.

``DQN(dim_info, is_continue, Qnet_lr, buffer_size, device, trick=None)`` with ``select_action / evaluate_action /
add / sample / learn(batch_size, gamma, tau) / update_target / save / load`` — drop-in for the class inside the
reference train script.  ``learn`` = ``frl_dqn_learn``: gather -> target max -> TD target -> MSE -> backward ->
Adam -> Polyak in one persistent kernel.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _common, _lib
from .Buffer import Buffer
from .nets import DeviceNet, alias_module, bind_module


class MLP(nn.Module):
    """Same construction order as ``DQN_file/DQN.py:32-41`` (hidden 128) so torch's RNG stream is consumed identically."""

    def __init__(self, obs_dim, action_dim, hidden=128):
        super().__init__()
        self.l1 = nn.Linear(obs_dim, hidden)
        self.l2 = nn.Linear(hidden, action_dim)


class Agent:
    def __init__(self, obs_dim, action_dim, Qnet_lr, device):
        dims = [(obs_dim, 128), (128, action_dim)]
        self._q = DeviceNet(dims, device, trainable=True)
        self._qt = DeviceNet(dims, device, trainable=False)
        self.Qnet = bind_module(self._q, MLP(obs_dim, action_dim), ("l1", "l2"))
        self._qt.copy_from(self._q)                      # deepcopy(self.Qnet)  (DQN.py:52)
        self.Qnet_target = alias_module(self._qt, ("l1", "l2"))
        self.lr = Qnet_lr
        self.step = 0                                    # torch.optim.Adam step counter (DQN.py:54)


class DQN(_common.ReplicaSyncMixin):
    def _replica_pairs(self):
        return [(self.agent._q.p, self.agent._q.sync_mirror), (self.agent._qt.p, self.agent._qt.sync_mirror)]

    def __init__(self, dim_info, is_continue, Qnet_lr, buffer_size, device, trick=None, mode=None):
        obs_dim, action_dim = dim_info
        self.device = _lib.require_device(device)
        self.agent = Agent(obs_dim, action_dim, Qnet_lr, self.device)
        self.buffer = Buffer(buffer_size, obs_dim, act_dim=action_dim if is_continue else 1, device=self.device)
        self.is_continue = is_continue
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.mode = _common.resolve_mode(mode)
        self._scratch = _common.DeviceScratch(self.device, self.agent._q.n_p)
        self._seed = _common.default_seed()
        self._n_learn = 0
        self.last_metrics = None

    def select_action(self, obs):
        """obs [obs_dim] -> numpy int scalar (reference); obs [N, obs_dim] -> int64 array [N] (vectorised envs)"""
        if self.is_continue:
            raise RuntimeError("DQN is not suitable for continuous action spaces (use dis_to_con)")   # DQN.py:78-80
        x, single = _common.as_obs_batch(obs, self.obs_dim)
        a = _common.infer(self.agent._q, x, _lib.INFER_ARGMAX, self.device, 1).reshape(-1).to(torch.int64).cpu().numpy()
        return a[0] if single else a

    def evaluate_action(self, obs):
        return self.select_action(obs)

    def add(self, obs, action, reward, next_obs, done):
        self.buffer.add(obs, action, reward, next_obs, done)

    def sample(self, batch_size):
        total_size = len(self.buffer)
        batch_size = min(total_size, batch_size)
        indices = np.random.choice(total_size, batch_size, replace=False)
        return self.buffer.sample(indices)

    def learn(self, batch_size, gamma, tau, *, n_updates=1, indices=None):
        """One (or ``n_updates`` sequential) DQN update(s).  ``indices`` ([n_updates, B] int64) overrides sampling."""
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
        a = _lib.DqnArgs()
        a.q, a.q_target, a.replay = ag._q.c_struct(), ag._qt.c_struct(), self.buffer.c_struct()
        a.indices, a.B, a.n_updates = idx.data_ptr(), B, n_updates
        a.gamma, a.tau = gamma, tau
        a.lr, a.beta1, a.beta2, a.eps = ag.lr, 0.9, 0.999, 1e-8
        a.step0 = ag.step
        out = self._scratch.out(n_updates, self.device)
        a.gpart, a.stats, a.out = self._scratch.gpart.data_ptr(), self._scratch.stats.data_ptr(), out.data_ptr()
        _lib.check(_lib.lib().frl_dqn_learn(ctypes.byref(a), _lib.stream_ptr(self.device)), "frl_dqn_learn")
        ag.step += n_updates
        self._n_learn += n_updates
        self.last_metrics = out[:n_updates]

    def update_target(self, tau):
        """Polyak update (already fused into learn(); provided for API parity, DQN.py:120-128)."""
        t, s = self.agent._qt, self.agent._q
        t.p.mul_(1.0 - tau).add_(s.p * tau)
        t.sync_mirror()

    def save(self, model_dir):
        torch.save({k: v.detach().clone().cpu() for k, v in self.agent.Qnet.state_dict().items()},
                   os.path.join(model_dir, "DQN.pt"))

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        device = device if device is not None else torch.device("cuda")
        policy = DQN(dim_info, is_continue, 0, 0, device=device, trick=trick)
        policy.agent.Qnet.load_state_dict(torch.load(os.path.join(model_dir, "DQN.pt"), map_location=device))
        return policy
