"""Device-resident replay buffers with the reference's ``Buffer.py`` interface.

Drop-in for ``from Buffer import Buffer`` (``SAC_file/Buffer.py:11-61`` = DQN/TD3/DDPG/MADDPG copies): same
constructor ``Buffer(capacity, obs_dim, act_dim, device)``, ``add(obs, action, reward, next_obs, done)``,
``sample(indices) -> (obs, actions, rewards[B,1], next_obs, dones[B,1])`` fresh fp32 tensors on ``device``,
``len()``, ``_index`` / ``_size``.  Storage is ONE fp32 row per transition in HBM,
``[obs | action | reward | done | next_obs | pad]`` (16-byte multiple), so a sampled transition is a single
contiguous 16-B-vectorised read.  The float64->float32 cast the reference applies in ``sample`` happens at
``add`` (same values: rounding a float64 once to fp32 is order-independent).

Extension: ``add`` also accepts a batch (``obs`` of shape ``[N, obs_dim]`` …) for N vectorised envs; rows are
written in env order exactly as N sequential single ``add`` calls would.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .nets import pad4


class Buffer:
    """replay buffer for each agent (device resident)"""

    def __init__(self, capacity, obs_dim, act_dim, device):
        self.capacity = capacity = int(capacity)
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.device = _lib.require_device(device)
        self.row_floats = pad4(2 * self.obs_dim + self.act_dim + 2)
        self.storage = torch.zeros((max(capacity, 1), self.row_floats), dtype=torch.float32, device=self.device)
        self._index = 0
        self._size = 0
        self._c = _lib.Replay(self.storage.data_ptr(), max(capacity, 1), self.row_floats, self.obs_dim, self.act_dim)

    # ---- C descriptor ------------------------------------------------------------------------------
    def c_struct(self):
        return self._c

    # ---- the reference's five arrays as views of the row store (SAC_file/Buffer.py:18-22) --------------------------
    # device fp32 views, not host float64 arrays: element [i] is what ``sample([i])`` returns; writing through them edits the store
    @property
    def obs(self):
        return self.storage[:, :self.obs_dim]

    @property
    def actions(self):
        return self.storage[:, self.obs_dim:self.obs_dim + self.act_dim]

    @property
    def rewards(self):
        return self.storage[:, self.obs_dim + self.act_dim]

    @property
    def dones(self):
        return self.storage[:, self.obs_dim + self.act_dim + 1]

    @property
    def next_obs(self):
        o = self.obs_dim + self.act_dim + 2
        return self.storage[:, o:o + self.obs_dim]

    # ---- add ---------------------------------------------------------------------------------------
    def _pack(self, obs, action, reward, next_obs, done):
        od, ad = self.obs_dim, self.act_dim
        obs = np.asarray(obs, dtype=np.float32).reshape(-1, od)
        n = obs.shape[0]
        rows = np.zeros((n, self.row_floats), np.float32)
        rows[:, :od] = obs
        rows[:, od:od + ad] = np.asarray(action, dtype=np.float32).reshape(n, ad)
        rows[:, od + ad] = np.asarray(reward, dtype=np.float32).reshape(n)
        rows[:, od + ad + 1] = np.asarray(done, dtype=np.float32).reshape(n)
        rows[:, od + ad + 2:2 * od + ad + 2] = np.asarray(next_obs, dtype=np.float32).reshape(n, od)
        return rows

    def add(self, obs, action, reward, next_obs, done):
        """add one experience (reference semantics) or a batch of N experiences (vectorised envs)"""
        rows = torch.from_numpy(self._pack(obs, action, reward, next_obs, done)).to(self.device)
        self.add_rows(rows)

    def add_rows(self, rows):
        """rows: device tensor [n, row_floats] already in storage layout"""
        n = rows.shape[0]
        if n > self.capacity:
            rows = rows[n - self.capacity:]       # only the last `capacity` rows survive n sequential adds
            self._index = (self._index + n - self.capacity) % self.capacity
            n = self.capacity
        first = min(n, self.capacity - self._index)
        self.storage[self._index:self._index + first].copy_(rows[:first])
        if first < n:
            self.storage[:n - first].copy_(rows[first:])
        self._index = (self._index + n) % self.capacity
        self._size = min(self._size + n, self.capacity)

    def add_device(self, obs, action, reward, next_obs, done):
        """Batched add of fields already resident on the device (fp32 tensors) — packs SoA -> AoS rows with the
        ``frl_replay_add_batch`` kernel.  ``done`` is a float tensor of 0/1."""
        n = obs.shape[0]
        assert n <= self.capacity
        _lib.check(_lib.lib().frl_replay_add_batch(
            ctypes.byref(self._c), self._index, _lib.ptr(obs), _lib.ptr(action), _lib.ptr(reward), _lib.ptr(next_obs),
            _lib.ptr(done), n, _lib.stream_ptr(self.device)), "frl_replay_add_batch")
        self._index = (self._index + n) % self.capacity
        self._size = min(self._size + n, self.capacity)

    # ---- sample ------------------------------------------------------------------------------------
    def _indices_to_device(self, indices):
        """Host (numpy / list) indices get numpy's fancy-indexing semantics before they reach the device: negative values count
        from the end of the store, anything outside ``[-capacity, capacity)`` raises ``IndexError`` (``SAC_file/Buffer.py:41-45``
        indexes ``self.obs[indices]``).  Device tensors are taken as they are — the caller owns their range."""
        if isinstance(indices, torch.Tensor):
            return indices.to(device=self.device, dtype=torch.int64).contiguous()
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        if idx.size:
            lo, hi = int(idx.min()), int(idx.max())
            if lo < -self.capacity or hi >= self.capacity:
                raise IndexError("index %d is out of bounds for axis 0 with size %d" % (lo if lo < -self.capacity else hi, self.capacity))
            if lo < 0:
                idx = np.where(idx < 0, idx + self.capacity, idx)
        return torch.from_numpy(idx).to(self.device)

    def sample(self, indices):
        idx = self._indices_to_device(indices)
        B = idx.numel()
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
        obs, actions, rewards = f(B, self.obs_dim), f(B, self.act_dim), f(B, 1)
        next_obs, dones = f(B, self.obs_dim), f(B, 1)
        _lib.check(_lib.lib().frl_replay_gather(
            ctypes.byref(self._c), _lib.ptr(idx), B, _lib.ptr(obs), _lib.ptr(actions), _lib.ptr(rewards),
            _lib.ptr(next_obs), _lib.ptr(dones), _lib.stream_ptr(self.device)), "frl_replay_gather")
        return obs, actions, rewards, next_obs, dones

    def __len__(self):
        return self._size


class Buffer_for_PPO:
    """Device-resident rollout store with the reference interface (``PPO_file/Buffer.py:266-323`` =
    ``MAPPO_file/Buffer.py:266-323``): ``Buffer_for_PPO(capacity, obs_dim, act_dim, device, trick=None)``,
    ``add(obs, action, reward, next_obs, done, action_log_probs, adv_done)``, ``all()`` -> 7 fp32 tensors covering
    the FULL capacity, ``clear()``, ``len()``.

    Vectorised extension: ``add`` of N rows at once stores one time step of N envs; the rollout is then the
    ``[T, N]`` time-major array the GAE kernel scans per env column (N = 1 reproduces the reference's flat order).
    """

    def __init__(self, capacity, obs_dim, act_dim, device, trick=None):
        self.capacity = capacity = int(capacity)
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.device = _lib.require_device(device)
        self.logp_dim = 1 if (trick is not None and trick.get('decaystd')) else self.act_dim
        c = max(capacity, 1)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.obs, self.actions, self.rewards = z(c, self.obs_dim), z(c, self.act_dim), z(c)
        self.next_obs, self.dones = z(c, self.obs_dim), z(c)
        self.action_log_probs, self.adv_dones = z(c, self.logp_dim), z(c)
        self._index = 0
        self._size = 0
        self.n_envs = 1

    def add(self, obs, action, reward, next_obs, done, action_log_probs, adv_done):
        od, ad = self.obs_dim, self.act_dim
        o = np.asarray(obs, dtype=np.float32).reshape(-1, od)
        n = o.shape[0]
        if self._size == 0:
            self.n_envs = n
        packed = np.concatenate([
            o, np.asarray(action, dtype=np.float32).reshape(n, ad), np.asarray(reward, dtype=np.float32).reshape(n, 1),
            np.asarray(next_obs, dtype=np.float32).reshape(n, od), np.asarray(done, dtype=np.float32).reshape(n, 1),
            np.asarray(action_log_probs, dtype=np.float32).reshape(n, self.logp_dim),
            np.asarray(adv_done, dtype=np.float32).reshape(n, 1)], axis=1)
        dev = torch.from_numpy(packed).to(self.device)
        i = self._index
        if i + n > self.capacity:
            raise ValueError("Buffer_for_PPO: a vectorised add must not wrap (capacity %d, index %d, n %d)" % (self.capacity, i, n))
        c0 = 0
        for dst, w in ((self.obs, od), (self.actions, ad), (self.rewards, 1), (self.next_obs, od), (self.dones, 1),
                       (self.action_log_probs, self.logp_dim), (self.adv_dones, 1)):
            dst[i:i + n].copy_(dev[:, c0:c0 + w].reshape(dst[i:i + n].shape))
            c0 += w
        self._index = (i + n) % self.capacity
        self._size = min(self._size + n, self.capacity)

    def __len__(self):
        return self._size

    def clear(self):
        self._index = 0
        self._size = 0

    def all(self):
        """Views of the device arrays, shaped like the reference's return values (rewards / dones / adv_dones [cap,1])."""
        return (self.obs, self.actions, self.rewards.reshape(-1, 1), self.next_obs, self.dones.reshape(-1, 1),
                self.action_log_probs, self.adv_dones.reshape(-1, 1))
