"""Prioritized / n-step replay with the reference's ``DQN_file/Buffer.py`` interface, device resident.

``SumTree`` (``Buffer.py:134-194``), ``PER_Buffer`` (``:66-132``), ``N_Step_Buffer`` (``:199-293``) and
``N_Step_PER_Buffer`` (``:333-399``).  The float64 array heap lives in HBM with the reference's exact layout
(``2*cap-1`` nodes, leaf i at ``i+cap-1``, no power-of-two padding) and is updated by ``frl_sumtree_update`` in
batch order, so leaf priorities AND every internal node stay bit-identical to the reference.  Sampled indices are
bit-exact in parity mode: the stratified targets are ``a + (b-a)*u`` with ``u`` taken from numpy's legacy global
stream exactly like ``np.random.uniform(a, b)``.  N-step folding is float64 host arithmetic per env, as in the
reference (python scalars), applied before the transition is stored.
"""
import ctypes
from collections import deque

import numpy as np
import torch

from . import _common, _lib
from .Buffer import Buffer


class SumTree:
    def __init__(self, capacity, device):
        self.capacity = capacity = int(capacity)
        self.device = _lib.require_device(device)
        self.tree = torch.zeros(max(2 * capacity - 1, 1), dtype=torch.float64, device=self.device)
        self._scratch = torch.zeros(4096, dtype=torch.float64, device=self.device)
        self._max = torch.zeros(1, dtype=torch.float64, device=self.device)

    def _update(self, idx, pri32=None, pri64=None, pri_const=0.0, idx0=0, is_range=False, B=None):
        B = int(B if B is not None else idx.numel())
        _lib.check(_lib.lib().frl_sumtree_update(
            _lib.ptr(self.tree), self.capacity, _lib.ptr(idx), _lib.ptr(pri32), _lib.ptr(pri64), float(pri_const), int(idx0),
            int(is_range), B, _lib.ptr(self._scratch), _lib.stream_ptr(self.device)), "frl_sumtree_update")

    def add(self, buffer_index, priority):
        """set one leaf (reference ``SumTree.add``)"""
        idx = torch.tensor([int(buffer_index)], dtype=torch.int64, device=self.device)
        self._update(idx, pri_const=float(np.asarray(priority).reshape(-1)[0]))

    def update(self, idx, priority):
        """reference ``SumTree.update`` (``DQN_file/Buffer.py:157-166``): ``idx`` is a TREE index of a leaf (``buffer_index + capacity - 1``)"""
        idx = int(idx)
        if not self.capacity - 1 <= idx < 2 * self.capacity - 1:
            raise IndexError("SumTree.update: tree index %d is not a leaf (leaves are [%d, %d))" % (idx, self.capacity - 1, 2 * self.capacity - 1))
        self.add(idx - self.capacity + 1, priority)

    def get(self, s):
        """reference ``SumTree.get`` for one value (host round trip; the batched path is ``PER_Buffer.sample``)"""
        tree = self.tree.cpu().numpy()
        node, n = 0, tree.shape[0]
        while 2 * node + 1 < n:
            left = 2 * node + 1
            if s <= tree[left]:
                node = left
            else:
                s = s - tree[left]
                node = left + 1
        return tree[node], node - self.capacity + 1

    def sum(self):
        return float(self.tree[0].item())

    def max_device(self):
        """fp64 device scalar = np.max(tree[-capacity:])"""
        _lib.check(_lib.lib().frl_sumtree_max(_lib.ptr(self.tree), self.capacity, _lib.ptr(self._scratch), 4096, _lib.ptr(self._max),
                                              _lib.stream_ptr(self.device)), "frl_sumtree_max")
        return self._max

    def max(self):
        return float(self.max_device().item())


class PER_Buffer:
    def __init__(self, capacity, obs_dim, act_dim, device, alpha=0.5, beta=0.4, beta_increment=0.001, epsilon=0.01,
                 prob_floor=1e-7, mode=None):
        self.capacity = int(capacity)
        self.alpha, self.beta, self.beta_increment, self.epsilon = alpha, beta, beta_increment, epsilon
        self.prob_floor = prob_floor
        self.device = _lib.require_device(device)
        self.sumtree = SumTree(self.capacity, self.device)
        self.buffer = Buffer(self.capacity, obs_dim, act_dim, self.device)
        self.mode = _common.resolve_mode(mode)
        self._seed = _common.default_seed()
        self._n_sample = 0

    # ---- add: new transitions get the current max priority (1.0 for the very first) ------------------------------
    def _add_priorities(self, n):
        tree = self.sumtree
        empty = len(self.buffer) == 0
        pmax = None if empty else tree.max_device()          # max priority BEFORE the batch (new leaves do not raise it)
        for s in range(0, n, 1024):                           # the ordered update kernel takes <= 1024 leaves per launch
            m = min(1024, n - s)
            i0 = (self.buffer._index + s) % self.capacity
            if empty:
                tree._update(None, pri_const=1.0, idx0=i0, is_range=True, B=m)
            else:
                tree._update(None, pri64=pmax, idx0=i0, is_range=True, B=m)

    def add(self, obs, action, reward, next_obs, done):
        n = int(np.asarray(obs).reshape(-1, self.buffer.obs_dim).shape[0])
        self._add_priorities(n)
        self.buffer.add(obs, action, reward, next_obs, done)

    # ---- sample ----------------------------------------------------------------------------------------------
    def sample_device(self, batch_size, u=None):
        """(indices int64 [B], is_weight fp32 [B], priorities fp32 [B]) as device tensors; no host sync."""
        self.beta = np.min([1., self.beta + self.beta_increment])
        if u is None and self.mode == "parity":
            u = np.array([np.random.random_sample() for _ in range(batch_size)])     # one draw per np.random.uniform(a, b)
        ud = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(self.device) if u is not None else None
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        pri = torch.empty(batch_size, dtype=torch.float32, device=self.device)
        w = torch.empty(batch_size, dtype=torch.float32, device=self.device)
        self._n_sample += 1
        _lib.check(_lib.lib().frl_sumtree_sample(
            _lib.ptr(self.sumtree.tree), self.capacity, _lib.ptr(ud), ctypes.c_uint64(self._seed), ctypes.c_uint64(self._n_sample),
            batch_size, len(self.buffer), float(self.beta), float(self.prob_floor), _lib.ptr(idx), _lib.ptr(pri), _lib.ptr(w),
            _lib.stream_ptr(self.device)), "frl_sumtree_sample")
        return idx, w, pri

    def sample(self, batch_size):
        """reference signature: ``-> (batch_indices np.int64 [B], is_weight tensor [B] on device)``"""
        idx, w, _ = self.sample_device(batch_size)
        return idx.cpu().numpy(), w

    # ---- priorities ------------------------------------------------------------------------------------------
    def update_priorities(self, indices, td_error):
        idx = self.buffer._indices_to_device(indices).reshape(-1)
        if isinstance(td_error, torch.Tensor):
            td = td_error.detach().to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        else:
            td = torch.from_numpy(np.ascontiguousarray(td_error, dtype=np.float32).reshape(-1)).to(self.device)
        n = min(idx.numel(), td.numel())             # `for idx, priority in zip(indices, priorities)` (DQN_file/Buffer.py:128): the shorter wins
        for s in range(0, n, 1024):                  # one launch per <= 1024 ordered updates: priority transform + heap update
            e = min(s + 1024, n)
            _lib.check(_lib.lib().frl_sumtree_update_td(
                _lib.ptr(self.sumtree.tree), self.capacity, _lib.ptr(idx[s:e].contiguous()), _lib.ptr(td[s:e].contiguous()),
                float(self.epsilon), float(self.alpha), e - s, _lib.ptr(self.sumtree._scratch), _lib.stream_ptr(self.device)),
                "frl_sumtree_update_td")

    def __len__(self):
        return len(self.buffer)


def _fold(window, gamma):
    """``_get_n_step_info`` (``DQN_file/Buffer.py:261-269`` / ``:350-358``): python float64 arithmetic."""
    obs, action = window[0][0], window[0][1]
    reward, next_obs, done = window[-1][2], window[-1][3], window[-1][4]
    for i in range(len(window) - 2, -1, -1):
        _, _, r, n_o, d = window[i]
        reward = r + gamma * reward * (1 - d)
        if d:
            next_obs, done = n_o, d
    return obs, action, reward, next_obs, done


class _NStepMixin:
    """Sliding n-step window(s).  One deque per env column; like the reference the window is never reset at episode
    end and once it is full EVERY add emits one folded transition."""

    def _init_nstep(self, gamma, n_step):
        self.gamma, self.n_step = gamma, n_step
        self.n_step_gamma = gamma ** n_step
        self.n_step_deque = deque(maxlen=n_step)          # env 0 (reference attribute name)
        self._deques = [self.n_step_deque]

    def _push(self, obs, action, reward, next_obs, done, obs_dim):
        o = np.asarray(obs).reshape(-1, obs_dim)
        n = o.shape[0]
        if n > 1 and len(self._deques) == 1 and len(self.n_step_deque) == 0:
            return self._push_vec(o, action, reward, next_obs, done, obs_dim)
        while len(self._deques) < n:
            self._deques.append(deque(maxlen=self.n_step))
        a = np.asarray(action).reshape(n, -1)
        r = np.asarray(reward, dtype=np.float64).reshape(n)
        o2 = np.asarray(next_obs).reshape(n, obs_dim)
        d = np.asarray(done).reshape(n)
        out = []
        for e in range(n):
            dq = self._deques[e]
            dq.append((o[e], a[e], float(r[e]), o2[e], bool(d[e])))
            if len(dq) == self.n_step:
                out.append(_fold(dq, self.gamma))
        if not out:
            return None
        return (np.stack([x[0] for x in out]), np.stack([x[1] for x in out]), np.array([x[2] for x in out], np.float64),
                np.stack([x[3] for x in out]), np.array([x[4] for x in out]))


    def _push_vec(self, o, action, reward, next_obs, done, obs_dim):
        """N vectorised envs stepping in lock-step: one window of the last n_step BATCHES, folded for all envs at once in
        numpy float64 with the reference's expression order (r + gamma * R * (1 - d): identical bits to N python folds)."""
        n = o.shape[0]
        win = self.__dict__.setdefault("_vec_window", deque(maxlen=self.n_step))
        if win and win[0][0].shape[0] != n:
            raise ValueError("vectorised n-step adds must keep the same number of envs")
        win.append((o, np.asarray(action).reshape(n, -1), np.asarray(reward, dtype=np.float64).reshape(n),
                    np.asarray(next_obs).reshape(n, obs_dim), np.asarray(done).reshape(n).astype(bool)))
        if len(win) < self.n_step:
            return None
        R, nobs, dn = win[-1][2].copy(), win[-1][3].copy(), win[-1][4].copy()
        for i in range(self.n_step - 2, -1, -1):
            _, _, r, n_o, d = win[i]
            R = r + self.gamma * R * (1 - d)
            nobs[d] = n_o[d]
            dn = dn | d
        return win[0][0], win[0][1], R, nobs, dn


class N_Step_Buffer(Buffer, _NStepMixin):
    def __init__(self, capacity, obs_dim, act_dim, device, gamma, n_step=2):
        Buffer.__init__(self, capacity, obs_dim, act_dim, device)
        self._init_nstep(gamma, n_step)

    def add(self, obs, action, reward, next_obs, done):
        folded = self._push(obs, action, reward, next_obs, done, self.obs_dim)
        if folded is not None:
            Buffer.add(self, *folded)


class N_Step_PER_Buffer(PER_Buffer, _NStepMixin):
    def __init__(self, capacity, obs_dim, act_dim, device, alpha=0.5, beta=0.4, beta_increment=0.001, epsilon=0.01, gamma=None,
                 n_step=3, mode=None):
        PER_Buffer.__init__(self, capacity, obs_dim, act_dim, device, alpha, beta, beta_increment, epsilon, mode=mode)
        self._init_nstep(gamma, n_step)

    def add(self, obs, action, reward, next_obs, done):
        folded = self._push(obs, action, reward, next_obs, done, self.buffer.obs_dim)
        if folded is not None:
            PER_Buffer.add(self, *folded)
