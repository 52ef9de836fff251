"""Per-step work of the reference train loops, for N vectorised envs on the device.

Same class names and call signatures as the reference's helpers, with a leading env axis:

* ``Normalization(shape)`` / ``RunningMeanStd``  — ``PPO_file/normalization.py:17-49`` (identical copies in ``MAPPO_file``,
  ``DDPG_file/DDPG.py:358-388``, ``SAC_file/SAC.py:357-388``): ``norm(x[N, shape], update=True)``.
* ``RewardScaling(shape=1, gamma)``             — ``PPO_file/normalization.py:87-101``: ``rs(reward[N])``, ``rs.reset(done_mask)``.
* ``OUNoise(action_dim, ...)``                  — ``SAC_file/SAC.py:334-355``: ``ou.noise()`` → ``[N, action_dim]``, ``ou.reset(done_mask)``.
* ``OUNoise.explore`` / ``explore_gauss``       — the action post-processing of ``DDPG_file/DDPG.py:519-522``.
* ``epsilon_greedy`` / ``dis_to_con``           — ``DQN_file/DQN.py:307-310`` and ``:195-217``.

The reference owns ONE statistics object and calls it once per env step.  Here the rows of a vector step are folded in env
order inside one kernel launch (``frl_vecnorm`` / ``frl_reward_scaling``), so statistics and outputs equal — bit for bit —
what the reference object produces when it is called for row 0, 1, … N-1 in turn (``tests/test_vecloop.py`` checks that
against the reference-generated fixture).  Inputs may be numpy (copied to the device) or device tensors; float32 and
float64 rows follow NumPy-2's dtype rules of the reference code (float32 rows keep a float32 mean).  Outputs are fresh
device tensors: fp32 by default (what the replay / rollout buffers store), float64 with ``out_dtype=torch.float64``.

CUDA only: every call goes through the C ABI (``include/freerl_b200.h``); there is no host fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev_rows(x, device, width):
    """-> contiguous [N, width] float32/float64 tensor on `device` (numpy float64/float32 keep their dtype, anything else -> float64
    like np.array of python scalars)."""
    if isinstance(x, torch.Tensor):
        t = x
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
    else:
        a = np.asarray(x)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    t = t.to(device).reshape(-1, width).contiguous()
    return t


class RunningMeanStd:
    """Attribute surface of the reference class: ``n``, ``mean``, ``S``, ``std`` (device float64 views of one [3, shape] tensor)."""

    def __init__(self, shape, device):
        self.n = 0
        self.shape = int(shape)
        self.state = torch.zeros((3, self.shape), dtype=torch.float64, device=device)

    @property
    def mean(self):
        return self.state[0]

    @property
    def S(self):
        return self.state[1]

    @property
    def std(self):
        return self.state[2]


class Normalization:
    def __init__(self, shape, device="cuda"):
        self.device = _lib.require_device(device)
        self.running_ms = RunningMeanStd(shape, self.device)

    def __call__(self, x, update=True, out_dtype=torch.float32):
        ms = self.running_ms
        rows = _dev_rows(x, self.device, ms.shape)
        n = rows.shape[0]
        out = torch.empty((n, ms.shape), dtype=out_dtype, device=self.device)
        o32, o64 = (_lib.ptr(out), None) if out_dtype == torch.float32 else (None, _lib.ptr(out))
        _lib.check(_lib.lib().frl_vecnorm(_lib.ptr(ms.state), ms.n, _lib.ptr(rows), int(rows.dtype == torch.float64), n, ms.shape,
                                          int(bool(update)), o32, o64, _lib.stream_ptr(self.device)), "frl_vecnorm")
        if update:
            ms.n += n
        return out


class RewardScaling:
    def __init__(self, shape, gamma, n_envs=1, device="cuda"):
        assert int(shape) == 1, "the reference scales a scalar reward (shape=1)"
        self.shape, self.gamma, self.n_envs = 1, float(gamma), int(n_envs)
        self.device = _lib.require_device(device)
        self.running_ms = RunningMeanStd(1, self.device)
        self.R = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)

    def __call__(self, x, out_dtype=torch.float32):
        ms = self.running_ms
        rows = _dev_rows(x, self.device, 1)
        assert rows.shape[0] == self.n_envs, "one reward per env"
        out = torch.empty(self.n_envs, dtype=out_dtype, device=self.device)
        o32, o64 = (_lib.ptr(out), None) if out_dtype == torch.float32 else (None, _lib.ptr(out))
        _lib.check(_lib.lib().frl_reward_scaling(_lib.ptr(ms.state), ms.n, _lib.ptr(self.R), _lib.ptr(rows), int(rows.dtype == torch.float64),
                                                 self.gamma, self.n_envs, o32, o64, _lib.stream_ptr(self.device)), "frl_reward_scaling")
        ms.n += self.n_envs
        return out

    def reset(self, done_mask=None):
        """``reset()`` zeroes every env's discounted return; ``reset(done_mask[N])`` only the envs whose episode ended."""
        _masked_reset(self.R.reshape(self.n_envs, 1), done_mask, 0.0, self.device)


def _masked_reset(state, mask, value, device):
    if mask is None:
        state.fill_(value)
        return
    m = torch.as_tensor(np.asarray(mask, dtype=np.uint8) if not isinstance(mask, torch.Tensor) else mask.to(torch.uint8)).to(device).contiguous()
    _lib.check(_lib.lib().frl_masked_reset(_lib.ptr(state), _lib.ptr(m), state.shape[0], state.shape[1], float(value),
                                           _lib.stream_ptr(device)), "frl_masked_reset")


class OUNoise:
    """``OUNoise(action_dim, mu=0, theta=0.15, sigma=0.1, dt=1e-2, scale=None)`` per env.  ``mode='parity'`` draws the normals with
    ``np.random.randn(n_envs, action_dim)`` — the same legacy-stream values the reference consumes when its ``noise()`` is called
    once per env in env order; ``mode='fast'`` uses the device Philox stream."""

    def __init__(self, action_dim, mu=0, theta=0.15, sigma=0.1, dt=1e-2, scale=None, n_envs=1, device="cuda", mode="parity", seed=0):
        self.action_dim, self.mu, self.theta, self.sigma, self.dt, self.scale = int(action_dim), mu, theta, sigma, dt, scale
        self.n_envs, self.mode, self.seed, self._counter = int(n_envs), mode, int(seed), 0
        self.device = _lib.require_device(device)
        self.state = torch.full((self.n_envs, self.action_dim), float(mu), dtype=torch.float64, device=self.device)

    def reset(self, done_mask=None):
        _masked_reset(self.state, done_mask, float(self.mu), self.device)

    def _args(self, kind, action, max_action, gauss_scale=0.0, gauss_sigma=0.0, z=None):
        a = _lib.ExploreArgs()
        a.kind, a.N, a.A = kind, self.n_envs, self.action_dim
        if z is None and self.mode == "parity":
            z = np.random.randn(self.n_envs, self.action_dim)
        zt = None if z is None else torch.as_tensor(np.asarray(z, dtype=np.float64)).to(self.device).contiguous()
        self._counter += 1
        a.action, a.ou_state, a.z = _lib.ptr(action), _lib.ptr(self.state), (None if zt is None else _lib.ptr(zt))
        a.seed, a.counter = self.seed, self._counter
        a.mu, a.theta, a.sigma, a.dt = float(self.mu), float(self.theta), float(self.sigma), float(self.dt)
        a.scale = -1.0 if self.scale is None else float(self.scale)
        a.gauss_scale, a.gauss_sigma, a.max_action = float(gauss_scale), float(gauss_sigma), float(max_action)
        return a, zt

    def noise(self, z=None):
        """Advance every env's OU state; returns ``state * scale`` as float64 ``[n_envs, action_dim]`` (``SAC.py:347-355``)."""
        zero = torch.zeros((self.n_envs, self.action_dim), dtype=torch.float32, device=self.device)
        out = torch.empty((self.n_envs, self.action_dim), dtype=torch.float64, device=self.device)
        a, keep = self._args(0, zero, 1.0, z=z)
        a.clip, a.out64 = 0, _lib.ptr(out)
        _lib.check(_lib.lib().frl_explore(C.byref(a), _lib.stream_ptr(self.device)), "frl_explore")
        return out

    def explore(self, action, max_action, out_dtype=torch.float64, z=None):
        """``np.clip(action * max_action + ou_noise.noise() * max_action, -max_action, max_action)`` (``DDPG.py:520``)."""
        act = _dev_rows(action, self.device, self.action_dim).to(torch.float32)
        out = torch.empty((self.n_envs, self.action_dim), dtype=out_dtype, device=self.device)
        a, keep = self._args(0, act, max_action, z=z)
        a.clip = 1
        if out_dtype == torch.float64:
            a.out64 = _lib.ptr(out)
        else:
            a.out = _lib.ptr(out)
        _lib.check(_lib.lib().frl_explore(C.byref(a), _lib.stream_ptr(self.device)), "frl_explore")
        return out


def explore_gauss(action, max_action, gauss_scale, gauss_sigma, device="cuda", z=None, mode="parity", seed=0, counter=0,
                  out_dtype=torch.float64):
    """``np.clip(action * max_action + gauss_scale * np.random.normal(scale=gauss_sigma * max_action, size=action_dim), -max_action,
    max_action)`` per env row (``DDPG_file/DDPG.py:522``, ``TD3_file/TD3.py`` main loop)."""
    device = _lib.require_device(device)
    act = (action if isinstance(action, torch.Tensor) else torch.as_tensor(np.asarray(action))).to(device).to(torch.float32)
    act = act.reshape(-1, act.shape[-1]).contiguous()
    n, adim = act.shape
    if z is None and mode == "parity":
        z = np.random.randn(n, adim)
    zt = None if z is None else torch.as_tensor(np.asarray(z, dtype=np.float64)).to(device).contiguous()
    out = torch.empty((n, adim), dtype=out_dtype, device=device)
    a = _lib.ExploreArgs()
    a.kind, a.N, a.A = 1, n, adim
    a.action, a.ou_state, a.z = _lib.ptr(act), None, (None if zt is None else _lib.ptr(zt))
    a.seed, a.counter = int(seed), int(counter)
    a.gauss_scale, a.gauss_sigma, a.max_action, a.clip = float(gauss_scale), float(gauss_sigma), float(max_action), 1
    if out_dtype == torch.float64:
        a.out64 = _lib.ptr(out)
    else:
        a.out = _lib.ptr(out)
    _lib.check(_lib.lib().frl_explore(C.byref(a), _lib.stream_ptr(device)), "frl_explore")
    return out


def epsilon_greedy(greedy, n_actions, epsilon, device="cuda", mode="parity", seed=0, counter=0):
    """``if np.random.rand() < epsilon: action = np.random.randint(action_dim) else: action = policy.select_action(obs)``
    (``DQN_file/DQN.py:307-310``) for N envs.  ``greedy``: the N greedy actions (int64).  Parity mode draws ``rand()`` — and a
    ``randint`` only where it fires — on the host in env order, i.e. the legacy-stream consumption of N reference iterations;
    fast mode draws on the device."""
    device = _lib.require_device(device)
    g = (greedy if isinstance(greedy, torch.Tensor) else torch.as_tensor(np.asarray(greedy, dtype=np.int64))).to(device).to(torch.int64).reshape(-1).contiguous()
    n = g.numel()
    out = torch.empty(n, dtype=torch.int64, device=device)
    u = r = None
    if mode == "parity":
        uh, rh = np.empty(n), np.zeros(n, dtype=np.int64)
        for i in range(n):
            uh[i] = np.random.rand()
            if uh[i] < epsilon:
                rh[i] = np.random.randint(n_actions)
        u, r = torch.from_numpy(uh).to(device), torch.from_numpy(rh).to(device)
    _lib.check(_lib.lib().frl_epsilon_greedy(_lib.ptr(g), n, int(n_actions), float(epsilon), None if u is None else _lib.ptr(u),
                                             None if r is None else _lib.ptr(r), C.c_uint64(int(seed)), C.c_uint64(int(counter)),
                                             _lib.ptr(out), _lib.stream_ptr(device)), "frl_epsilon_greedy")
    return out


def dis_to_con(discrete_action, low, high, action_dim, device="cuda", out_dtype=torch.float64):
    """``dis_to_con(discrete_action, env, action_dim)`` (``DQN_file/DQN.py:195-217``) for N actions: ``low`` / ``high`` are the Box
    bounds (``env.action_space.low / high``, float32), result ``[N, len(low)]``."""
    device = _lib.require_device(device)
    a = (discrete_action if isinstance(discrete_action, torch.Tensor) else torch.as_tensor(np.asarray(discrete_action, dtype=np.int64)))
    a = a.to(device).to(torch.int64).reshape(-1).contiguous()
    lo = torch.as_tensor(np.asarray(low, dtype=np.float32).reshape(-1)).to(device)
    hi = torch.as_tensor(np.asarray(high, dtype=np.float32).reshape(-1)).to(device)
    shape = lo.numel()
    per = int(action_dim ** (1 / shape)) if shape > 1 else 0
    out = torch.empty((a.numel(), shape), dtype=out_dtype, device=device)
    o64, o32 = (_lib.ptr(out), None) if out_dtype == torch.float64 else (None, _lib.ptr(out))
    _lib.check(_lib.lib().frl_dis_to_con(_lib.ptr(a), a.numel(), int(action_dim), shape, per, _lib.ptr(lo), _lib.ptr(hi), o64, o32,
                                         _lib.stream_ptr(device)), "frl_dis_to_con")
    return out
