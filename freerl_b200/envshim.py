"""Synthetic fixed-shape stand-in for the subset of the gymnasium API the reference train scripts use
(SURVEY §8b last rows, §8d).  Injected as ``sys.modules['gymnasium']`` by ``freerl_b200.launcher`` only when the real
``gymnasium`` is not importable.  Dynamics: ``obs ~ N(0,1)`` fp32, ``reward ~ N(0,1)``, ``terminated ~ Bernoulli(p_term)``,
``truncated`` every ``t_max`` steps — the shapes (and only the shapes) of the named environments."""
import types

import numpy as np

_SPECS = {                   # name: (obs_dim, ('box', act_dim, high) | ('discrete', n), p_term, t_max, reward_threshold)
    "CartPole-v1": (4, ("discrete", 2), 1 / 200, 500, 475.0),
    "MountainCar-v0": (2, ("discrete", 3), 0.0, 200, -110.0),
    "LunarLander-v2": (8, ("discrete", 4), 1 / 300, 1000, 200.0),
    "LunarLander-v3": (8, ("discrete", 4), 1 / 300, 1000, 200.0),
    "Pendulum-v1": (3, ("box", 1, 2.0), 0.0, 200, None),
    "MountainCarContinuous-v0": (2, ("box", 1, 1.0), 0.0, 999, 90.0),
    "BipedalWalker-v3": (24, ("box", 4, 1.0), 1 / 500, 1600, 300.0),
    "HalfCheetah-v4": (17, ("box", 6, 1.0), 0.0, 1000, 4800.0),
}


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype)
        self.dtype = dtype
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(self.n))


class SyntheticEnv:
    def __init__(self, name):
        obs_dim, act, self.p_term, self.t_max, _ = _SPECS[name]
        self.observation_space = Box(-np.inf, np.inf, (obs_dim,))
        self.action_space = Box(-act[2], act[2], (act[1],)) if act[0] == "box" else Discrete(act[1])
        self._rng = np.random.default_rng(0)
        self._t = 0

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        self._t = 0
        return self._rng.standard_normal(self.observation_space.shape).astype(np.float32), {}

    def step(self, action):
        self._t += 1
        obs = self._rng.standard_normal(self.observation_space.shape).astype(np.float32)
        terminated = bool(self._rng.random() < self.p_term)
        truncated = self._t >= self.t_max
        return obs, float(self._rng.standard_normal()), terminated, truncated, {}

    def close(self):
        pass


def make(name, **kwargs):
    if name not in _SPECS:
        raise ValueError("envshim knows the shapes of %s only" % sorted(_SPECS))
    return SyntheticEnv(name)


def spec(name):
    return types.SimpleNamespace(id=name, reward_threshold=_SPECS[name][4])


spaces = types.SimpleNamespace(Box=Box, Discrete=Discrete)


def as_module():
    m = types.ModuleType("gymnasium")
    m.make, m.spec, m.spaces = make, spec, spaces
    m.__freerl_b200_shim__ = True
    return m
