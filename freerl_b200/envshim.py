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


# ---- PettingZoo MPE parallel-env subset (MADDPG.py:255-282, 479-548; MAPPO.py:514-523, 662-689) ----------------------------
_MPE = {        # env module name: (agent-name pattern, default N, obs_dim(N), continuous action dim)
    "simple_spread_v3": (lambda n: ["agent_%d" % i for i in range(n)], 3, lambda n: 6 * n, 5),
    "simple_adversary_v3": (lambda n: ["adversary_0"] + ["agent_%d" % i for i in range(n)], 2, None, 5),
    "simple_tag_v3": (lambda n: ["adversary_%d" % i for i in range(3)] + ["agent_%d" % i for i in range(n)], 1, None, 5),
}


class SyntheticParallelEnv:
    """Shape-only stand-in for ``pettingzoo.mpe.<name>.parallel_env``: dict-in / dict-out ``reset`` / ``step``, per-agent
    ``observation_space(agent)`` / ``action_space(agent)`` (Box [0, 1]^5 when ``continuous_actions`` else Discrete(5)),
    ``agents`` emptied when the episode is truncated at ``max_cycles`` (PettingZoo semantics)."""

    def __init__(self, name, max_cycles=25, continuous_actions=False, N=None, num_good=None, **_):
        names, n_def, obs_of, self._act = _MPE[name]
        n = N if N is not None else (num_good if num_good is not None else n_def)
        self.possible_agents = names(n)
        if name == "simple_spread_v3":
            dims = {a: obs_of(n) for a in self.possible_agents}
        elif name == "simple_adversary_v3":
            dims = {a: (4 * n if a.startswith("adversary") else 4 * n + 2) for a in self.possible_agents}
        else:
            dims = {a: (12 + 2 * n if a.startswith("adversary") else 10 + 2 * n) for a in self.possible_agents}
        self._obs_space = {a: Box(-np.inf, np.inf, (d,)) for a, d in dims.items()}
        self._act_space = {a: (Box(0.0, 1.0, (self._act,)) if continuous_actions else Discrete(self._act)) for a in self.possible_agents}
        self.max_cycles, self.agents = int(max_cycles), []
        self._rng, self._t = np.random.default_rng(0), 0

    def observation_space(self, agent):
        return self._obs_space[agent]

    def action_space(self, agent):
        return self._act_space[agent]

    def _obs(self):
        return {a: self._rng.standard_normal(self._obs_space[a].shape).astype(np.float32) for a in self.possible_agents}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        self._t, self.agents = 0, list(self.possible_agents)
        return self._obs(), {a: {} for a in self.agents}

    def step(self, actions):
        self._t += 1
        trunc = self._t >= self.max_cycles
        live = list(self.possible_agents)
        out = (self._obs(), {a: float(self._rng.standard_normal()) for a in live}, {a: False for a in live},
               {a: trunc for a in live}, {a: {} for a in live})
        if trunc:
            self.agents = []
        return out

    def close(self):
        pass


def mpe_modules():
    """{'pettingzoo': ..., 'pettingzoo.mpe': ..., 'pettingzoo.mpe.<name>': module with parallel_env(...)}"""
    mods = {"pettingzoo": types.ModuleType("pettingzoo"), "pettingzoo.mpe": types.ModuleType("pettingzoo.mpe")}
    mods["pettingzoo"].mpe = mods["pettingzoo.mpe"]
    mods["pettingzoo"].__freerl_b200_shim__ = True
    for name in _MPE:
        m = types.ModuleType("pettingzoo.mpe." + name)
        m.parallel_env = (lambda _n: (lambda **kw: SyntheticParallelEnv(_n, **kw)))(name)
        setattr(mods["pettingzoo.mpe"], name, m)
        mods["pettingzoo.mpe." + name] = m
    return mods


def as_module():
    m = types.ModuleType("gymnasium")
    m.make, m.spec, m.spaces = make, spec, spaces
    m.__freerl_b200_shim__ = True
    return m
