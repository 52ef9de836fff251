"""CPU restatement of the per-step helpers of the reference train loops, fed row by row.

TEST INFRASTRUCTURE ONLY — only ``tests/`` may import this; the product path (``freerl_b200/vecloop.py``) runs CUDA kernels.
Pinned by ``tests/golden/vecloop.npz``, generated from the unmodified reference classes by ``oracle/make_golden_vecloop.py``.
"""
import numpy as np


class RunningMeanStd:
    """Welford statistics over single observations as ``PPO_file/normalization.py:17-35`` keeps them (same code in
    ``MAPPO_file/normalization.py``, ``DDPG_file/DDPG.py:358-376``, ``SAC_file/SAC.py:357-375``).  Restated as one fold step:

    * call 1 ALIASES the observation: ``mean`` and ``std`` both become ``x`` itself (so they carry x's dtype — float32 observations
      give a float32 mean from then on — and the first normalised output is exactly 0);
    * call n > 1: ``mean += (x - mean) / n`` in mean's dtype (python-int ``n`` is a weak scalar under NumPy 2), the float64
      accumulator ``S += (x - mean_old) * (x - mean_new)``, ``std = sqrt(S / n)`` (float64).
    """

    def __init__(self, shape):
        self.n, self.mean, self.S = 0, np.zeros(shape), np.zeros(shape)
        self.std = np.sqrt(self.S)

    def update(self, x):
        x = np.array(x)
        self.n += 1
        if self.n == 1:
            self.mean = self.std = x
            return
        before = self.mean.copy()
        step = (x - before) / self.n
        self.mean = before + step
        self.S = self.S + (x - before) * (x - self.mean)
        self.std = np.sqrt(self.S / self.n)


def normalize_rows(ms, rows, update=True):
    """``Normalization.__call__`` (``normalization.py:38-49``) applied to each row in order; returns the stacked float64 results."""
    out = []
    for x in rows:
        if update:
            ms.update(x)
        out.append(np.asarray((x - ms.mean) / (ms.std + 1e-8), dtype=np.float64))
    return np.stack(out)


def reward_scaling_rows(ms, R, gamma, rewards):
    """``RewardScaling.__call__`` (``normalization.py:87-97``) with one discounted return per env and a shared RunningMeanStd fed in
    env order.  ``R`` is modified in place."""
    out = np.empty(len(rewards))
    for i, x in enumerate(rewards):
        R[i] = gamma * R[i] + x
        ms.update(R[i:i + 1].copy())
        out[i] = (x / (ms.std + 1e-8))[0]
    return out


def ou_rows(state, z, mu=0.0, theta=0.15, sigma=0.1, dt=1e-2, scale=None):
    """``OUNoise.noise`` (``SAC_file/SAC.py:347-355``) per env row with the given standard normals; ``state`` updated in place."""
    out = np.empty_like(state)
    for i in range(state.shape[0]):
        x = state[i]
        dx = theta * (mu - x) + np.sqrt(dt) * sigma * z[i]
        state[i] = x + dx
        out[i] = state[i] if scale is None else state[i] * scale
    return out


def explore_ou(action, noise, max_action):
    """``DDPG_file/DDPG.py:520``"""
    return np.clip(action * max_action + noise * max_action, -max_action, max_action)


def explore_gauss(action, z, max_action, gauss_scale, gauss_sigma):
    """``DDPG_file/DDPG.py:522``: ``np.random.normal(scale=s, size)`` is ``0.0 + s * z`` on the legacy stream."""
    return np.clip(action * max_action + gauss_scale * (0.0 + (gauss_sigma * max_action) * z), -max_action, max_action)


def dis_to_con(discrete_action, low, high, action_dim):
    """``DQN_file/DQN.py:195-217`` for one np.int64 action; ``low`` / ``high`` float32 arrays."""
    if len(low) == 1:
        return np.array([low[0] + (discrete_action / (action_dim - 1)) * (high[0] - low[0])])
    per = int(action_dim ** (1 / len(low)))
    idx = [discrete_action // (per ** i) % per for i in range(len(low))]
    return np.array([low[i] + idx[i] / (per - 1) * (high[i] - low[i]) for i in range(len(low))])


def epsilon_greedy(greedy, n_actions, epsilon):
    """``DQN_file/DQN.py:307-310`` iterated over envs on numpy's legacy global stream."""
    out = np.empty(len(greedy), dtype=np.int64)
    for i, gA in enumerate(greedy):
        out[i] = np.random.randint(n_actions) if np.random.rand() < epsilon else gA
    return out
