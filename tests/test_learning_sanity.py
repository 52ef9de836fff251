"""End-to-end learning sanity on the GPU (beyond per-step parity): the fused kernels actually improve a policy on toy vectorised
environments stepped on the host — SAC (continuous), DQN (discrete), PPO (continuous, GAE + minibatch epochs)."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class PointEnv:
    """N independent 1-D points: obs = [x, target], x' = clip(x + 0.2 a), reward = -|x' - target|, episodes of 20 steps."""

    def __init__(self, n, seed):
        self.n, self.rng, self.t = n, np.random.default_rng(seed), 0
        self.reset()

    def reset(self):
        self.x = self.rng.uniform(-1, 1, self.n).astype(np.float32)
        self.g = self.rng.uniform(-1, 1, self.n).astype(np.float32)
        self.t = 0
        return self.obs()

    def obs(self):
        return np.stack([self.x, self.g], axis=1)

    def step(self, a):
        self.x = np.clip(self.x + 0.2 * np.asarray(a, np.float32).reshape(self.n), -1.5, 1.5)
        self.t += 1
        r = -np.abs(self.x - self.g)
        trunc = self.t >= 20
        nxt = self.obs()
        if trunc:
            self.reset()
        return nxt, r.astype(np.float32), np.zeros(self.n, bool), np.full(self.n, trunc)


def _quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def test_sac_learns_point_env():
    from freerl_b200.SAC import SAC
    torch.manual_seed(0); np.random.seed(0)
    dev = torch.device("cuda")
    pol = _quiet(lambda: SAC([2, 1], True, 1e-3, 1e-3, 100_000, dev, trick={}, mode="fast"))
    env = PointEnv(64, 0)
    obs, rets = env.obs(), []
    for step in range(500):
        act = pol.select_action(obs) if step >= 20 else np.random.uniform(-1, 1, (64, 1)).astype(np.float32)
        nxt, r, term, trunc = env.step(act)
        pol.add(obs, act, r, nxt if not trunc.any() else nxt, term)      # truncation: done stays False (bootstrapped)
        obs = env.obs()
        rets.append(float(r.mean()))
        if step >= 20:
            pol.learn(256, 0.95, 0.01, n_updates=16)
    random_phase, late = np.mean(rets[:20]), np.mean(rets[-100:])     # the first 20 steps act uniformly at random (~ -0.6)
    assert np.isfinite(pol.last_metrics.cpu().numpy()).all()
    assert late > random_phase + 0.3 and late > -0.2, (random_phase, late)   # measured on B200: -0.19 after 50 steps, -0.08 at the end


def test_dqn_learns_point_env():
    from freerl_b200.DQN import DQN
    torch.manual_seed(0); np.random.seed(0)
    dev = torch.device("cuda")
    pol = _quiet(lambda: DQN([2, 3], False, 1e-3, 100_000, dev, mode="fast"))
    env = PointEnv(64, 1)
    amap = np.array([-1.0, 0.0, 1.0], np.float32)
    obs, rets = env.obs(), []
    for step in range(500):
        eps = max(0.05, 1.0 - step / 200)
        greedy = pol.select_action(obs)
        rand = np.random.randint(0, 3, 64)
        a = np.where(np.random.rand(64) < eps, rand, greedy)
        nxt, r, term, trunc = env.step(amap[a])
        pol.add(obs, a.reshape(-1, 1), r, nxt, term)
        obs = env.obs()
        rets.append(float(r.mean()))
        if step >= 20:
            pol.learn(256, 0.95, 0.01, n_updates=16)
    early, late = np.mean(rets[:50]), np.mean(rets[-100:])
    assert late > early + 0.15 and late > -0.4, (early, late)


def test_ppo_learns_point_env():
    from freerl_b200.PPO_advance import PPO
    torch.manual_seed(0); np.random.seed(0)
    dev = torch.device("cuda")
    N, T = 64, 40
    pol = _quiet(lambda: PPO([2, 1], True, 3e-3, 3e-3, N * T, dev, trick={"adv_norm": False}, mode="fast"))
    env = PointEnv(N, 2)
    obs, curve = env.obs(), []
    for it in range(25):
        rs = []
        for t in range(T):
            act, logp = pol.select_action(obs)
            nxt, r, term, trunc = env.step(np.clip(act, -1, 1))
            pol.add(obs, act, r, nxt, term, logp, term | trunc)
            obs = env.obs()
            rs.append(float(r.mean()))
        curve.append(np.mean(rs))
        pol.learn(512, 0.95, 0.95, 0.2, 4, 0.0)
    early, late = np.mean(curve[:3]), np.mean(curve[-5:])
    assert late > early + 0.1, (early, late, curve)
