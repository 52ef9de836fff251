"""Per-step train-loop helpers for N vectorised envs (freerl_b200/vecloop.py -> frl_vecnorm / frl_reward_scaling / frl_explore /
frl_masked_reset) against the fixture generated from the UNMODIFIED reference classes (oracle/make_golden_vecloop.py) and against
the oracle restatement (oracle/vecloop.py) on larger random cases.  Everything is bit-exact: float64 statistics, float64 outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import vecloop as ov

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vecloop.npz"))
STEPS, N = 6, 5


# ---- oracle pinned to the reference-generated fixture (CPU) -------------------------------------------------------
def test_oracle_vecloop_matches_reference_fixture():
    for tag in ("f32", "f64"):
        ms = ov.RunningMeanStd(7)
        x = G["norm_%s_x" % tag]
        assert np.array_equal(ov.normalize_rows(ms, x), G["norm_%s_y" % tag])
        assert np.array_equal(ov.normalize_rows(ms, x[:4], update=False), G["norm_%s_eval" % tag])
        assert np.array_equal(np.asarray(ms.std, np.float64), G["norm_%s_std" % tag])
    ms, R = ov.RunningMeanStd(1), np.zeros(N)
    for t in range(STEPS):
        assert np.array_equal(ov.reward_scaling_rows(ms, R, float(G["rs_gamma"]), G["rs_x"][t]), G["rs_y"][t])
        R[G["rs_done"][t]] = 0.0
    st = np.zeros((N, 3))
    for t in range(STEPS):
        nz = ov.ou_rows(st, G["ou_z"][t], sigma=0.2, dt=1e-2, scale=0.3)
        assert np.array_equal(nz, G["ou_noise"][t])
        assert np.array_equal(ov.explore_ou(G["ou_act"][t], nz, float(G["ou_max_action"])), G["ou_out"][t])
        if t == 2:
            st[[1, 3]] = 0.0
        assert np.array_equal(ov.explore_gauss(G["ou_act"][t], G["gauss_z"][t], 2.0, float(G["gauss_scale"]), float(G["gauss_sigma"])),
                              G["gauss_out"][t])


    a1 = np.stack([ov.dis_to_con(a, np.array([-2.0], np.float32), np.array([2.0], np.float32), 11) for a in G["d2c1_a"]])
    assert np.array_equal(a1.astype(np.float64), G["d2c1_out"])
    a4 = np.stack([ov.dis_to_con(a, G["d2c4_low"], G["d2c4_high"], 81) for a in G["d2c4_a"]])
    assert np.array_equal(a4.astype(np.float64), G["d2c4_out"])
    np.random.seed(23)
    assert np.array_equal(ov.epsilon_greedy(G["eg_greedy"], 4, 0.3), G["eg_out"])


# ---- product (C ABI) vs fixture and oracle ------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


def _run(device):
    from freerl_b200 import vecloop as vl
    # Normalization: [N, D] blocks in env order == the reference object called row by row
    for tag in ("f32", "f64"):
        x, want = G["norm_%s_x" % tag], G["norm_%s_y" % tag]
        nm = vl.Normalization(7, device)
        got = np.concatenate([_np(nm(x[t * N:(t + 1) * N], out_dtype=torch.float64)) for t in range(STEPS)])
        assert np.array_equal(got, want), tag
        assert nm.running_ms.n == STEPS * N
        for name in ("mean", "S", "std"):
            assert np.array_equal(_np(getattr(nm.running_ms, name)), G["norm_%s_%s" % (tag, name)]), (tag, name)
        assert np.array_equal(_np(nm(x[:4], update=False, out_dtype=torch.float64)), G["norm_%s_eval" % tag])
        # fp32 output = what the replay hands to the networks (float64 result rounded once)
        nm2 = vl.Normalization(7, device)
        assert np.array_equal(_np(nm2(x[:N])), want[:N].astype(np.float32))
    # reward_norm = Normalization(shape=1) on float64 rewards
    nm = vl.Normalization(1, device)
    got = np.concatenate([_np(nm(G["rnorm_x"][t * N:(t + 1) * N], out_dtype=torch.float64)).reshape(-1) for t in range(STEPS)])
    assert np.array_equal(got, G["rnorm_y"])
    # RewardScaling: vector of envs with per-env resets, and the single-env reference object with a reset
    rs = vl.RewardScaling(1, float(G["rs_gamma"]), n_envs=N, device=device)
    for t in range(STEPS):
        assert np.array_equal(_np(rs(G["rs_x"][t], out_dtype=torch.float64)), G["rs_y"][t]), t
        rs.reset(G["rs_done"][t])
    assert np.array_equal(_np(rs.running_ms.std), G["rs_std"])
    rs1 = vl.RewardScaling(1, 0.99, n_envs=1, device=device)
    for i, v in enumerate(G["rs1_x"]):
        assert _np(rs1(np.array([v]), out_dtype=torch.float64))[0] == G["rs1_y"][i]
        if i == 6:
            rs1.reset()
    # OUNoise + clip, Gaussian exploration
    ou = vl.OUNoise(3, sigma=0.2, dt=1e-2, scale=0.3, n_envs=N, device=device)
    ou2 = vl.OUNoise(3, sigma=0.2, dt=1e-2, scale=0.3, n_envs=N, device=device)
    for t in range(STEPS):
        assert np.array_equal(_np(ou.noise(z=G["ou_z"][t])), G["ou_noise"][t]), t
        assert np.array_equal(_np(ou2.explore(G["ou_act"][t], float(G["ou_max_action"]), z=G["ou_z"][t])), G["ou_out"][t]), t
        if t == 2:
            m = np.zeros(N, bool); m[[1, 3]] = True
            ou.reset(m); ou2.reset(m)
        g = vl.explore_gauss(G["ou_act"][t], 2.0, float(G["gauss_scale"]), float(G["gauss_sigma"]), device=device, z=G["gauss_z"][t])
        assert np.array_equal(_np(g), G["gauss_out"][t]), t
    # parity mode consumes the legacy numpy stream exactly like N reference calls
    np.random.seed(7)
    ou3 = vl.OUNoise(3, sigma=0.2, dt=1e-2, scale=0.3, n_envs=N, device=device)
    assert np.array_equal(_np(ou3.noise()), G["ou_noise"][0])
    # dis_to_con and epsilon-greedy of the DQN mains
    assert np.array_equal(_np(vl.dis_to_con(G["d2c1_a"], [-2.0], [2.0], 11, device=device)), G["d2c1_out"])
    assert np.array_equal(_np(vl.dis_to_con(G["d2c4_a"], G["d2c4_low"], G["d2c4_high"], 81, device=device)), G["d2c4_out"])
    np.random.seed(23)
    assert np.array_equal(_np(vl.epsilon_greedy(G["eg_greedy"], 4, 0.3, device=device)), G["eg_out"])
    gr = np.zeros(20000, dtype=np.int64) + 7
    fe = _np(vl.epsilon_greedy(gr, 4, 0.25, device=device, mode="fast", seed=5, counter=1))
    fired = fe != 7
    assert 0.22 < fired.mean() < 0.28 and set(np.unique(fe[fired])) == {0, 1, 2, 3}
    assert np.array_equal(fe, _np(vl.epsilon_greedy(gr, 4, 0.25, device=device, mode="fast", seed=5, counter=1)))
    # larger random cases vs the oracle: 1024 envs (C3), wide observations (more than one CTA of columns), many steps
    rng = np.random.default_rng(5)
    for n_env, d, dt in ((1024, 8, np.float32), (64, 300, np.float64), (3, 54, np.float32)):
        nm, ms = vl.Normalization(d, device), ov.RunningMeanStd(d)
        for _ in range(3):
            x = (rng.standard_normal((n_env, d)) * 2 - 0.5).astype(dt)
            assert np.array_equal(_np(nm(x, out_dtype=torch.float64)), ov.normalize_rows(ms, x)), (n_env, d)
        assert np.array_equal(_np(nm.running_ms.S), np.asarray(ms.S, np.float64))
    rs, ms, R = vl.RewardScaling(1, 0.97, n_envs=512, device=device), ov.RunningMeanStd(1), np.zeros(512)
    for _ in range(3):
        x = rng.standard_normal(512)
        assert np.array_equal(_np(rs(x, out_dtype=torch.float64)), ov.reward_scaling_rows(ms, R, 0.97, x))
    # fast mode: Philox normals on the device — bounded, reproducible per (seed, counter), different across calls
    f1 = vl.OUNoise(6, n_envs=256, device=device, mode="fast", seed=3)
    f2 = vl.OUNoise(6, n_envs=256, device=device, mode="fast", seed=3)
    a, b = _np(f1.noise()), _np(f2.noise())
    assert np.array_equal(a, b) and np.isfinite(a).all() and 0.005 < a.std() < 0.02          # sqrt(dt) * sigma = 0.01
    assert not np.array_equal(_np(f1.noise()), a)


def test_vecloop_emulated(emul):
    _run(torch.device("cpu"))


@pytest.mark.gpu
def test_vecloop_gpu():
    _run(torch.device("cuda"))


def test_vecloop_random_interleavings_emulated(emul):
    """Random env counts / widths / dtypes with update and evaluation calls interleaved, against the oracle fed row by row."""
    from freerl_b200 import vecloop as vl
    device = torch.device("cpu")
    rng = np.random.default_rng(77)
    for trial in range(12):
        d, dt = int(rng.integers(1, 40)), (np.float32 if trial % 2 else np.float64)
        nm, ms = vl.Normalization(d, device), ov.RunningMeanStd(d)
        for step in range(int(rng.integers(2, 7))):
            n_env = int(rng.choice([1, 2, 5, 33, 130]))
            x = (rng.standard_normal((n_env, d)) * rng.uniform(0.1, 5) + rng.uniform(-3, 3)).astype(dt)
            upd = bool(rng.random() < 0.7) or step == 0
            assert np.array_equal(_np(nm(x, update=upd, out_dtype=torch.float64)), ov.normalize_rows(ms, x, update=upd)), (trial, step, upd)
        assert nm.running_ms.n == ms.n
        assert np.array_equal(_np(nm.running_ms.mean), np.asarray(ms.mean, np.float64).reshape(-1))
        assert np.array_equal(_np(nm.running_ms.std), np.asarray(ms.std, np.float64).reshape(-1))
    for trial in range(6):
        n_env, gamma = int(rng.choice([1, 3, 64])), float(rng.uniform(0.9, 0.999))
        rs, ms, R = vl.RewardScaling(1, gamma, n_envs=n_env, device=device), ov.RunningMeanStd(1), np.zeros(n_env)
        for step in range(8):
            x = rng.standard_normal(n_env) * 3
            assert np.array_equal(_np(rs(x, out_dtype=torch.float64)), ov.reward_scaling_rows(ms, R, gamma, x))
            done = rng.random(n_env) < 0.3
            rs.reset(done)
            R[done] = 0.0
