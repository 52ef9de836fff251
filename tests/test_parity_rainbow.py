"""freerl_b200 Rainbow DQN (fused C51+Dueling+Noisy kernel, device PER) vs the oracle and the reference golden."""
import numpy as np
import pytest
import torch

from oracle.rainbow import RainbowOracle
from test_oracle_rainbow import eps_of, rainbow_setup

TRICK = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}


def _run(golden, device):
    from freerl_b200.DQN_with_tricks import DQN
    g = golden("rainbow")
    q, per = rainbow_setup(g)
    orc = RainbowOracle(q, 1e-3, 4)
    pol = DQN([8, 4], False, 1e-3, 500, device, trick=TRICK, gamma=0.99, batch_size=32)
    sd = {k: v.clone() for k, v in q.items()}
    pol.agent.Qnet.load_state_dict(sd)
    pol.agent.Qnet_target.load_state_dict(sd)
    # same replay content + sum-tree as the reference run
    n = g["buf/obs"].shape[0]
    pol.buffer.buffer.add(g["buf/obs"], g["buf/act"], g["buf/rew"], g["buf/nobs"], g["buf/done"])
    assert [pol.buffer.buffer._index, pol.buffer.buffer._size] == [int(x) for x in g["init/index"]]
    pol.buffer.sumtree.tree.copy_(torch.from_numpy(g["init/tree"]))
    gamma_n = float(g["n_step_gamma"])
    assert abs(pol.buffer.n_step_gamma - gamma_n) < 1e-15
    B = 32
    for it in range(3):
        u = g["u/%d" % it]
        raw = eps_of(g, it)
        # oracle side: PER sample + learn + priorities
        seg = per.sumtree.total() / B
        per.beta = np.min([1., per.beta + per.beta_increment])
        idx = np.zeros(B, np.int64); pri = np.zeros(B, np.float32)
        for i in range(B):
            a, b = seg * i, seg * (i + 1)
            pri[i], idx[i] = per.sumtree.find(a + (b - a) * u[i])
        prob = np.clip(pri / per.sumtree.total(), 1e-7, None)
        w = (len(per) * prob) ** (-per.beta)
        w = (w / w.max()).astype(np.float32)
        batch = tuple(torch.from_numpy(x) for x in per.buffer.sample(idx))
        r = orc.learn(batch, raw, gamma_n, 0.01, is_weight=torch.from_numpy(w), double_q=True)
        per.update_priorities(idx, r["error"].numpy())
        # ours
        pol.learn(B, 0.99, 0.01, u=u, noise=[tuple(x.numpy() for x in f) for f in raw])
        assert np.array_equal(pol.last_indices.cpu().numpy(), idx)                      # bit-exact sampled indices
        loss = float(pol.last_metrics[0])
        assert abs(loss - r["loss"]) <= 1e-5 * abs(r["loss"]), (loss, r["loss"])
        assert abs(loss - g["losses"][it]) <= 1e-5 * abs(g["losses"][it])
        np.testing.assert_allclose(pol.last_error.cpu().numpy(), r["error"].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(pol.buffer.sumtree.tree.cpu().numpy(), g["tree/%d" % it], rtol=2e-6, atol=1e-9)
        assert float(pol.buffer.beta) == float(g["beta/%d" % it][1])
    got = pol.agent.Qnet.state_dict()
    got_t = pol.agent.Qnet_target.state_dict()
    for k, v in orc.q.items():
        np.testing.assert_allclose(got[k].cpu().numpy(), v.detach().numpy(), rtol=1e-5, atol=2e-6, err_msg=k)
        np.testing.assert_allclose(got[k].cpu().numpy(), g["final/q/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
        np.testing.assert_allclose(got_t[k].cpu().numpy(), g["final/q_target/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
    # state_dict schema of the checked-in Rainbow checkpoints
    assert list(got.keys()) == [k[len("init/q/"):] for k in g.files if k.startswith("init/q/")]
    a = pol.select_action(g["buf/obs"][0].astype(np.float32))
    assert 0 <= int(a) < 4
    acts = pol.select_action(g["buf/obs"][:16].astype(np.float32))
    assert acts.shape == (16,)


def test_rainbow_emulated(golden, emul):
    _run(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_rainbow_gpu(golden):
    _run(golden, torch.device("cuda"))


def _other_action_counts(device):
    """n_actions 2 and 6 (different head widths / column-block splits / scratch sizes than the 4-action golden): fused learn vs
    the oracle with explicit PER uniforms and NoisyLinear draws."""
    from collections import OrderedDict
    from freerl_b200.DQN_with_tricks import DQN
    from oracle import buffers
    rng = np.random.default_rng(21)
    for nA in (2, 6):
        torch.manual_seed(nA)
        B, od = 24, 5
        pol = DQN([od, nA], False, 1e-3, 200, device, trick=TRICK, gamma=0.97, batch_size=B)
        q = OrderedDict((k, v.detach().cpu().clone()) for k, v in pol.agent.Qnet.state_dict().items())
        orc = RainbowOracle(q, 1e-3, nA)
        per = buffers.NStepPrioritizedReplay(200, od, 1, gamma=0.97, n_step=3)
        for _ in range(90):
            o, a_, r = rng.standard_normal(od).astype(np.float32), int(rng.integers(0, nA)), float(np.float32(rng.standard_normal()))
            o2, d = rng.standard_normal(od).astype(np.float32), bool(rng.random() < 0.1)
            pol.add(o, a_, r, o2, d)
            per.add(o, a_, r, o2, d)
        assert np.array_equal(pol.buffer.sumtree.tree.cpu().numpy(), per.sumtree.tree)
        gamma_n = per.n_step_gamma
        for it in range(2):
            u = rng.random(B)
            raw = [tuple(torch.randn(n) for n in (128, 51, 128, nA * 51)) for _ in range(3)]
            seg = per.sumtree.total() / B
            per.beta = np.min([1., per.beta + per.beta_increment])
            idx, pri = np.zeros(B, np.int64), np.zeros(B, np.float32)
            for i in range(B):
                lo, hi = seg * i, seg * (i + 1)
                pri[i], idx[i] = per.sumtree.find(lo + (hi - lo) * u[i])
            prob = np.clip(pri / per.sumtree.total(), 1e-7, None)
            w = (len(per) * prob) ** (-per.beta)
            w = (w / w.max()).astype(np.float32)
            batch = tuple(torch.from_numpy(x) for x in per.buffer.sample(idx))
            r = orc.learn(batch, raw, gamma_n, 0.01, is_weight=torch.from_numpy(w), double_q=True)
            per.update_priorities(idx, r["error"].numpy())
            pol.learn(B, 0.97, 0.01, u=u, noise=[tuple(x.numpy() for x in f) for f in raw])
            assert np.array_equal(pol.last_indices.cpu().numpy(), idx)
            loss = float(pol.last_metrics[0])
            assert abs(loss - r["loss"]) <= 1e-5 * abs(r["loss"]), (nA, it, loss, r["loss"])
            np.testing.assert_allclose(pol.last_error.cpu().numpy(), r["error"].numpy(), rtol=1e-5, atol=1e-6)
        from parity_util import assert_module_close           # outlier-aware (Adam-conditioned elements, see parity_util)
        assert_module_close(pol.agent.Qnet, orc.q, "nA=%d" % nA)


def test_rainbow_other_action_counts_emulated(emul):
    _other_action_counts(torch.device("cpu"))


@pytest.mark.gpu
def test_rainbow_other_action_counts_gpu():
    _other_action_counts(torch.device("cuda"))


def test_rainbow_fast_mode_in_kernel_noise(emul):
    """fast mode: the kernel draws the factorised noise f(eps) = sign(eps) sqrt|eps| itself (Philox): fresh per learn, with the
    moments of the transform (E|f| = 0.822, E f^2 = E|eps| = 0.798), published for the weight_epsilon bookkeeping."""
    from freerl_b200.DQN_with_tricks import DQN
    torch.manual_seed(0)
    pol = DQN([8, 4], False, 1e-3, 256, torch.device("cpu"), trick=TRICK, gamma=0.99, batch_size=16, mode="fast")
    rng = np.random.default_rng(0)
    for _ in range(6):                      # 16 lock-stepped envs; the 3-step window emits from the third add on
        pol.add(rng.standard_normal((16, 8)), rng.integers(0, 4, (16, 1)), rng.standard_normal(16), rng.standard_normal((16, 8)), rng.random(16) < 0.1)
    assert len(pol.buffer) == 64
    pol.learn(16, 0.99, 0.01)
    e1 = pol._eps.clone()
    pol.learn(16, 0.99, 0.01)
    e2 = pol._eps.clone()
    used = pol.agent.eps_off["A_out"] + pol.agent.nA * pol.agent.n_atoms
    a, b = e1[:, :used].numpy(), e2[:, :used].numpy()
    assert not np.array_equal(a, b) and not np.array_equal(a[0], a[1])
    assert abs(np.abs(a).mean() - 0.822) < 0.03 and abs((a ** 2).mean() - 0.798) < 0.04 and abs(a.mean()) < 0.05
    assert np.isfinite(float(pol.last_metrics[0]))
    pol.agent._refresh_buffers()                                                   # bookkeeping reads the published noise
    we = pol.agent.Qnet.V.weight_epsilon.cpu().numpy()
    o = pol.agent.eps_off
    np.testing.assert_allclose(we, np.outer(b[2, o["V_out"]:o["V_out"] + 51], b[2, o["V_in"]:o["V_in"] + 128]), rtol=1e-6)
