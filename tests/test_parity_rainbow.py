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
