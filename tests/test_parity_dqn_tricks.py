"""freerl_b200.DQN_with_tricks non-distributional branch (frl_dqn_learn with Double / Dueling / PER / N-step flags) vs the
oracle and the fixtures generated from the unmodified reference (DQN_file/DQN_with_tricks.py:261-283)."""
import numpy as np
import pytest
import torch

from oracle.dqn_tricks import DQNTricksOracle
from test_oracle_dqn_tricks import CASES, per_draw, tricks_setup

TOL = dict(rtol=1e-5, atol=2e-6)


def _run(golden, device, name):
    from freerl_b200.DQN_with_tricks import DQN
    g, cfg = golden("dqn_tricks_" + name), CASES[name]
    trick = {"Double": cfg["double_q"], "Dueling": cfg["dueling"], "PER": cfg["per"], "Noisy": False, "N_Step": cfg["per"], "Categorical": False}
    q, store = tricks_setup(g, cfg["per"])
    orc = DQNTricksOracle(q, 1e-3, dueling=cfg["dueling"], double_q=cfg["double_q"])
    pol = DQN([8, 4], False, 1e-3, 500, device, trick=trick, gamma=0.99, batch_size=32)
    sd = {k: v.clone() for k, v in q.items()}
    assert list(pol.agent.Qnet.state_dict().keys()) == list(sd.keys())               # reference checkpoint schema
    pol.agent.Qnet.load_state_dict(sd)
    pol.agent.Qnet_target.load_state_dict(sd)
    buf = pol.buffer.buffer if cfg["per"] else pol.buffer
    buf.add(g["buf/obs"], g["buf/act"], g["buf/rew"], g["buf/nobs"], g["buf/done"])
    assert [buf._index, buf._size] == [int(x) for x in g["init/index"]]
    if cfg["per"]:
        pol.buffer.sumtree.tree.copy_(torch.from_numpy(g["init/tree"]))
        assert abs(pol.buffer.n_step_gamma - float(g["gamma_used"])) < 1e-15
    gamma = float(g["gamma_used"])
    # greedy actions before training == the reference's select_action on the same observations
    load_acts = None
    for it in range(4):
        if cfg["per"]:
            u = g["u/%d" % it]
            idx, w = per_draw(store, u, 32)
            batch = tuple(torch.from_numpy(x) for x in store.buffer.sample(idx))
            r = orc.learn(batch, gamma, 0.01, is_weight=torch.from_numpy(w))
            store.update_priorities(idx, r["td_error"].numpy())
            pol.learn(32, 0.99, 0.01, u=u)
            assert np.array_equal(pol.last_indices.cpu().numpy(), idx)                 # bit-exact sampled indices
            np.testing.assert_allclose(pol.buffer.sumtree.tree.cpu().numpy(), g["tree/%d" % it], rtol=2e-6, atol=1e-9)
        else:
            idx = g["idx/%d" % it]
            batch = tuple(torch.from_numpy(x) for x in store.sample(idx))
            r = orc.learn(batch, gamma, 0.01)
            pol.learn(32, 0.99, 0.01, indices=idx)
        loss = float(pol.last_metrics[0])
        assert abs(loss - r["loss"]) <= 1e-5 * abs(r["loss"]), (it, loss, r["loss"])
        assert abs(loss - g["losses"][it]) <= 1e-5 * abs(g["losses"][it])
        np.testing.assert_allclose(pol.last_error.cpu().numpy(), r["td_error"].numpy(), rtol=1e-5, atol=2e-6)
        got = pol.agent.Qnet.state_dict()
        for k, v in orc.q.items():
            np.testing.assert_allclose(got[k].cpu().numpy(), v.detach().numpy(), err_msg="%s after learn %d" % (k, it), **TOL)
    got, got_t = pol.agent.Qnet.state_dict(), pol.agent.Qnet_target.state_dict()
    for k in orc.q:
        np.testing.assert_allclose(got[k].cpu().numpy(), g["final/q/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
        np.testing.assert_allclose(got_t[k].cpu().numpy(), g["final/q_target/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
    assert np.array_equal(pol.select_action(g["act/obs"]), g["act/action"])            # reference greedy actions, batched
    assert int(pol.select_action(g["act/obs"][3])) == int(g["act/action"][3])


@pytest.mark.parametrize("name", sorted(CASES))
def test_dqn_tricks_emulated(golden, emul, name):
    _run(golden, torch.device("cpu"), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_dqn_tricks_gpu(golden, name):
    _run(golden, torch.device("cuda"), name)


def test_trick_dispatch_errors(emul):
    from freerl_b200.DQN_with_tricks import DQN
    base = {"Double": False, "Dueling": False, "PER": False, "Noisy": True, "N_Step": False, "Categorical": False}
    with pytest.raises(NotImplementedError):
        DQN([8, 4], False, 1e-3, 100, torch.device("cpu"), trick=base, gamma=0.99, batch_size=8)
