"""oracle/dqn_tricks.py (non-distributional DQN_with_tricks branch: Double / Dueling / PER / N-step) replayed against the
fixtures generated from the unmodified reference."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import buffers
from oracle.dqn_tricks import DQNTricksOracle

CASES = {"double": dict(dueling=False, double_q=True, per=False), "dueling": dict(dueling=True, double_q=False, per=False),
         "d3per": dict(dueling=True, double_q=True, per=True)}


def tricks_setup(g, per):
    q = OrderedDict((k[len("init/q/"):], torch.from_numpy(g[k].copy())) for k in g.files if k.startswith("init/q/"))
    store = buffers.PrioritizedReplay(500, 8, 1) if per else buffers.RingReplay(500, 8, 1)
    b = store.buffer if per else store
    n = g["buf/obs"].shape[0]
    b.obs[:n], b.actions[:n], b.rewards[:n] = g["buf/obs"], g["buf/act"], g["buf/rew"]
    b.next_obs[:n], b.dones[:n] = g["buf/nobs"], g["buf/done"]
    b._index, b._size = [int(x) for x in g["init/index"]]
    if per:
        store.sumtree.tree[:] = g["init/tree"]
    return q, store


def per_draw(per, u, B):
    """PER_Buffer.sample with recorded unit uniforms (DQN_file/Buffer.py:99-124)"""
    seg = per.sumtree.total() / B
    per.beta = np.min([1., per.beta + per.beta_increment])
    idx, pri = np.zeros(B, np.int64), np.zeros(B, np.float32)
    for i in range(B):
        a, b = seg * i, seg * (i + 1)
        pri[i], idx[i] = per.sumtree.find(a + (b - a) * u[i])
    prob = np.clip(pri / per.sumtree.total(), 1e-7, None)
    w = (len(per) * prob) ** (-per.beta)
    return idx, (w / w.max()).astype(np.float32)


@pytest.mark.parametrize("name", sorted(CASES))
def test_dqn_tricks_oracle_vs_reference(golden, name):
    g, cfg = golden("dqn_tricks_" + name), CASES[name]
    q, store = tricks_setup(g, cfg["per"])
    orc = DQNTricksOracle(q, 1e-3, dueling=cfg["dueling"], double_q=cfg["double_q"])
    gamma = float(g["gamma_used"])
    for it in range(4):
        if cfg["per"]:
            idx, w = per_draw(store, g["u/%d" % it], 32)
            batch = tuple(torch.from_numpy(x) for x in store.buffer.sample(idx))
            r = orc.learn(batch, gamma, 0.01, is_weight=torch.from_numpy(w))
            store.update_priorities(idx, r["td_error"].numpy())
            np.testing.assert_allclose(store.sumtree.tree, g["tree/%d" % it], rtol=1e-6, atol=1e-9)
        else:
            batch = tuple(torch.from_numpy(x) for x in store.sample(g["idx/%d" % it]))
            r = orc.learn(batch, gamma, 0.01)
        np.testing.assert_allclose(r["loss"], g["losses"][it], rtol=1e-6)
        for k, v in orc.q.items():
            np.testing.assert_allclose(v.detach().numpy(), g["after/%d/q/%s" % (it, k)], rtol=2e-5, atol=2e-6, err_msg=k)
    for k, v in orc.q_target.items():
        np.testing.assert_allclose(v.detach().numpy(), g["final/q_target/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
