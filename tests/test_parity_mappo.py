"""freerl_b200.MAPPO (fused GAE / adv-norm / LayerNorm PPO kernels) vs the oracle and the reference golden."""
import numpy as np
import pytest
import torch

from oracle.make_golden_marl import MAPPO_TRICK
from oracle.marl import MAPPOOracle
from parity_util import assert_module_close, load_into
from test_oracle_marl import IDS, maddpg_nets, mappo_data


def _run(golden, device):
    from freerl_b200.MAPPO import MAPPO
    g = golden("mappo")
    dim_info = {k: [18, 5] for k in IDS}
    pol = MAPPO(dim_info, True, 1e-3, 1e-3, 64, device, dict(MAPPO_TRICK))
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k])
        load_into(pol.agents[k].critic, ic[k])
    orc = MAPPOOracle(ia, ic, 1e-3, MAPPO_TRICK)
    data = mappo_data(g)
    d = {k: [x.numpy() for x in data[k]] for k in IDS}
    for t in range(64):
        pol.add({k: d[k][0][t] for k in IDS}, {k: d[k][1][t] for k in IDS}, {k: float(d[k][2][t, 0]) for k in IDS},
                {k: d[k][3][t] for k in IDS}, {k: bool(d[k][4][t, 0]) for k in IDS}, {k: d[k][5][t] for k in IDS},
                {k: bool(d[k][6][t, 0]) for k in IDS})
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(data, perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    pol.learn(32, 0.95, 0.95, 0.2, 2, 0.01, 10.0, permutations=perms)
    np.testing.assert_allclose(pol.last_adv.cpu().numpy(), r["adv"].numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pol.last_v_target.cpu().numpy(), r["v_target"].numpy(), rtol=1e-5, atol=2e-6)
    m = pol.last_metrics.cpu().numpy()
    ref = np.array(r["losses"])
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, :2], g["losses"], rtol=5e-5, atol=5e-6)
    tol = dict(rtol=1e-4, atol=1e-5)
    for k in IDS:
        assert_module_close(pol.agents[k].actor, orc.actor[k], "actor " + k, tol)
        assert_module_close(pol.agents[k].critic, orc.critic[k], "critic " + k, tol)
        assert_module_close(pol.agents[k].actor, maddpg_nets(g, "final", "actor")[k], "actor vs reference " + k, tol)
    acts, lps = pol.select_action({k: d[k][0][0] for k in IDS})
    assert acts["agent_0"].shape == (5,) and lps["agent_0"].shape == (5,)


def test_mappo_emulated(golden, emul):
    _run(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_mappo_gpu(golden):
    _run(golden, torch.device("cuda"))
