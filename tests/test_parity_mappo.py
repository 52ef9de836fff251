"""freerl_b200.MAPPO (fused GAE / adv-norm / LayerNorm PPO kernels) vs the oracle and the reference golden."""
import numpy as np
import pytest
import torch

from oracle.make_golden_marl import MAPPO_TRICK
from oracle.marl import MAPPOOracle
from parity_util import assert_module_close, load_into
from test_oracle_marl import IDS, maddpg_nets, mappo_data


def _run(golden, device, is_continue=True):
    from freerl_b200.MAPPO import MAPPO
    g = golden("mappo" if is_continue else "mappo_disc")
    dim_info = {k: [18, 5] for k in IDS}
    pol = MAPPO(dim_info, is_continue, 1e-3, 1e-3, 64, device, dict(MAPPO_TRICK))
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k])
        load_into(pol.agents[k].critic, ic[k])
    orc = MAPPOOracle(ia, ic, 1e-3, MAPPO_TRICK, is_continue=is_continue)
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
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, :2], g["losses"], rtol=1e-5, atol=2e-6)
    tol = dict(rtol=1e-4, atol=1e-5)
    for k in IDS:
        assert_module_close(pol.agents[k].actor, orc.actor[k], "actor " + k, tol)
        assert_module_close(pol.agents[k].critic, orc.critic[k], "critic " + k, tol)
        assert_module_close(pol.agents[k].actor, maddpg_nets(g, "final", "actor")[k], "actor vs reference " + k, tol)
    acts, lps = pol.select_action({k: d[k][0][0] for k in IDS})
    if is_continue:
        assert acts["agent_0"].shape == (5,) and lps["agent_0"].shape == (5,)
    else:
        assert 0 <= int(acts["agent_0"]) < 5 and float(lps["agent_0"]) <= 0.0
        ev = pol.evaluate_action({k: d[k][0][:4] for k in IDS})
        assert ev["agent_1"].shape == (4,)


def test_mappo_emulated(golden, emul):
    _run(golden, torch.device("cpu"))


def test_mappo_discrete_emulated(golden, emul):
    _run(golden, torch.device("cpu"), is_continue=False)


@pytest.mark.gpu
def test_mappo_discrete_gpu(golden):
    _run(golden, torch.device("cuda"), is_continue=False)


@pytest.mark.gpu
def test_mappo_gpu(golden):
    _run(golden, torch.device("cuda"))


def _ippo(golden, device, name, is_continue):
    """freerl_b200.IPPO (independent per-agent PPO on the fused kernels) vs oracle + MAPPO_file/IPPO.py golden"""
    from freerl_b200.IPPO import IPPO
    from oracle.marl import IPPOOracle
    from test_oracle_marl import IDS, IPPO_TRICK, ippo_data, maddpg_nets
    from parity_util import assert_module_close, load_into
    g = golden(name)
    pol = IPPO({k: [18, 5] for k in IDS}, is_continue, 1e-3, 5e-4, 64, device, dict(IPPO_TRICK))
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k])
        load_into(pol.agents[k].critic, ic[k])
    orc = IPPOOracle(ia, ic, 1e-3, 5e-4, IPPO_TRICK, is_continue)
    data = ippo_data(g)
    for t in range(64):
        d = {n: {k: data[k][i][t].numpy() for k in IDS} for i, n in enumerate(("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))}
        pol.add(d["obs"], d["act"], {k: float(v[0]) for k, v in d["rew"].items()}, d["nobs"], {k: bool(v[0]) for k, v in d["done"].items()},
                d["logp"], {k: bool(v[0]) for k, v in d["adv_done"].items()})
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(data, perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    pol.learn(32, 0.95, 0.95, 0.2, 2, 0.01, 10.0, permutations=perms)
    for k in IDS:
        np.testing.assert_allclose(pol.last_adv[k].cpu().numpy(), r["adv"][k].numpy(), rtol=2e-5, atol=5e-6)
        np.testing.assert_allclose(pol.last_v_target[k].cpu().numpy(), r["v_target"][k].numpy(), rtol=1e-5, atol=2e-6)
    m = pol.last_metrics.cpu().numpy()
    ref = np.array(r["losses"])
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, :2], g["losses"], rtol=1e-4, atol=1e-5)
    tol = dict(rtol=5e-5, atol=5e-6)
    for k in IDS:
        assert_module_close(pol.agents[k].actor, orc.actor[k], "actor " + k, tol)
        assert_module_close(pol.agents[k].critic, orc.critic[k], "critic " + k, tol)
        assert_module_close(pol.agents[k].actor, maddpg_nets(g, "final", "actor")[k], "final actor " + k, tol)
        assert_module_close(pol.agents[k].critic, maddpg_nets(g, "final", "critic")[k], "final critic " + k, tol)
    ev = pol.evaluate_action({k: g["act/%s/obs" % k] for k in IDS})
    for k in IDS:
        if is_continue:
            np.testing.assert_allclose(ev[k], g["act/%s/eval" % k], rtol=1e-5, atol=2e-6)
        else:
            assert int(ev[k]) == int(g["act/%s/eval" % k])
    a, lp = pol.select_action({k: g["act/%s/obs" % k] for k in IDS})
    assert set(a) == set(IDS) and (np.asarray(lp["agent_0"]).shape == ((5,) if is_continue else ()))


def test_ippo_continuous_emulated(golden, emul):
    _ippo(golden, torch.device("cpu"), "ippo_cont", True)


def test_ippo_discrete_emulated(golden, emul):
    _ippo(golden, torch.device("cpu"), "ippo_disc", False)


@pytest.mark.gpu
def test_ippo_continuous_gpu(golden):
    _ippo(golden, torch.device("cuda"), "ippo_cont", True)


@pytest.mark.gpu
def test_ippo_discrete_gpu(golden):
    _ippo(golden, torch.device("cuda"), "ippo_disc", False)


def _happo(golden, device):
    """freerl_b200.HAPPO (sequential agents, factor folded into the advantages) vs oracle + MAPPO_file/HAPPO.py golden"""
    from freerl_b200.HAPPO import HAPPO
    from oracle.marl import HAPPOOracle
    g = golden("happo")
    pol = HAPPO({k: [18, 5] for k in IDS}, True, 1e-3, 5e-4, 64, device, dict(MAPPO_TRICK))
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k])
        load_into(pol.agents[k].critic, ic[k])
    orc = HAPPOOracle(ia, ic, 1e-3, 5e-4, MAPPO_TRICK)
    data = mappo_data(g)
    d = {k: [x.numpy() for x in data[k]] for k in IDS}
    for t in range(64):
        pol.add({k: d[k][0][t] for k in IDS}, {k: d[k][1][t] for k in IDS}, {k: float(d[k][2][t, 0]) for k in IDS},
                {k: d[k][3][t] for k in IDS}, {k: bool(d[k][4][t, 0]) for k in IDS}, {k: d[k][5][t] for k in IDS},
                {k: bool(d[k][6][t, 0]) for k in IDS})
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(data, g["order"], perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    pol.learn(32, 0.95, 0.95, 0.2, 2, 0.01, 10.0, permutations=perms, order=g["order"])
    np.testing.assert_allclose(pol.last_factor.cpu().numpy(), r["factor"], rtol=5e-5, atol=1e-6)
    m = pol.last_metrics.cpu().numpy()
    ref = np.array(r["losses"])
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, :2], g["losses"], rtol=1e-4, atol=1e-5)
    tol = dict(rtol=1e-4, atol=1e-5)
    for k in IDS:
        assert_module_close(pol.agents[k].actor, orc.actor[k], "actor " + k, tol)
        assert_module_close(pol.agents[k].critic, orc.critic[k], "critic " + k, tol)
        assert_module_close(pol.agents[k].actor, maddpg_nets(g, "final", "actor")[k], "final actor " + k, tol)
        assert_module_close(pol.agents[k].critic, maddpg_nets(g, "final", "critic")[k], "final critic " + k, tol)


def _happo_discrete(golden, device):
    """HAPPO with Categorical actors against the fixture generated from MAPPO_file/HAPPO.py (happo_disc.npz, minibatch == horizon — the
    only setting in which the reference's discrete factor update broadcasts, HAPPO.py:449-450): losses of every update of every agent in
    visiting order and the final networks are the reference's own numbers; any other minibatch size raises like upstream."""
    from freerl_b200.HAPPO import HAPPO
    g = golden("happo_disc")
    pol = HAPPO({k: [18, 5] for k in IDS}, False, 1e-3, 5e-4, 64, device, dict(MAPPO_TRICK))
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k])
        load_into(pol.agents[k].critic, ic[k])
    d = {k: [g["data/%s/%s" % (k, n)] for n in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done")] for k in IDS}
    for t in range(64):
        pol.add({k: d[k][0][t] for k in IDS}, {k: d[k][1][t] for k in IDS}, {k: float(d[k][2][t, 0]) for k in IDS},
                {k: d[k][3][t] for k in IDS}, {k: bool(d[k][4][t, 0]) for k in IDS}, {k: d[k][5][t] for k in IDS},
                {k: bool(d[k][6][t, 0]) for k in IDS})
    with pytest.raises(RuntimeError, match="minibatch_size == horizon"):
        pol.learn(32, 0.95, 0.95, 0.2, 2, 0.01, 10.0)
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    pol.learn(64, 0.95, 0.95, 0.2, 2, 0.01, 10.0, permutations=perms, order=g["order"])
    m = pol.last_metrics.cpu().numpy()
    np.testing.assert_allclose(m[:, :2], g["losses"], rtol=1e-4, atol=1e-5)
    assert not np.allclose(pol.last_factor.cpu().numpy(), 1.0)                     # the sequential factor did move
    tol = dict(rtol=1e-4, atol=1e-5)
    for k in IDS:
        assert_module_close(pol.agents[k].actor, maddpg_nets(g, "final", "actor")[k], "final actor " + k, tol)
        assert_module_close(pol.agents[k].critic, maddpg_nets(g, "final", "critic")[k], "final critic " + k, tol)


def test_happo_discrete_emulated(golden, emul):
    _happo_discrete(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_happo_discrete_gpu(golden):
    _happo_discrete(golden, torch.device("cuda"))


def test_happo_emulated(golden, emul):
    _happo(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_happo_gpu(golden):
    _happo(golden, torch.device("cuda"))


def test_lr_decay_of_the_mappo_family_emulated(emul):
    """IPPO.py:324-331 / HAPPO.py:460-467 decay both learning rates linearly (the next learn reads them from the agents);
    MAPPO.py:484-491 cannot run upstream (its merged-optimiser Agent has no actor_optimizer) and raises here too."""
    from freerl_b200.HAPPO import HAPPO
    from freerl_b200.IPPO import IPPO
    from freerl_b200.MAPPO import MAPPO
    dim_info, dev = {k: [18, 5] for k in IDS}, torch.device("cpu")
    for cls in (IPPO, HAPPO):
        pol = cls(dim_info, True, 1e-3, 5e-4, 64, dev, dict(MAPPO_TRICK))
        pol.lr_decay(25, 100)
        for ag in pol.agents.values():
            assert ag.lr == 1e-3 * 0.75 and ag.lr_critic == 5e-4 * 0.75
    with pytest.raises(AttributeError, match="actor_optimizer"):
        MAPPO(dim_info, True, 1e-3, 5e-4, 64, dev, dict(MAPPO_TRICK)).lr_decay(25, 100)
