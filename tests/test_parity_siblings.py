"""Sibling scripts on the same fused kernels: DDPG_simple, MADDPG_simple, MATD3_simple vs their oracles and
the fixtures generated from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import algos
from oracle.marl import MADDPGOracle, MATD3Oracle
from parity_util import assert_module_close, fill_buffer_from_batches, golden_batch, load_into, net_from_golden
from test_oracle_marl import IDS, maddpg_batch, maddpg_nets, matd3_noises

NETS = ("actor", "critic", "actor_target", "critic_target")


def _ddpg_simple(golden, device):
    from freerl_b200.DDPG_simple import DDPG
    g = golden("ddpg_simple")
    pol = DDPG([17, 6], True, 1e-3, 1e-3, 1000, device, trick=None)
    for n in NETS:
        load_into(getattr(pol.agent, n), net_from_golden(g, "init/%s/" % n.replace("_target", "")))
    orc = algos.DDPGOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, weight_decay=False)
    idxs = fill_buffer_from_batches(pol.buffer, g, 3)
    for it in range(3):
        r = orc.learn(golden_batch(g, it), 0.99, 0.01)
        pol.learn(64, 0.99, 0.01, indices=idxs[it][None])
        m = pol.last_metrics[0].cpu().numpy()
        assert abs(m[0] - r["critic_loss"]) <= 1e-5 * abs(r["critic_loss"])
        assert abs(m[1] - r["actor_loss"]) <= 2e-5 * abs(r["actor_loss"])
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)


def _setup_ma(pol, g):
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k]); load_into(pol.agents[k].actor_target, ia[k])
        load_into(pol.agents[k].critic, ic[k]); load_into(pol.agents[k].critic_target, ic[k])
        pol.buffers[k].add(g["buf/%s/obs" % k], g["buf/%s/act" % k], g["buf/%s/rew" % k], g["buf/%s/nobs" % k], g["buf/%s/done" % k])
    return ia, ic


def _final_ma(pol, g):
    for k in IDS:
        for kind in NETS:
            assert_module_close(getattr(pol.agents[k], kind), maddpg_nets(g, "final", kind)[k], "final %s %s" % (kind, k))
    acts = pol.select_action({k: g["act/%s/obs" % k] for k in IDS})
    for k in IDS:
        np.testing.assert_allclose(acts[k], g["act/%s/action" % k], rtol=1e-5, atol=2e-6)


def _maddpg_simple(golden, device):
    from freerl_b200.MADDPG_simple import MADDPG
    g = golden("maddpg_simple")
    pol = MADDPG({k: [18, 5] for k in IDS}, True, 1e-3, 1e-3, 1000, device, trick=None)
    ia, ic = _setup_ma(pol, g)
    orc = MADDPGOracle(ia, ic, 1e-3, 1e-3, weight_decay=False)
    for it in range(2):
        idxs = [g["idx/%d/%d" % (it, j)] for j in range(3)]
        r = orc.learn([maddpg_batch(g, ix) for ix in idxs], 0.95, 0.01)
        pol.learn(64, 0.95, 0.01, indices=idxs)
        m = pol.last_metrics.cpu().numpy()
        for j in range(3):
            assert abs(m[j, 0] - r[j][0]) <= 1e-5 * abs(r[j][0]), (it, j, m[j, 0], r[j][0])
            assert abs(m[j, 1] - r[j][1]) <= 3e-5 * abs(r[j][1]) + 1e-8, (it, j, m[j, 1], r[j][1])
        for k in IDS:
            assert_module_close(pol.agents[k].actor, orc.actor[k], "actor %s" % k)
            assert_module_close(pol.agents[k].critic, orc.critic[k], "critic %s" % k)
    _final_ma(pol, g)


def _matd3(golden, device):
    from freerl_b200.MATD3_simple import MATD3
    g = golden("matd3")
    realize = {'clip_double': True, 'policy_noise': True, 'twin_delay': True}
    pol = MATD3({k: [18, 5] for k in IDS}, True, 1e-3, 1e-3, 1000, device, trick=None, realize=realize)
    assert list(pol.agents["agent_0"].critic.state_dict().keys())[-1] == "l6.bias"           # Critic_TD3 schema
    ia, ic = _setup_ma(pol, g)
    orc = MATD3Oracle(ia, ic, 1e-3, 1e-3)
    ref = g["losses"]
    pos = 0
    for it in range(3):
        idxs = [g["idx/%d/%d" % (it, j)] for j in range(3)]
        nz = matd3_noises(g, it)
        r = orc.learn([maddpg_batch(g, ix) for ix in idxs], nz, 0.95, 0.01, 1.0, 0.1, 0.5, 1.0, 2)
        pol.learn(64, 0.95, 0.01, 1.0, 0.1, 0.5, 1.0, 2, indices=idxs, noise=[[x.numpy() for x in row] for row in nz])
        m = pol.last_metrics.cpu().numpy()
        for j in range(3):
            assert abs(m[j, 0] - r[j][0]) <= 1e-5 * abs(r[j][0]), (it, j, m[j, 0], r[j][0])
            assert ref[pos, 0] == 0 and abs(m[j, 0] - ref[pos, 1]) <= 2e-5 * abs(ref[pos, 1])
            pos += 1
            if r[j][1] is not None:
                assert abs(m[j, 1] - r[j][1]) <= 3e-5 * abs(r[j][1]) + 1e-8, (it, j, m[j, 1], r[j][1])
                assert ref[pos, 0] == 1 and abs(m[j, 1] - ref[pos, 1]) <= 5e-5 * abs(ref[pos, 1]) + 1e-8
                pos += 1
        for k in IDS:
            for kind in NETS:
                assert_module_close(getattr(pol.agents[k], kind), getattr(orc, kind)[k], "%s %s after learn %d" % (kind, k, it))
    assert pos == 12 and pol.total_it == 3
    assert [pol.agents[k].actor_step for k in IDS] == [1, 1, 1] and [pol.agents[k].critic_step for k in IDS] == [3, 3, 3]
    _final_ma(pol, g)


CASES = {"ddpg_simple": _ddpg_simple, "maddpg_simple": _maddpg_simple, "matd3": _matd3}


@pytest.mark.parametrize("name", sorted(CASES))
def test_sibling_emulated(golden, emul, name):
    CASES[name](golden, torch.device("cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_sibling_gpu(golden, name):
    CASES[name](golden, torch.device("cuda"))
