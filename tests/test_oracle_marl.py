"""MADDPG oracle replayed against the reference-generated fixture."""
from collections import OrderedDict

import numpy as np
import torch

from oracle.marl import MADDPGOracle

IDS = ["agent_0", "agent_1", "agent_2"]


def maddpg_nets(g, which, kind):
    out = OrderedDict()
    for k in IDS:
        pre = "%s/%s/%s/" % (which, k, kind)
        out[k] = OrderedDict((n[len(pre):], torch.from_numpy(g[n].copy())) for n in g.files if n.startswith(pre))
    return out


def maddpg_norms():
    from oracle.algos import BatchObsNorm
    return {k: BatchObsNorm(18) for k in IDS}


def maddpg_batch(g, idx):
    f = lambda x: torch.from_numpy(np.asarray(x, dtype=np.float32))
    return {k: (f(g["buf/%s/obs" % k][idx]), f(g["buf/%s/act" % k][idx]), f(g["buf/%s/rew" % k][idx]).reshape(-1, 1),
                f(g["buf/%s/nobs" % k][idx]), f(g["buf/%s/done" % k][idx]).reshape(-1, 1)) for k in IDS}


def test_maddpg_oracle_vs_reference(golden):
    g = golden("maddpg")
    orc = MADDPGOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 1e-3)
    ls = []
    for it in range(2):
        r = orc.learn([maddpg_batch(g, g["idx/%d/%d" % (it, j)]) for j in range(3)], 0.95, 0.01)
        for cl, al in r:
            ls += [cl, al]
    np.testing.assert_allclose(np.array(ls), g["losses"], rtol=2e-5, atol=1e-7)
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic), ("actor_target", orc.actor_target), ("critic_target", orc.critic_target)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6)


def test_maddpg_batch_obs_norm_oracle_vs_reference(golden):
    g = golden("maddpg_bon")
    norms = maddpg_norms()
    orc = MADDPGOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 1e-3, obs_norms=norms)
    ls = []
    for it in range(2):
        r = orc.learn([maddpg_batch(g, g["idx/%d/%d" % (it, j)]) for j in range(3)], 0.95, 0.01)
        for cl, al in r:
            ls += [cl, al]
    np.testing.assert_allclose(np.array(ls), g["losses"], rtol=2e-5, atol=1e-7)
    for k in IDS:
        assert norms[k].n == int(g["final/norm/%s/n" % k]) == 6          # 2 learns x 3 agents' samples
        np.testing.assert_array_equal(norms[k].std.numpy(), g["final/norm/%s/std" % k])
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic), ("actor_target", orc.actor_target)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6)


from oracle.marl import MAPPOOracle  # noqa: E402
from oracle.make_golden_marl import MAPPO_TRICK  # noqa: E402


def mappo_data(g):
    return {k: tuple(torch.from_numpy(g["data/%s/%s" % (k, n)]) for n in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
            for k in IDS}


def test_mappo_oracle_vs_reference(golden):
    g = golden("mappo")
    orc = MAPPOOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, MAPPO_TRICK)
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(mappo_data(g), perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=2e-5, atol=1e-6)
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6)


def test_mappo_discrete_oracle_vs_reference(golden):
    g = golden("mappo_disc")
    orc = MAPPOOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, MAPPO_TRICK, is_continue=False)
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(mappo_data(g), perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=2e-5, atol=1e-6)
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6)


def _final_close(orc, g):
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic), ("actor_target", orc.actor_target), ("critic_target", orc.critic_target)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6, err_msg="%s %s %s" % (k, kind, n))


def test_maddpg_simple_oracle_vs_reference(golden):
    """MADDPG_simple.py = MADDPG without supplements (no weight decay, default init)"""
    g = golden("maddpg_simple")
    orc = MADDPGOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 1e-3, weight_decay=False)
    ls = []
    for it in range(2):
        r = orc.learn([maddpg_batch(g, g["idx/%d/%d" % (it, j)]) for j in range(3)], 0.95, 0.01)
        for cl, al in r:
            ls += [cl, al]
    np.testing.assert_allclose(np.array(ls), g["losses"][:, 1], rtol=2e-5, atol=1e-7)
    _final_close(orc, g)


def matd3_noises(g, it):
    return [[torch.from_numpy(g["noise/%d/%d/%d" % (it, i, j)]) for j in range(3)] for i in range(3)]


def test_matd3_oracle_vs_reference(golden):
    from oracle.marl import MATD3Oracle
    g = golden("matd3")
    orc = MATD3Oracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 1e-3)
    ls = []
    for it in range(3):
        r = orc.learn([maddpg_batch(g, g["idx/%d/%d" % (it, j)]) for j in range(3)], matd3_noises(g, it), 0.95, 0.01, 1.0, 0.1, 0.5, 1.0, 2)
        for cl, al in r:
            ls += [cl] + ([al] if al is not None else [])
    assert len(ls) == g["losses"].shape[0] == 12                      # 9 critic + 3 actor updates (policy_freq 2)
    np.testing.assert_allclose(np.array(ls), g["losses"][:, 1], rtol=2e-5, atol=1e-7)
    _final_close(orc, g)


IPPO_TRICK = {'adv_norm': True, 'ObsNorm': True, 'reward_norm': False, 'reward_scaling': True, 'orthogonal_init': True,
              'adam_eps': True, 'lr_decay': False, 'ValueClip': True, 'huber_loss': True, 'LayerNorm': True, 'feature_norm': True}


def ippo_data(g):
    return {k: tuple(torch.from_numpy(g["data/%s/%s" % (k, n)]) for n in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done")) for k in IDS}


def _ippo_oracle(golden, name, is_continue):
    from oracle.marl import IPPOOracle
    g = golden(name)
    orc = IPPOOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 5e-4, IPPO_TRICK, is_continue)
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(ippo_data(g), perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=3e-5, atol=1e-6)
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6, err_msg="%s %s %s" % (k, kind, n))


def test_ippo_continuous_oracle_vs_reference(golden):
    _ippo_oracle(golden, "ippo_cont", True)


def test_ippo_discrete_oracle_vs_reference(golden):
    _ippo_oracle(golden, "ippo_disc", False)


def test_happo_oracle_vs_reference(golden):
    from oracle.marl import HAPPOOracle
    g = golden("happo")
    assert list(g["order"]) != [0, 1, 2]                       # the fixture exercises a non-trivial agent order
    orc = HAPPOOracle(maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic"), 1e-3, 5e-4, MAPPO_TRICK)
    perms = {k: [g["perm/%s/%d" % (k, e)] for e in range(2)] for k in IDS}
    r = orc.learn(mappo_data(g), g["order"], perms, 32, 0.95, 0.95, 0.2, 0.01, 10.0)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=3e-5, atol=1e-6)
    for k in IDS:
        for kind, nets in (("actor", orc.actor), ("critic", orc.critic)):
            for n, v in nets[k].items():
                np.testing.assert_allclose(v.detach().numpy(), g["final/%s/%s/%s" % (k, kind, n)], rtol=2e-5, atol=2e-6, err_msg="%s %s %s" % (k, kind, n))
