"""freerl_b200.MADDPG (multi-agent path of the fused actor-critic kernel) vs the oracle and the reference golden."""
import numpy as np
import pytest
import torch

from oracle.marl import MADDPGOracle
from parity_util import assert_module_close, load_into
from test_oracle_marl import IDS, maddpg_batch, maddpg_nets, maddpg_norms

SUP = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": False}


def _run(golden, device, bon=False):
    from freerl_b200.MADDPG import MADDPG
    g = golden("maddpg_bon" if bon else "maddpg")
    dim_info = {k: [18, 5] for k in IDS}
    pol = MADDPG(dim_info, True, 1e-3, 1e-3, 1000, device, trick=None, supplement=dict(SUP, Batch_ObsNorm=bon))
    norms = maddpg_norms() if bon else None
    ia, ic = maddpg_nets(g, "init", "actor"), maddpg_nets(g, "init", "critic")
    for k in IDS:
        load_into(pol.agents[k].actor, ia[k]); load_into(pol.agents[k].actor_target, ia[k])
        load_into(pol.agents[k].critic, ic[k]); load_into(pol.agents[k].critic_target, ic[k])
        pol.buffers[k].add(g["buf/%s/obs" % k], g["buf/%s/act" % k], g["buf/%s/rew" % k], g["buf/%s/nobs" % k], g["buf/%s/done" % k])
    orc = MADDPGOracle(ia, ic, 1e-3, 1e-3, obs_norms=norms)
    ref_losses = g["losses"].reshape(2, 3, 2)
    for it in range(2):
        idxs = [g["idx/%d/%d" % (it, j)] for j in range(3)]
        r = orc.learn([maddpg_batch(g, ix) for ix in idxs], 0.95, 0.01)
        pol.learn(64, 0.95, 0.01, indices=idxs)
        m = pol.last_metrics.cpu().numpy()
        for j in range(3):
            assert abs(m[j, 0] - r[j][0]) <= 1e-5 * abs(r[j][0]), (it, j, m[j, 0], r[j][0])
            assert abs(m[j, 1] - r[j][1]) <= 3e-5 * abs(r[j][1]) + 1e-8, (it, j, m[j, 1], r[j][1])
            assert abs(m[j, 0] - ref_losses[it, j, 0]) <= 2e-5 * abs(ref_losses[it, j, 0])
        for k in IDS:
            assert_module_close(pol.agents[k].actor, orc.actor[k], "actor %s" % k)
            assert_module_close(pol.agents[k].critic, orc.critic[k], "critic %s" % k)
            assert_module_close(pol.agents[k].actor_target, orc.actor_target[k], "actor_target %s" % k)
    for k in IDS:
        for kind in ("actor", "critic", "actor_target", "critic_target"):
            assert_module_close(getattr(pol.agents[k], kind), maddpg_nets(g, "final", kind)[k], "final %s %s" % (kind, k))
    acts = pol.select_action({k: g["buf/%s/obs" % k][0].astype(np.float32) for k in IDS})
    assert acts["agent_0"].shape == (5,)
    if bon:      # per-agent statistics (bit-identical) and the normalised select_action of the reference
        acts = pol.select_action({k: g["act/%s/obs" % k] for k in IDS})
        for k in IDS:
            ms = pol.batch_size_obs_norm[k].running_ms
            assert ms.n == 6
            np.testing.assert_array_equal(ms.mean.cpu().numpy(), g["final/norm/%s/mean" % k])
            np.testing.assert_array_equal(ms.std.cpu().numpy(), g["final/norm/%s/std" % k])
            np.testing.assert_allclose(acts[k], g["act/%s/action" % k], rtol=1e-5, atol=2e-6)


def test_maddpg_emulated(golden, emul):
    _run(golden, torch.device("cpu"))


def test_maddpg_batch_obs_norm_emulated(golden, emul):
    _run(golden, torch.device("cpu"), bon=True)


@pytest.mark.gpu
def test_maddpg_batch_obs_norm_gpu(golden):
    _run(golden, torch.device("cuda"), bon=True)


@pytest.mark.gpu
def test_maddpg_gpu(golden):
    _run(golden, torch.device("cuda"))
