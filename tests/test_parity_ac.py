"""freerl_b200 SAC / TD3 / DDPG (fused actor-critic kernel) vs the oracle and the reference-generated goldens."""
import numpy as np
import pytest
import torch

from oracle import algos
from parity_util import (assert_module_close, fill_buffer_from_batches, golden_batch, load_into, net_from_golden)

NETS = ("actor", "critic", "actor_target", "critic_target")


def _load(pol, g):
    for n in NETS:
        src = "actor" if n == "actor_target" else ("critic" if n == "critic_target" else n)
        load_into(getattr(pol.agent, n), net_from_golden(g, "init/%s/" % src))


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def _sac(golden, device):
    from freerl_b200.SAC import SAC
    g = golden("sac")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    pol = SAC([17, 6], True, 1e-3, 1e-3, 1000, device, trick=trick)
    _load(pol, g)
    orc = algos.SACOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, act_dim=6)
    idxs = fill_buffer_from_batches(pol.buffer, g, 3)
    for it in range(3):
        n0, n1 = g["noise/%d/0" % it], g["noise/%d/1" % it]
        r = orc.learn(golden_batch(g, it), torch.from_numpy(n0), torch.from_numpy(n1), 0.99, 0.01)
        pol.learn(64, 0.99, 0.01, indices=idxs[it][None], noise_next=n0[None], noise_new=n1[None])
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5, (m[0], r["critic_loss"])
        assert _rel(m[1], r["actor_loss"]) < 2e-5, (m[1], r["actor_loss"])
        assert _rel(m[4], r["critic_gnorm"]) < 1e-4 and _rel(m[5], r["actor_gnorm"]) < 1e-4
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
        assert _rel(float(pol.alphas.log_alpha), orc.log_alpha.item()) < 1e-5
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)
    assert _rel(float(pol.alphas.log_alpha), float(g["final/log_alpha"])) < 1e-5


def _td3(golden, device):
    from freerl_b200.TD3 import TD3
    g = golden("td3")
    realize = {"clip_double": True, "policy_noise": True, "twin_delay": True}
    pol = TD3([17, 6], True, 1e-3, 1e-3, 1000, device, trick=None, realize=realize)
    _load(pol, g)
    orc = algos.TD3Oracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3)
    idxs = fill_buffer_from_batches(pol.buffer, g, 4)
    for it in range(4):
        nz = g["noise/%d/0" % it]
        r = orc.learn(golden_batch(g, it), torch.from_numpy(nz), 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0)
        pol.learn(64, 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0, indices=idxs[it][None], noise=nz[None])
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5
        if "actor_loss" in r:
            assert _rel(m[1], r["actor_loss"]) < 2e-5
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)


def _ddpg(golden, device):
    from freerl_b200.DDPG import DDPG
    g = golden("ddpg")
    sup = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": False}
    pol = DDPG([17, 6], True, 1e-3, 1e-3, 1000, device, trick=None, supplement=sup)
    _load(pol, g)
    orc = algos.DDPGOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, weight_decay=True)
    idxs = fill_buffer_from_batches(pol.buffer, g, 3)
    for it in range(3):
        r = orc.learn(golden_batch(g, it), 0.99, 0.01)
        pol.learn(64, 0.99, 0.01, indices=idxs[it][None])
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5
        assert _rel(m[1], r["actor_loss"]) < 2e-5
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)


def test_sac_emulated(golden, emul):
    _sac(golden, torch.device("cpu"))


def test_td3_emulated(golden, emul):
    _td3(golden, torch.device("cpu"))


def test_ddpg_emulated(golden, emul):
    _ddpg(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_sac_gpu(golden):
    _sac(golden, torch.device("cuda"))


@pytest.mark.gpu
def test_td3_gpu(golden):
    _td3(golden, torch.device("cuda"))


@pytest.mark.gpu
def test_ddpg_gpu(golden):
    _ddpg(golden, torch.device("cuda"))
