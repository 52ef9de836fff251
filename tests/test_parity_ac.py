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


def EXPECT_PATH():
    import os
    return 0 if os.environ.get("FREERL_B200_AC_PATH") == "generic" else 1


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
        assert pol.last_path == EXPECT_PATH()      # the small-batch schedule (csrc/algo_acfx.cuh) unless the env forces the generic kernel
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5, (m[0], r["critic_loss"])
        assert _rel(m[1], r["actor_loss"]) < 1e-5, (m[1], r["actor_loss"])
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
        assert pol.last_path == EXPECT_PATH()
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5
        if "actor_loss" in r:
            assert _rel(m[1], r["actor_loss"]) < 1e-5
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
        assert pol.last_path == EXPECT_PATH()
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5
        assert _rel(m[1], r["actor_loss"]) < 1e-5
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)


def _bon(golden, device, alg):
    """Batch_ObsNorm inside the fused kernel (running statistics over batch means, SAC.py:390-421) vs oracle + golden.
    The statistics are BIT-IDENTICAL to torch's (the kernel reproduces the association of torch's CPU mean(dim=0)); that
    matters because the normalisation divides by std = sqrt(S/n) of batch-mean differences (a different summation order
    moves the n = 2 losses by 1e-3)."""
    from freerl_b200.DDPG import DDPG
    from freerl_b200.SAC import SAC
    g = golden(alg + "_bon")
    bon = algos.BatchObsNorm(17)
    if alg == "sac":
        trick = {"ObsNorm": False, "Batch_ObsNorm": True, "OUNoise": True, "GaussNoise": False}
        pol = SAC([17, 6], True, 1e-3, 1e-3, 1000, device, trick=trick)
        orc = algos.SACOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, act_dim=6, obs_norm=bon)
    else:
        sup = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": True}
        pol = DDPG([17, 6], True, 1e-3, 1e-3, 1000, device, trick=None, supplement=sup)
        orc = algos.DDPGOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, weight_decay=True, obs_norm=bon)
    _load(pol, g)
    idxs = fill_buffer_from_batches(pol.buffer, g, 4)
    tol = dict(rtol=2e-5, atol=2e-6)
    for it in range(4):
        if alg == "sac":
            n0, n1 = g["noise/%d/0" % it], g["noise/%d/1" % it]
            r = orc.learn(golden_batch(g, it), torch.from_numpy(n0), torch.from_numpy(n1), 0.99, 0.01)
            pol.learn(64, 0.99, 0.01, indices=idxs[it][None], noise_next=n0[None], noise_new=n1[None])
        else:
            r = orc.learn(golden_batch(g, it), 0.99, 0.01)
            pol.learn(64, 0.99, 0.01, indices=idxs[it][None])
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5, (it, m[0], r["critic_loss"])
        assert _rel(m[1], r["actor_loss"]) < 1e-5, (it, m[1], r["actor_loss"])
        ms = pol.batch_size_obs_norm.running_ms
        assert ms.n == it + 1
        np.testing.assert_array_equal(ms.mean.cpu().numpy(), bon.mean.numpy())
        np.testing.assert_array_equal(ms.std.cpu().numpy(), bon.std.numpy())
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it), tol)
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n, tol)
    np.testing.assert_array_equal(pol.batch_size_obs_norm.running_ms.std.cpu().numpy(), g["final/norm/std"])
    if alg == "ddpg":
        np.testing.assert_allclose(pol.select_action(g["act/obs"]), g["act/action"], rtol=1e-5, atol=2e-6)


def _bon_sizes(device):
    """Ragged batch sizes (lane remainders, < 16 rows, > 256 rows) and both summation paths of torch's outer sum (obs 19:
    interleaved `row_sum` lanes; obs 6: columns 0-3 plain `multi_row_sum`, 4-5 `row_sum` — the same split under AVX2 and
    AVX-512 builds of torch): running statistics bit-identical to torch's."""
    from freerl_b200.DDPG import DDPG
    sup = {"weight_decay": False, "OUNoise": False, "ObsNorm": False, "net_init": False, "Batch_ObsNorm": True}
    rng = np.random.default_rng(5)
    for B, od in ((7, 19), (100, 19), (300, 19), (7, 6), (37, 6), (300, 6)):
        torch.manual_seed(B)
        pol = DDPG([od, 3], True, 1e-3, 1e-3, 512, device, trick=None, supplement=sup)
        obs = (rng.standard_normal((400, od)) + rng.uniform(1, 3, od)).astype(np.float32)
        pol.add(obs, rng.uniform(-1, 1, (400, 3)).astype(np.float32), rng.standard_normal(400).astype(np.float32),
                obs[::-1].copy(), rng.random(400) < 0.1)
        bon = algos.BatchObsNorm(od)
        for it in range(3):
            idx = rng.permutation(400)[:B]
            bon(torch.from_numpy(obs[idx]))
            pol.learn(B, 0.99, 0.01, indices=idx[None])
            ms = pol.batch_size_obs_norm.running_ms
            np.testing.assert_array_equal(ms.mean.cpu().numpy(), bon.mean.numpy(), err_msg="B=%d it=%d" % (B, it))
            np.testing.assert_array_equal(ms.std.cpu().numpy(), bon.std.numpy(), err_msg="B=%d it=%d" % (B, it))
            assert np.isfinite(pol.last_metrics.cpu().numpy()).all()


def test_bon_sizes_emulated(emul):
    _bon_sizes(torch.device("cpu"))


@pytest.mark.gpu
def test_bon_sizes_gpu():
    _bon_sizes(torch.device("cuda"))


def test_sac_bon_emulated(golden, emul):
    _bon(golden, torch.device("cpu"), "sac")


def test_ddpg_bon_emulated(golden, emul):
    _bon(golden, torch.device("cpu"), "ddpg")


@pytest.mark.gpu
def test_sac_bon_gpu(golden):
    _bon(golden, torch.device("cuda"), "sac")


@pytest.mark.gpu
def test_ddpg_bon_gpu(golden):
    _bon(golden, torch.device("cuda"), "ddpg")


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


# ---- the generic kernel (csrc/algo_ac.cuh) stays covered: same checks with the small-batch schedule switched off ----
def test_sac_generic_kernel_emulated(golden, emul, monkeypatch):
    monkeypatch.setenv("FREERL_B200_AC_PATH", "generic")
    _sac(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_sac_td3_generic_kernel_gpu(golden, monkeypatch):
    monkeypatch.setenv("FREERL_B200_AC_PATH", "generic")
    _sac(golden, torch.device("cuda"))
    _td3(golden, torch.device("cuda"))


# ---- one launch of K fused learns == K launches of one learn, BIT FOR BIT (same indices / noise): a stale-weights or
#      stale-exchange bug between the fused updates of the persistent kernel would show up here ----
def _fused_equals_sequential(device, alg, B, K):
    from freerl_b200.SAC import SAC
    from freerl_b200.TD3 import TD3
    rng = np.random.default_rng(11)
    n = 2048
    obs, nobs = rng.standard_normal((n, 17), dtype=np.float32), rng.standard_normal((n, 17), dtype=np.float32)
    act, rew = rng.uniform(-1, 1, (n, 6)).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    done = rng.random(n) < 0.05
    idx = np.stack([rng.permutation(n)[:B] for _ in range(K)])
    nz0 = rng.standard_normal((K, B, 6)).astype(np.float32)
    nz1 = rng.standard_normal((K, B, 6)).astype(np.float32)
    pols = []
    for fused in (True, False):
        torch.manual_seed(3)
        if alg == "sac":
            pol = SAC([17, 6], True, 1e-3, 1e-3, 4096, device, trick={})
        else:
            pol = TD3([17, 6], True, 1e-3, 1e-3, 4096, device, trick=None,
                      realize={"clip_double": True, "policy_noise": True, "twin_delay": True})
        pol.add(obs, act, rew, nobs, done)
        metrics = []
        for k0 in ([0] if fused else range(K)):
            sl = slice(0, K) if fused else slice(k0, k0 + 1)
            nu = K if fused else 1
            if alg == "sac":
                pol.learn(B, 0.99, 0.01, indices=idx[sl], noise_next=nz0[sl], noise_new=nz1[sl], n_updates=nu)
            else:
                pol.learn(B, 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0, indices=idx[sl], noise=nz0[sl], n_updates=nu)
            assert pol.last_path == EXPECT_PATH()
            metrics.append(pol.last_metrics.cpu().numpy().copy())
        pols.append((pol, np.concatenate(metrics)))
    (pa, ma), (pb, mb) = pols
    cols = [0, 4] if alg == "td3" else [0, 1, 2, 4, 5, 6]      # TD3 off-steps leave the actor columns unwritten
    np.testing.assert_array_equal(ma[:, cols], mb[:, cols])
    for nname in NETS:
        for (k, va), (_, vb) in zip(getattr(pa.agent, nname).state_dict().items(), getattr(pb.agent, nname).state_dict().items()):
            np.testing.assert_array_equal(va.cpu().numpy(), vb.cpu().numpy(), err_msg="%s.%s" % (nname, k))
    if alg == "sac":
        assert float(pa.alphas.log_alpha) == float(pb.alphas.log_alpha)


def test_fused_equals_sequential_emulated(emul):
    _fused_equals_sequential(torch.device("cpu"), "sac", 40, 3)
    _fused_equals_sequential(torch.device("cpu"), "td3", 40, 4)


@pytest.mark.gpu
def test_fused_equals_sequential_gpu():
    _fused_equals_sequential(torch.device("cuda"), "sac", 256, 8)
    _fused_equals_sequential(torch.device("cuda"), "td3", 256, 8)


# ---- the mode the bench times (mode="fast": device sampler + in-kernel Philox noise, K learns per launch) against the oracle:
#      the indices the launch sampled are read back, the noise it drew is regenerated by the library's audit hook
#      (frl_debug_randn: the same randn call, streams 1 = eps of a', 2 = eps of the new action, counter = learn number) and both
#      are handed to the oracle — B = 256, 8 fused updates, the bench's batch shape ----
def _fast_mode_audit(device, B=256, K=8, n=4096):
    import ctypes
    from collections import OrderedDict
    from freerl_b200 import _lib
    from freerl_b200.SAC import SAC
    from freerl_b200 import _common
    torch.manual_seed(21)
    np.random.seed(21)
    _common._fast_seed_counter[0] = 0      # the device generators' seed counts the policies built in this process: pin it, so that the
    #                                        draws (and the margins below) do not depend on which tests ran before this one
    pol = SAC([17, 6], True, 1e-3, 1e-3, n, device, trick={}, mode="fast")
    rng = np.random.default_rng(8)
    obs, act = rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    rew, nobs = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32)
    done = rng.random(n) < 0.05
    pol.add(obs, act, rew, nobs, done)
    sd = lambda m: OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())
    orc = algos.SACOracle(sd(pol.agent.actor), sd(pol.agent.critic), 1e-3, 1e-3, act_dim=6)
    seed, it0 = pol._seed, pol._n_learn
    pol.learn(B, 0.99, 0.01, n_updates=K)
    idx = pol._keepalive[0].cpu().numpy()
    assert idx.shape == (K, B) and pol._keepalive[1] is None and pol._keepalive[2] is None       # nothing was injected
    m = pol.last_metrics.cpu().numpy()
    for u in range(K):
        assert len(set(idx[u].tolist())) == B and idx[u].min() >= 0 and idx[u].max() < n       # without replacement, in range
        nz = []
        for stream in (1, 2):
            t = torch.empty(B * 6, dtype=torch.float32, device=device)
            _lib.check(_lib.lib().frl_debug_randn(ctypes.c_uint64(seed), stream, it0 + u, B * 6, _lib.ptr(t), _lib.stream_ptr(device)), "frl_debug_randn")
            nz.append(t.cpu().reshape(B, 6))
        i = idx[u]
        batch = (torch.from_numpy(obs[i]), torch.from_numpy(act[i]), torch.from_numpy(rew[i]).reshape(-1, 1), torch.from_numpy(nobs[i]),
                 torch.from_numpy(done[i].astype(np.float32)).reshape(-1, 1))
        r = orc.learn(batch, nz[0], nz[1], 0.99, 0.01)
        assert _rel(m[u, 0], r["critic_loss"]) < 1e-5, (u, m[u, 0], r["critic_loss"])
        # the actor loss is a mean of signed O(1) terms (-Q - alpha * entropy) that sits near -0.17 here, and this is a FREE-running chain
        # (update u starts from the kernel's own parameters after u fused updates): allclose form.  Measured on B200 over seeds: <= 2e-7
        # relative on the first update, 3.4e-6 absolute at worst by update 2 (tools/parity_margin.py, profiles/r4_parity_margin.txt)
        np.testing.assert_allclose(m[u, 1], r["actor_loss"], rtol=2e-5, atol=2e-6 if u == 0 else 5e-6, err_msg="actor loss, update %d" % u)
    z = torch.cat(nz).numpy()                                  # the regenerated draws are standard normals
    assert abs(z.mean()) < 0.08 and abs(z.std() - 1.0) < 0.05
    for name in NETS:
        assert_module_close(getattr(pol.agent, name), getattr(orc, name), "fast mode, %s after %d fused learns" % (name, K))
    assert _rel(float(pol.alphas.log_alpha), orc.log_alpha.item()) < 1e-5


def test_fast_mode_audit_emulated(emul):
    _fast_mode_audit(torch.device("cpu"), B=64, K=3, n=512)


@pytest.mark.gpu
def test_fast_mode_audit_gpu():
    _fast_mode_audit(torch.device("cuda"))


# ---- the bench's batch shape against the REFERENCE: B = 256, ten chained learns (fixture sac_b256.npz, generated from the unmodified
#      SAC_file/SAC.py by oracle/make_golden.py::gen_sac_b256); losses of every learn vs the reference's own numbers ----
def _sac_b256(golden, device):
    from freerl_b200.SAC import SAC
    g = golden("sac_b256")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    pol = SAC([17, 6], True, 1e-3, 1e-3, 4096, device, trick=trick)
    _load(pol, g)
    orc = algos.SACOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, act_dim=6)
    idxs = fill_buffer_from_batches(pol.buffer, g, 10)
    for it in range(10):
        n0, n1 = g["noise/%d/0" % it], g["noise/%d/1" % it]
        r = orc.learn(golden_batch(g, it), torch.from_numpy(n0), torch.from_numpy(n1), 0.99, 0.01)
        pol.learn(256, 0.99, 0.01, indices=idxs[it][None], noise_next=n0[None], noise_new=n1[None])
        m = pol.last_metrics[0].cpu().numpy()
        ref_c, ref_a = float(g["loss/%03d/update_critic" % (2 * it)][0]), float(g["loss/%03d/update_actor" % (2 * it + 1)][0])
        assert _rel(r["critic_loss"], ref_c) < 1e-6 and _rel(r["actor_loss"], ref_a) < 2e-6      # the oracle IS the reference here
        assert _rel(m[0], ref_c) < 1e-5, (it, m[0], ref_c)
        assert _rel(m[1], ref_a) < 1e-5, (it, m[1], ref_a)
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)
    assert _rel(float(pol.alphas.log_alpha), float(g["final/log_alpha"])) < 1e-5


def test_sac_b256_k10_vs_reference_emulated(golden, emul):
    _sac_b256(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_sac_b256_k10_vs_reference_gpu(golden):
    _sac_b256(golden, torch.device("cuda"))


# ---- k = 100 learns at B = 256, TEACHER-FORCED: every learn starts from the oracle's state (parameters, targets, Adam moments, step
#      counts, temperature), so the per-step agreement is measured over a long trajectory without the drift of a free-running chain
#      (Adam divides by sqrt(v) + eps: elements whose gradient is ~eps amplify last-bit differences, see parity_util) ----
def _push_sac(pol, orc):
    from parity_util import push_block_state
    ag = pol.agent
    push_block_state(ag._actor, ag.actor, orc.actor, orc.opt_a.m, orc.opt_a.v)
    push_block_state(ag._critic, ag.critic, orc.critic, orc.opt_c.m, orc.opt_c.v)
    push_block_state(ag._actor_t, ag.actor_target, orc.actor_target)
    push_block_state(ag._critic_t, ag.critic_target, orc.critic_target)
    ag.actor_step, ag.critic_step = orc.opt_a.step, orc.opt_c.step
    st = pol.alphas.state
    with torch.no_grad():
        st[0], st[1], st[2] = float(orc.log_alpha), float(orc.opt_alpha.m[0]), float(orc.opt_alpha.v[0])
    pol.alphas.step = orc.opt_alpha.step


def _sac_k100(device, K=100, B=256, n=4096):
    from collections import OrderedDict
    from freerl_b200.SAC import SAC
    torch.manual_seed(33)
    np.random.seed(33)
    pol = SAC([17, 6], True, 1e-3, 1e-3, n, device, trick={})
    rng = np.random.default_rng(17)
    obs, act = rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    rew, nobs = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32)
    done = rng.random(n) < 0.05
    pol.add(obs, act, rew, nobs, done)
    sd = lambda m: OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())
    orc = algos.SACOracle(sd(pol.agent.actor), sd(pol.agent.critic), 1e-3, 1e-3, act_dim=6)
    worst_c = worst_a = 0.0
    for u in range(K):
        _push_sac(pol, orc)
        i = rng.choice(n, B, replace=False)
        z0, z1 = rng.standard_normal((B, 6)).astype(np.float32), rng.standard_normal((B, 6)).astype(np.float32)
        batch = (torch.from_numpy(obs[i]), torch.from_numpy(act[i]), torch.from_numpy(rew[i]).reshape(-1, 1), torch.from_numpy(nobs[i]),
                 torch.from_numpy(done[i].astype(np.float32)).reshape(-1, 1))
        r = orc.learn(batch, torch.from_numpy(z0), torch.from_numpy(z1), 0.99, 0.01)
        pol.learn(B, 0.99, 0.01, indices=i[None], noise_next=z0[None], noise_new=z1[None])
        m = pol.last_metrics[0].cpu().numpy()
        worst_c, worst_a = max(worst_c, _rel(m[0], r["critic_loss"])), max(worst_a, _rel(m[1], r["actor_loss"]))
        if u % 10 == 9 or u == K - 1:
            for name in NETS:
                assert_module_close(getattr(pol.agent, name), getattr(orc, name), "%s after teacher-forced learn %d" % (name, u))
            assert _rel(float(pol.alphas.log_alpha), orc.log_alpha.item()) < 1e-5
    assert worst_c < 1e-5 and worst_a < 1e-5, (worst_c, worst_a)


def test_sac_k100_emulated(emul):
    _sac_k100(torch.device("cpu"), K=40, B=64, n=1024)


@pytest.mark.gpu
def test_sac_k100_gpu():
    _sac_k100(torch.device("cuda"))
