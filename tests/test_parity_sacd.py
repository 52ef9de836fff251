"""Discrete-action SAC — the ``hands_on`` branch of ``SAC_file/SAC_add_discrete.py`` (:137-177, 226-360) — on ``frl_sacd_learn``
(csrc/algo_sacd.cuh) vs the oracle (``oracle.algos.SACDiscreteOracle``) and the reference-generated fixture ``sac_discrete.npz``
(oracle/make_golden.py::gen_sac_discrete: obs 8, 4 actions, B = 64, three learns, eight sampled + greedy actions)."""
import numpy as np
import pytest
import torch

from oracle import algos
from parity_util import assert_module_close, fill_buffer_from_batches, golden_batch, load_into, net_from_golden

NETS = ("actor", "critic", "actor_target", "critic_target")


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def test_oracle_sac_discrete_vs_reference(golden):
    g = golden("sac_discrete")
    orc = algos.SACDiscreteOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, n_actions=4)
    for it in range(3):
        r = orc.learn(golden_batch(g, it), 0.99, 0.01)
        assert _rel(r["critic_loss"], float(g["loss/%03d/update_critic" % (2 * it)][0])) < 1e-6
        assert _rel(r["actor_loss"], float(g["loss/%03d/update_actor" % (2 * it + 1)][0])) < 2e-6
    for n in NETS:
        ref = net_from_golden(g, "final/%s/" % n)
        for k, v in getattr(orc, n).items():
            np.testing.assert_allclose(v.detach().numpy(), ref[k].numpy(), rtol=1e-5, atol=1e-6, err_msg=n + "/" + k)
    assert _rel(orc.log_alpha.item(), float(g["final/log_alpha"])) < 1e-6


def _run(golden, device, inject):
    from freerl_b200.SAC_add_discrete import SAC
    g = golden("sac_discrete")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    pol = SAC([8, 4], False, 1e-3, 1e-3, 1000, device, trick=trick)
    assert type(pol).__name__ == "_DiscreteSAC" and pol.buffer.act_dim == 1
    for n in NETS:
        src = "actor" if n == "actor_target" else ("critic" if n == "critic_target" else n)
        load_into(getattr(pol.agent, n), net_from_golden(g, "init/%s/" % src))
    orc = algos.SACDiscreteOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, n_actions=4)
    idxs = fill_buffer_from_batches(pol.buffer, g, 3)
    for it in range(3):
        r = orc.learn(golden_batch(g, it), 0.99, 0.01)
        pol.learn(64, 0.99, 0.01, indices=idxs[it][None])
        m = pol.last_metrics[0].cpu().numpy()
        assert _rel(m[0], r["critic_loss"]) < 1e-5, (it, m[0], r["critic_loss"])
        assert abs(m[1] - r["actor_loss"]) < 2e-5 * abs(r["actor_loss"]) + 2e-6, (it, m[1], r["actor_loss"])
        assert _rel(m[4], r["critic_gnorm"]) < 1e-4 and _rel(m[5], r["actor_gnorm"]) < 1e-4
        assert _rel(m[0], float(g["loss/%03d/update_critic" % (2 * it)][0])) < 1e-5           # the reference's own numbers
        for n in NETS:
            assert_module_close(getattr(pol.agent, n), getattr(orc, n), "%s after learn %d" % (n, it))
        assert _rel(float(pol.alphas.log_alpha), orc.log_alpha.item()) < 1e-5
    for n in NETS:
        assert_module_close(getattr(pol.agent, n), net_from_golden(g, "final/%s/" % n), "final " + n)
    assert _rel(float(pol.alphas.log_alpha), float(g["final/log_alpha"])) < 1e-5
    # acting: Categorical(probs).sample() (torch.multinomial's exponential trick) and the greedy arg-max, vs the reference's draws
    torch.set_rng_state(torch.from_numpy(g["sel/rng_state"].copy()))
    for i in range(8):
        if inject:
            a = pol.select_action(g["sel/obs"][i], noise=torch.empty((1, 4)).exponential_(1))
        else:
            a = pol.select_action(g["sel/obs"][i])
        assert int(a) == int(g["sel/action"][i]), (i, a, g["sel/action"][i])
        assert int(pol.evaluate_action(g["sel/obs"][i])) == int(g["sel/greedy"][i])
    # fused == sequential
    pol.learn(64, 0.99, 0.01, n_updates=2, indices=np.stack([idxs[0], idxs[1]]))
    assert pol.last_metrics.shape[0] == 2 and bool(torch.isfinite(pol.last_metrics).all())


def test_sac_discrete_emulated(golden, emul):
    _run(golden, torch.device("cpu"), False)


@pytest.mark.gpu
def test_sac_discrete_gpu(golden):
    _run(golden, torch.device("cuda"), True)
