"""freerl_b200.DQN (fused kernel) vs the oracle and the reference-generated golden fixture."""
import numpy as np
import pytest
import torch

from oracle import algos
from parity_util import (assert_module_close, fill_buffer_from_batches, golden_batch, load_into, net_from_golden)


def _run(golden, device):
    from freerl_b200.DQN import DQN
    g = golden("dqn")
    init = net_from_golden(g, "init/q/")
    pol = DQN([4, 2], False, 1e-3, 1000, device)
    load_into(pol.agent.Qnet, init)
    load_into(pol.agent.Qnet_target, init)
    orc = algos.DQNOracle(init, 1e-3)
    idxs = fill_buffer_from_batches(pol.buffer, g, 3)
    ref_losses = [g[k][0] for k in sorted(g.files) if k.startswith("loss/")]
    for it in range(3):
        r = orc.learn(golden_batch(g, it), 0.99, 0.01)
        pol.learn(64, 0.99, 0.01, indices=idxs[it][None])
        loss = float(pol.last_metrics[0, 0])
        assert abs(loss - r["loss"]) <= 1e-5 * abs(r["loss"])
        assert abs(loss - ref_losses[it]) <= 1e-5 * abs(ref_losses[it])
        assert_module_close(pol.agent.Qnet, orc.q, "q after learn %d" % it)
    assert_module_close(pol.agent.Qnet, net_from_golden(g, "final/q/"), "final q")
    assert_module_close(pol.agent.Qnet_target, net_from_golden(g, "final/q_target/"), "final q_target")
    # multi-update launch == sequential launches
    pol2 = DQN([4, 2], False, 1e-3, 1000, device)
    load_into(pol2.agent.Qnet, init)
    load_into(pol2.agent.Qnet_target, init)
    fill_buffer_from_batches(pol2.buffer, g, 3)
    pol2.learn(64, 0.99, 0.01, n_updates=3, indices=np.stack(idxs))
    assert_module_close(pol2.agent.Qnet, pol.agent.Qnet.state_dict(), "fused 3 updates", tol=dict(rtol=0, atol=0))


def test_dqn_emulated(golden, emul):
    _run(golden, torch.device("cpu"))


@pytest.mark.gpu
def test_dqn_gpu(golden):
    _run(golden, torch.device("cuda"))
