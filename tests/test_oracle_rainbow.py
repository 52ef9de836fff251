"""Rainbow oracle (oracle/rainbow.py + oracle/buffers.py PER) replayed against the reference-generated fixture."""
from collections import OrderedDict

import numpy as np
import torch

from oracle import buffers
from oracle.rainbow import RainbowOracle


def rainbow_setup(g):
    q = OrderedDict((k[len("init/q/"):], torch.from_numpy(g[k].copy())) for k in g.files if k.startswith("init/q/"))
    per = buffers.PrioritizedReplay(500, 8, 1)
    n = g["buf/obs"].shape[0]
    per.buffer.obs[:n], per.buffer.actions[:n], per.buffer.rewards[:n] = g["buf/obs"], g["buf/act"], g["buf/rew"]
    per.buffer.next_obs[:n], per.buffer.dones[:n] = g["buf/nobs"], g["buf/done"]
    per.buffer._index, per.buffer._size = [int(x) for x in g["init/index"]]
    per.sumtree.tree[:] = g["init/tree"]
    return q, per


def eps_of(g, it):
    return [[torch.from_numpy(g["eps/%d/%d/%d" % (it, f, j)]) for j in range(4)] for f in range(3)]


def test_rainbow_oracle_vs_reference(golden):
    g = golden("rainbow")
    q, per = rainbow_setup(g)
    orc = RainbowOracle(q, 1e-3, 4)
    B = 32
    gamma_n = float(g["n_step_gamma"])
    for it in range(3):
        u = g["u/%d" % it]
        # PER sample with the recorded uniforms (np.random.uniform(a,b) = a + (b-a)*u)
        seg = per.sumtree.total() / B
        per.beta = np.min([1., per.beta + per.beta_increment])
        idx = np.zeros(B, np.int64)
        pri = np.zeros(B, np.float32)
        for i in range(B):
            a, b = seg * i, seg * (i + 1)
            pri[i], idx[i] = per.sumtree.find(a + (b - a) * u[i])
        prob = np.clip(pri / per.sumtree.total(), 1e-7, None)
        w = (len(per) * prob) ** (-per.beta)
        w = (w / w.max()).astype(np.float32)
        batch = tuple(torch.from_numpy(x) for x in per.buffer.sample(idx))
        r = orc.learn(batch, eps_of(g, it), gamma_n, 0.01, is_weight=torch.from_numpy(w), double_q=True)
        np.testing.assert_allclose(r["loss"], g["losses"][it], rtol=1e-6)
        per.update_priorities(idx, r["error"].numpy())
        np.testing.assert_allclose(per.sumtree.tree, g["tree/%d" % it], rtol=1e-6, atol=1e-9)
    for k, v in orc.q.items():
        np.testing.assert_allclose(v.detach().numpy(), g["final/q/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
    for k, v in orc.q_target.items():
        np.testing.assert_allclose(v.detach().numpy(), g["final/q_target/" + k], rtol=2e-5, atol=2e-6, err_msg=k)
