"""Diagnostic for the teacher-forced PPO chain (tests/test_parity_ppo.py::_ppo_teacher_forced): per step, compare OUR reduced gradient
(net.g) with the oracle's autograd gradient and count cautious-mask bits (m * g > 0) that differ, with the size of the gradients involved."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from collections import OrderedDict
from oracle import algos
import test_parity_ppo as T
from freerl_b200.PPO import PPO

dev = torch.device("cuda")
is_continue = True
Tn, N, mb = 16, 1024, 1024
torch.manual_seed(9)
ad = 2
H = Tn * N
pol = PPO([8, ad], is_continue, 1e-3, 1e-3, H, dev)
rng = np.random.default_rng(21)
cols = []
for t in range(Tn):
    o, o2 = rng.standard_normal((N, 8), dtype=np.float32), rng.standard_normal((N, 8), dtype=np.float32)
    act, lp = pol.select_action(o)
    act, lp = np.asarray(act, dtype=np.float32).reshape(N, -1), np.asarray(lp, dtype=np.float32).reshape(N, -1)
    r = rng.standard_normal(N).astype(np.float32)
    d = rng.random(N) < 0.02
    adn = d | (rng.random(N) < 0.02)
    pol.add(o, act, r, o2, d, lp, adn)
    cols.append((o, act, r.reshape(N, 1), o2, d.reshape(N, 1).astype(np.float32), lp, adn.reshape(N, 1).astype(np.float32)))
data = tuple(torch.from_numpy(np.concatenate([c[k] for c in cols])) for k in range(7))
sd = lambda m: OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())
orc = algos.PPOOracle(sd(pol.agent.actor), sd(pol.agent.critic), 1e-3, is_continue)
adv, vt = pol.compute_gae(0.99, 0.95)
adv_o, vt_o = adv.cpu(), vt.cpu()
perm = rng.permutation(H)
cap = {}
orig = algos.cautious_adamw_step


def tap(params, grads, st):
    cap["g"] = [g.clone() for g in grads]
    cap["m_prev"] = [m.clone() for m in st.m]
    return orig(params, grads, st)


algos.cautious_adamw_step = tap
names = list(orc.actor.keys()) + ["critic." + k for k in orc.critic.keys()]
net = pol.agent._net
for u in range(6):
    T._push_oracle_state(pol, orc, is_continue)
    index = perm[u * mb:(u + 1) * mb]
    want = orc.minibatch(data, adv_o, vt_o, index, 0.2, 0.01)
    idx = torch.from_numpy(index.astype(np.int64))[None].to(dev)
    rows = torch.tensor([mb], dtype=torch.int32, device=dev)
    pol._minibatch_plan = lambda *a, **k: (idx, rows, 1)
    pol._update(adv, vt, mb, 1, 0.2, 0.01, None)
    m = pol.last_metrics.cpu().numpy()[0]
    print("step %d  losses ours %.7f %.7f  oracle %.7f %.7f  gnorm ours %.6f %.6f" % (u, m[0], m[1], want[0], want[1], m[3], m[4]))
    g_ours = net.g.cpu()
    for i, nm in enumerate(names):
        if not nm.endswith("weight") or nm.startswith("critic"):
            continue
        li = {"l1": 0, "l2": 1, "l3": 2, "mean_layer": 2}[nm.split(".")[0]]
        go = net._state_like(g_ours, li).numpy()
        ga_clipped = cap["g"][i].numpy()
        # ours is unclipped: rescale by the oracle's clip coefficient implied by ga_clipped / raw (estimate from norms)
        coef = np.linalg.norm(ga_clipped) / max(np.linalg.norm(go), 1e-30)
        gs = go * coef
        mp = cap["m_prev"][i].numpy()
        m_or, m_us = 0.9 * mp + 0.1 * ga_clipped, 0.9 * mp + 0.1 * gs
        bit_or, bit_us = (m_or * ga_clipped > 0), (m_us * gs > 0)
        diff = bit_or != bit_us
        rel = np.abs(gs - ga_clipped).max() / np.abs(ga_clipped).max()
        print("   %-18s max|g| %.3e  max|dg|/max|g| %.2e  mask bits differing %d / %d  |g| of those (ours) max %.3e  median %.3e" % (
            nm, np.abs(ga_clipped).max(), rel, diff.sum(), diff.size, np.abs(gs[diff]).max() if diff.any() else 0.0,
            np.median(np.abs(gs[diff])) if diff.any() else 0.0))
        if diff.any() and nm == "l1.weight":
            rr = np.argwhere(diff)[:6]
            for (a_, b_) in rr:
                print("        [%d,%d] g ours %.4e oracle %.4e  m_prev %.4e" % (a_, b_, gs[a_, b_], ga_clipped[a_, b_], mp[a_, b_]))
