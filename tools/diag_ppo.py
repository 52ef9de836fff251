import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')  # run from the repo root: python tools/diag_ppo.py
import numpy as np, torch
from collections import OrderedDict
dev = torch.device(os.environ.get("DEV", "cpu"))
from freerl_b200.PPO import PPO
from oracle import algos
def _sd(m): return OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())
def run(T, N, mb, K, real_logp):
    torch.manual_seed(0)
    pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev)
    rng = np.random.default_rng(2)
    cols = []
    for t in range(T):
        o, o2 = rng.standard_normal((N, 8), dtype=np.float32), rng.standard_normal((N, 8), dtype=np.float32)
        if real_logp:
            a, lp = pol.select_action(o)
            a = a.reshape(N, 1).astype(np.float32); lp = lp.reshape(N, 1).astype(np.float32)
        else:
            a = rng.integers(0, 4, (N, 1)).astype(np.float32)
            lp = -np.abs(rng.standard_normal((N, 1))).astype(np.float32)
        r = rng.standard_normal(N).astype(np.float32)
        d = rng.random(N) < 1 / 300
        ad = d | (rng.random(N) < 1 / 500)
        pol.add(o, a, r, o2, d, lp, ad)
        cols.append((o, a, r.reshape(N, 1), o2, d.reshape(N, 1).astype(np.float32), lp, ad.reshape(N, 1).astype(np.float32)))
    data = tuple(torch.from_numpy(np.concatenate([c[k] for c in cols])) for k in range(7))
    orc = algos.PPOOracle(_sd(pol.agent.actor), _sd(pol.agent.critic), 1e-3, False)
    perms = [rng.permutation(T * N) for _ in range(K)]
    with torch.no_grad():
        vs, vn = algos.mlp2(orc.critic, data[0]), algos.mlp2(orc.critic, data[3])
        td = (data[2] + 0.99 * (1.0 - data[4]) * vn - vs).numpy().reshape(T, N).astype(np.float64)
    adn = data[6].numpy().reshape(T, N).astype(np.float64)
    want = np.zeros((T, N)); g = np.zeros(N)
    for t in reversed(range(T)):
        g = td[t] + 0.99 * 0.95 * g * (1.0 - adn[t]); want[t] = g
    adv_o = torch.from_numpy(want.astype(np.float32).reshape(-1, 1)); vt_o = adv_o + vs
    r = {"losses": [orc.minibatch(data, adv_o, vt_o, perm[s:s + mb], 0.2, 0.01) for perm in perms for s in range(0, T * N, mb)]}
    pol.learn(mb, 0.99, 0.95, 0.2, K, 0.01, permutations=perms)
    m = pol.last_metrics.cpu().numpy(); ref = np.array(r["losses"])
    print("T%d N%d mb%d K%d real_logp=%s: loss rel err actor %.2e critic %.2e" % (T, N, mb, K, real_logp, np.abs(m[:,0]/ref[:,0]-1).max(), np.abs(m[:,1]/ref[:,1]-1).max()))
    print("   per-update actor loss rel err:", " ".join("%.1e" % x for x in np.abs(m[:,0]/ref[:,0]-1)))
    print("   per-update critic loss rel err:", " ".join("%.1e" % x for x in np.abs(m[:,1]/ref[:,1]-1)))
    print("   grad norms a/c:", " ".join("%.3g/%.3g" % (x, y) for x, y in zip(m[:,3], m[:,4])))
    for name, mod, o in (("actor", pol.agent.actor, orc.actor), ("critic", pol.agent.critic, orc.critic)):
        for k, v in mod.state_dict().items():
            a, b = v.cpu().numpy(), o[k].detach().numpy()
            print("   %-7s %-18s max|d| %.2e  frac>2e-5 %.3f" % (name, k, np.abs(a - b).max(), (np.abs(a - b) > 2e-5).mean()))
for cfg in eval(os.environ.get("CFGS", "[(16,128,2048,1,False),(16,128,2048,1,True),(16,128,512,4,False)]")):
    run(*cfg)
