"""GPU check of the tensor-core PPO path (csrc/algo_ppo_umma.cuh):
  1. ONE 2048-row update, tensor-core path vs the FFMA tile path (FREERL_B200_NO_UMMA=1) from identical parameters: reduced gradient
     net.g per tensor, losses, parameters;
  2. two 1024-row updates vs the oracle (the body of tests/test_parity_ppo.py::_ppo_large_minibatch) with per-tensor differences."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from collections import OrderedDict
from oracle import algos
from freerl_b200.PPO import PPO

dev = torch.device("cuda")


def build(is_continue, T, N, umma):
    if umma:
        os.environ.pop("FREERL_B200_NO_UMMA", None)
    else:
        os.environ["FREERL_B200_NO_UMMA"] = "1"
    torch.manual_seed(4)
    ad = 2 if is_continue else 4
    pol = PPO([8, ad], is_continue, 1e-3, 1e-3, T * N, dev)
    rng = np.random.default_rng(12)
    cols = []
    for t in range(T):
        o, o2 = rng.standard_normal((N, 8), dtype=np.float32), rng.standard_normal((N, 8), dtype=np.float32)
        act, lp = pol.select_action(o)
        act = np.asarray(act, dtype=np.float32).reshape(N, -1)
        lp = np.asarray(lp, dtype=np.float32).reshape(N, -1)
        r = rng.standard_normal(N).astype(np.float32)
        d = rng.random(N) < 0.02
        adn = d | (rng.random(N) < 0.02)
        pol.add(o, act, r, o2, d, lp, adn)
        cols.append((o, act, r.reshape(N, 1), o2, d.reshape(N, 1).astype(np.float32), lp, adn.reshape(N, 1).astype(np.float32)))
    data = tuple(torch.from_numpy(np.concatenate([c[k] for c in cols])) for k in range(7))
    return pol, data, rng


def seg_report(net, ga, gb, what):
    L = net.layers if hasattr(net, "layers") else None
    a, b = ga.cpu().numpy(), gb.cpu().numpy()
    print("  %s: max|diff| %.3e  max|ref| %.3e  (n = %d, first bad index %s)" % (what, np.abs(a - b).max(), np.abs(b).max(), a.size,
          np.argmax(np.abs(a - b) > 1e-5 * np.abs(b).max() + 1e-9) if (np.abs(a - b) > 1e-5 * np.abs(b).max() + 1e-9).any() else None))
    c = net.c_struct()
    for li in range(c.n_layers):
        Ly = c.L[li]
        for nm, off, ln in (("W", Ly.w_off, Ly.out_pad * Ly.in_pad), ("b", Ly.b_off, Ly.out_pad)):
            x, y = a[off:off + ln], b[off:off + ln]
            print("    layer %d %s  max|diff| %.3e  max|ref| %.3e" % (li, nm, np.abs(x - y).max(), np.abs(y).max()))
    if c.x_len:
        x, y = a[c.x_off:c.x_off + c.x_len], b[c.x_off:c.x_off + c.x_len]
        print("    extra      max|diff| %.3e  max|ref| %.3e" % (np.abs(x - y).max(), np.abs(y).max()))


def ab(is_continue):
    T, N = 16, 128
    res = {}
    for umma in (False, True):
        pol, data, rng = build(is_continue, T, N, umma)
        perm = rng.permutation(T * N)
        pol.learn(T * N, 0.99, 0.95, 0.2, 1, 0.01, permutations=[perm])
        torch.cuda.synchronize()
        res[umma] = (pol.agent._net.g.clone(), pol.last_metrics.cpu().numpy().copy(), pol.agent._net.p.clone(), pol)
    print("A/B one 2048-row update,", "continuous" if is_continue else "discrete")
    print("  metrics FFMA", res[False][1][0, :5])
    print("  metrics UMMA", res[True][1][0, :5])
    seg_report(res[True][3].agent._net, res[True][0], res[False][0], "gradient UMMA vs FFMA")
    d = (res[True][2] - res[False][2]).abs().max().item()
    print("  parameters after the update: max|diff| %.3e" % d)


def vs_oracle(is_continue, mb=1024, T=16, N=128):
    pol, data, rng = build(is_continue, T, N, True)
    sd = lambda m: OrderedDict((k, v.detach().cpu().clone()) for k, v in m.state_dict().items())
    orc = algos.PPOOracle(sd(pol.agent.actor), sd(pol.agent.critic), 1e-3, is_continue)
    with torch.no_grad():
        vs, vn = algos.mlp2(orc.critic, data[0]), algos.mlp2(orc.critic, data[3])
        td = (data[2] + 0.99 * (1.0 - data[4]) * vn - vs).numpy().reshape(T, N).astype(np.float64)
    adn = data[6].numpy().reshape(T, N).astype(np.float64)
    want, g = np.zeros((T, N)), np.zeros(N)
    for t in reversed(range(T)):
        g = td[t] + 0.99 * 0.95 * g * (1.0 - adn[t])
        want[t] = g
    adv_o = torch.from_numpy(want.astype(np.float32).reshape(-1, 1))
    perm = rng.permutation(T * N)
    ref = [orc.minibatch(data, adv_o, adv_o + vs, perm[s:s + mb], 0.2, 0.01) for s in range(0, T * N, mb)]
    pol.learn(mb, 0.99, 0.95, 0.2, 1, 0.01, permutations=[perm])
    torch.cuda.synchronize()
    m = pol.last_metrics.cpu().numpy()
    print("vs oracle, two 1024-row updates,", "continuous" if is_continue else "discrete")
    print("  actor loss  got", m[:, 0], "want", [float(x[0]) for x in ref])
    print("  critic loss got", m[:, 1], "want", [float(x[1]) for x in ref])
    for name, mine, theirs in (("actor", pol.agent.actor, orc.actor), ("critic", pol.agent.critic, orc.critic)):
        a = mine.state_dict()
        for k in a:
            x, y = a[k].detach().cpu().numpy(), theirs[k].detach().cpu().numpy()
            print("  %-6s %-12s max|diff| %.3e  max|ref| %.3e" % (name, k, np.abs(x - y).max(), np.abs(y).max()))


def mappo_ab(is_continue=True):
    """MAPPO (LayerNorm nets, centralised critic on the joint observation, huber value loss): one 2048-row full-batch update per
    agent, tensor-core path vs FFMA tile path from identical parameters."""
    from freerl_b200.MAPPO import MAPPO
    from oracle.make_golden_marl import MAPPO_TRICK          # the trick dict only
    H, E = 32, 64
    ids = ["agent_%d" % i for i in range(3)]
    res = {}
    for umma in (False, True):
        if umma:
            os.environ.pop("FREERL_B200_NO_UMMA", None)
        else:
            os.environ["FREERL_B200_NO_UMMA"] = "1"
        torch.manual_seed(7)
        np.random.seed(7)
        pol = MAPPO({k: [18, 5] for k in ids}, is_continue, 1e-3, 1e-3, H * E, dev, dict(MAPPO_TRICK))
        rng = np.random.default_rng(3)
        for t in range(H):
            obs = {k: rng.standard_normal((E, 18), dtype=np.float32) for k in ids}
            act, lp = {}, {}
            for k in ids:
                if is_continue:
                    act[k] = rng.uniform(-1, 1, (E, 5)).astype(np.float32)
                    lp[k] = (-np.abs(rng.standard_normal((E, 5))) * 0.1 - 0.9).astype(np.float32)
                else:
                    act[k] = rng.integers(0, 5, (E, 1)).astype(np.float32)
                    lp[k] = (-np.abs(rng.standard_normal((E, 1))) * 0.1 - 1.5).astype(np.float32)
            rew = {k: rng.standard_normal(E).astype(np.float32) for k in ids}
            nobs = {k: rng.standard_normal((E, 18), dtype=np.float32) for k in ids}
            done = {k: np.zeros(E, bool) for k in ids}
            trunc = np.full(E, (t % 25) == 24)
            pol.add(obs, act, rew, nobs, done, lp, {k: trunc for k in ids})
        pol.learn(H * E, 0.95, 0.95, 0.2, 1, 0.01, 10.0)
        torch.cuda.synchronize()
        res[umma] = (pol, pol.last_metrics.cpu().numpy().copy())
    print("MAPPO A/B one 2048-row update per agent,", "continuous" if is_continue else "discrete")
    print("  metrics FFMA", res[False][1][:, :5].tolist())
    print("  metrics UMMA", res[True][1][:, :5].tolist())
    for k in ids:
        nu, nf = res[True][0].agents[k]._net, res[False][0].agents[k]._net
        seg_report(nu, nu.g, nf.g, "%s gradient UMMA vs FFMA" % k)
        print("  %s parameters after the update: max|diff| %.3e" % (k, (nu.p - nf.p).abs().max().item()))


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "mappo"):
    mappo_ab(True)
    mappo_ab(False)
for cont in (False, True):
    if which in ("all", "ab"):
        ab(cont)
    if which in ("all", "oracle"):
        vs_oracle(cont)
