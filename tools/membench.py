"""Micro-benchmarks of the memory-bound members of the hot path (replay add / gather, GAE scan, advantage
normalisation, PER sum-tree, Polyak sweep, batched policy inference) against the measured HBM peak.

    python tools/membench.py [--json profiles/rX_membench.json]

Every kernel is timed with CUDA events on the launching stream over `reps` launches after warm-up; working sets are
rotated through buffers larger than the 126 MB L2 where the kernel is meant to stream from HBM.  `GB/s` = ALGORITHMIC
bytes / mean launch time.  Not the bench: `bench.py` measures the training step."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from freerl_b200 import _common, _lib          # noqa: E402
from freerl_b200.Buffer import Buffer          # noqa: E402

dev = torch.device("cuda")


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timeit_flushed(fn, reps=10, warm=2, flush_mb=512):
    """every iteration timed alone with the L2 flushed first (a buffer of flush_mb > 126 MB is overwritten before each call): nothing of the
    previous iteration's input or output can be served by L2"""
    junk = torch.empty(flush_mb << 18, dtype=torch.float32, device=dev)
    for _ in range(warm):
        fn()
    tot = 0.0
    for i in range(reps):
        junk.fill_(float(i))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    L = _lib.lib()
    st = _lib.stream_ptr(dev)
    pk = peak()
    rows = []

    def rec(name, ms, nbytes, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "ms": ms, "algorithmic_bytes": int(nbytes), "GB/s": gbs, "frac_of_hbm_peak": gbs / pk, "note": note})
        print("%-44s %9.3f ms %10.1f MB %9.1f GB/s  %5.1f %%  %s" % (name, ms, nbytes / 1e6, gbs, 100 * gbs / pk, note))

    OBS, ACT = 17, 6
    g = torch.Generator(device=dev); g.manual_seed(0)
    # ---- replay add_batch: C2 rows (168 B of payload per transition read from 5 SoA tensors, 176 B row written) ----
    cap = 4_000_000                                 # 704 MB ring: every add streams to HBM
    buf = Buffer(cap, OBS, ACT, dev)
    n = 1_000_000
    o = torch.randn((n, OBS), device=dev, generator=g); a_ = torch.rand((n, ACT), device=dev, generator=g)
    r = torch.randn(n, device=dev, generator=g); o2 = torch.randn((n, OBS), device=dev, generator=g); d = torch.zeros(n, device=dev)
    ms = timeit(lambda: buf.add_device(o, a_, r, o2, d), reps=8)
    rec("frl_replay_add_batch n=1e6 (C2 rows)", ms, n * (168 + 176), "read SoA 168 B + write row 176 B per transition")
    rep4 = lambda x: x.repeat(*([4] + [1] * (x.dim() - 1)))
    o4, a4, r4, o24, d4 = rep4(o), rep4(a_), rep4(r), rep4(o2), rep4(d)
    ms = timeit(lambda: buf.add_device(o4, a4, r4, o24, d4), reps=4)
    rec("frl_replay_add_batch n=4e6 (C2 rows)", ms, 4 * n * (168 + 176), "the whole 704 MB ring in one launch (fixed launch / ramp cost amortised)")
    del o4, a4, r4, o24, d4
    # ---- replay gather: B rows sampled uniformly from the 704 MB ring ----
    for B in (65536, 1 << 20):
        idx = torch.randint(0, cap, (8, B), device=dev, generator=g)
        k = [0]
        def fn():
            buf.sample(idx[k[0] % 8]); k[0] += 1
        ms = timeit(fn, reps=8)
        rec("frl_replay_gather B=%d (C2 rows, 704 MB ring)" % B, ms, B * (176 + 8 + 168), "read row 176 B + index 8 B, write 168 B")
    # ---- small gather at the reference batch (launch-latency bound) ----
    idx = torch.randint(0, cap, (8, 256), device=dev, generator=g)
    k = [0]
    def fn():
        buf.sample(idx[k[0] % 8]); k[0] += 1
    ms = timeit(fn, reps=50)
    rec("frl_replay_gather B=256 (reference batch)", ms, 256 * (176 + 8 + 168), "includes 5 torch.empty allocations; latency bound")
    del buf, o, a_, r, o2, d
    torch.cuda.empty_cache()
    # ---- uniform sampling without replacement ----
    out = torch.empty((256, 256), dtype=torch.int64, device=dev)
    ms = timeit(lambda: L.frl_sample_uniform(_lib.ptr(out), 1_000_000, 256, 256, ctypes.c_uint64(1), ctypes.c_uint64(0), st), reps=20)
    rec("frl_sample_uniform 256 x B=256 of 1e6", ms, 256 * 256 * 8, "compute bound (Philox + duplicate scan); bytes = indices written")
    # ---- GAE scan ----
    for T, N in ((128, 1024), (2048, 1), (256, 512 * 3), (1024, 16384)):
        f = lambda: torch.randn((T, N), device=dev, generator=g)
        rw, dn, ad, vs, vn = f(), (f() > 2).float(), (f() > 1.5).float(), f(), f()
        adv, vt = torch.empty((T, N), device=dev), torch.empty((T, N), device=dev)
        ms = timeit(lambda: L.frl_gae(_lib.ptr(rw), _lib.ptr(dn), _lib.ptr(ad), _lib.ptr(vs), _lib.ptr(vn), T, N, 0.99, 0.95,
                                      _lib.ptr(adv), _lib.ptr(vt), st), reps=20)
        rec("frl_gae T=%d N=%d" % (T, N), ms, T * N * 28, "20 B read + 8 B written per element")
    # ---- advantage normalisation (MAPPO C5: 256 x 512 envs x 3 agents) ----
    for n_ in (256 * 3, 256 * 512 * 3, 1 << 24):
        x = torch.randn(n_, device=dev, generator=g); y = torch.empty_like(x)
        ms = timeit(lambda: L.frl_adv_norm(_lib.ptr(x), n_, ctypes.c_float(1e-8), _lib.ptr(y), st), reps=20)
        rec("frl_adv_norm n=%d" % n_, ms, n_ * 8, "4 B read + 4 B written" + (" (two launches: + one re-read)" if n_ > (1 << 21) else " (resident: one launch)"))
    x = torch.randn(1 << 24, device=dev, generator=g); y = torch.empty_like(x)
    ms = timeit_flushed(lambda: L.frl_adv_norm(_lib.ptr(x), 1 << 24, ctypes.c_float(1e-8), _lib.ptr(y), st))
    rec("frl_adv_norm n=16777216, L2 flushed before every call", ms, (1 << 24) * 8, "4 B read + 4 B written; cold caches, each call timed alone (includes both launches' ramp)")
    del x, y
    # ---- PER sum-tree (cap 1e6 like the reference default, non power of two) ----
    capt = 1_000_000
    tree = torch.zeros(2 * capt - 1, dtype=torch.float64, device=dev)
    tscr = torch.zeros(4096, dtype=torch.float64, device=dev)
    pri = torch.rand(capt, device=dev, generator=g) + 0.01
    # build by ordered range update in chunks (also the timed "add" path)
    B = 256
    ms = timeit(lambda: L.frl_sumtree_update(_lib.ptr(tree), capt, None, _lib.ptr(pri), None, 0.0, 0, 1, B, _lib.ptr(tscr), st), reps=20)
    rec("frl_sumtree_update B=256 cap=1e6 (ordered)", ms, B * 20 * 16, "B x ceil(log2 cap)=20 levels x (8 B read + 8 B write); one CTA per tree level, batch-ordered fp64 chains (bit-exact)")
    for i0 in range(0, capt, 1024):
        nb = min(1024, capt - i0)
        L.frl_sumtree_update(_lib.ptr(tree), capt, None, _lib.ptr(pri[i0:]), None, 0.0, i0, 1, nb, _lib.ptr(tscr), st)
    oi = torch.empty(B, dtype=torch.int64, device=dev); op = torch.empty(B, device=dev); ow = torch.empty(B, device=dev)
    ms = timeit(lambda: L.frl_sumtree_sample(_lib.ptr(tree), capt, None, ctypes.c_uint64(3), ctypes.c_uint64(0), B, capt, 0.4, 1e-7,
                                             _lib.ptr(oi), _lib.ptr(op), _lib.ptr(ow), st), reps=20)
    rec("frl_sumtree_sample B=256 cap=1e6", ms, B * 20 * 16, "B descents x 20 levels x 2 child reads of 8 B")
    scratch = torch.empty(1024, dtype=torch.float64, device=dev); mx = torch.empty(1, dtype=torch.float64, device=dev)
    ms = timeit(lambda: L.frl_sumtree_max(_lib.ptr(tree), capt, _lib.ptr(scratch), 1024, _lib.ptr(mx), st), reps=20)
    rec("frl_sumtree_max cap=1e6", ms, capt * 8, "np.max over the 1e6 float64 leaves (8 MB, L2 resident)")
    # ---- batched policy inference (select_action for N envs) ----
    from freerl_b200.SAC import SAC
    pol = SAC([OBS, ACT], True, 1e-3, 1e-3, 1024, dev, trick={}, mode="fast")
    for n_ in (256, 1024, 65536):
        x = torch.randn((n_, OBS), device=dev, generator=g)
        ms = timeit(lambda: _common.infer(pol.agent._actor, x, _lib.INFER_SAC_SAMPLE, dev, ACT, seed=1, counter=1), reps=20)
        rec("frl_policy_infer SAC n=%d" % n_, ms, n_ * (OBS + ACT) * 4 + 19596 * 4, "FLOP bound: %.2f GFLOP/s" % (2 * 19328 * n_ / (ms * 1e-3) / 1e9))
    # ---- Polyak sweep (standalone, MADDPG path) ----
    ag = pol.agent
    ms = timeit(lambda: L.frl_polyak(ctypes.byref(ag._critic.c_struct()), ctypes.byref(ag._critic_t.c_struct()), ctypes.c_float(0.01), st), reps=50)
    rec("frl_polyak twin critic (39 426 params)", ms, 39426 * 16, "12 B/param + 4 B mirror write; launch-latency bound at this size")
    if args.json:
        json.dump({"hbm_peak_gbs": pk, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
