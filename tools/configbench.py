"""learn() throughput of every BASELINE.json configuration at its SURVEY §8(d) shape on one B200 — NOT the bench (bench.py
measures config 2 end to end); this is the secondary table DESIGN.md §3 quotes for the other algorithm kernels.

    python tools/configbench.py [--json profiles/rX_configs.json]

C1 DQN CartPole (obs 4, 2 actions, B 256) · C2 SAC (obs 17, act 6, B 256, 1M replay) · C3 PPO (1024 envs x T 128 = 131 072
rows, minibatch 8192, K 10 = 160 updates / learn) · C4 Rainbow (obs 8, 4 actions, 51 atoms, PER + n-step, B 256, capacity 1e6)
· C5 MAPPO (3 agents, obs 18, act 5, 512 envs x horizon 256 = 131 072 rows, minibatch = all rows, K 15).
Timed with CUDA events over whole learn() calls (fast mode: on-device sampling / noise), after warm-up.
"""
import argparse
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = torch.device("cuda")


def timeit(fn, reps, warm=2, median=False):
    """mean over `reps` back-to-back calls (short calls: host launch work overlaps the previous call, as in a training loop), or with
    median=True the median over calls bracketed one by one (the 20-70 ms cooperative launches occasionally run 10-30 % slow; a mean of
    a few calls would carry that into the table)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if not median:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    each = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        each.append(e0.elapsed_time(e1))
    return float(np.median(each))


def run(which=("C1", "C2", "C3", "C4", "C5"), log=print):
    """Time the learn() of the named configurations; returns the list of row dicts (bench.py puts it under extra.configs)."""
    rows = []
    rng = np.random.default_rng(0)

    def rec(name, ms, updates, transitions, note):
        rows.append({"config": name, "ms_per_learn_call": ms, "updates_per_call": updates, "updates_per_sec": updates / ms * 1e3,
                     "us_per_update": ms * 1e3 / updates, "transitions_per_sec": transitions / ms * 1e3, "note": note})
        log("%-34s %9.3f ms/call  %6d updates  %9.1f us/update  %10.0f updates/s  %s" % (name, ms, updates, ms * 1e3 / updates, updates / ms * 1e3, note))

    with contextlib.redirect_stdout(sys.stderr):
        from freerl_b200.DQN import DQN
        from freerl_b200.DQN_with_tricks import DQN as Rainbow
        from freerl_b200.MAPPO import MAPPO
        from freerl_b200.PPO import PPO
        from freerl_b200.SAC import SAC

    if "C1" in which:
        _c1(rec, rng, DQN)
    if "C2" in which:
        _c2(rec, rng, SAC)
    if "C3" in which:
        _c3(rec, rng, PPO)
    if "C4" in which:
        _c4(rec, rng, Rainbow)
    if "C5" in which:
        _c5(rec, rng, MAPPO)
    return rows


def _c1(rec, rng, DQN):
    # ---- C1: DQN CartPole dims ----
    with contextlib.redirect_stdout(sys.stderr):
        pol = DQN([4, 2], False, 1e-3, 1e6, dev, mode="fast")
    n = 100_000
    pol.add(rng.standard_normal((n, 4)), rng.integers(0, 2, (n, 1)), rng.standard_normal(n), rng.standard_normal((n, 4)), rng.random(n) < 0.005)
    ms = timeit(lambda: pol.learn(256, 0.99, 0.01, n_updates=256), 5)
    rec("C1 DQN CartPole B=256", ms, 256, 256 * 256, "256 sequential learns per launch")


def _c2(rec, rng, SAC):
    # ---- C2: SAC (the bench workload's kernel) ----
    with contextlib.redirect_stdout(sys.stderr):
        pol = SAC([17, 6], True, 1e-3, 1e-3, 1e6, dev, trick={}, mode="fast")
    for _ in range(4):
        n = 250_000
        pol.add(rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32), rng.standard_normal(n).astype(np.float32),
                rng.standard_normal((n, 17), dtype=np.float32), rng.random(n) < 0.001)
    ms = timeit(lambda: pol.learn(256, 0.99, 0.01, n_updates=256), 5)
    rec("C2 SAC HalfCheetah B=256 (1M)", ms, 256, 256 * 256, "256 sequential learns per launch")
    ms = timeit(lambda: pol.learn(65536, 0.99, 0.01, n_updates=1), 5)
    rec("C2 SAC B=65536 (1 update/vec step)", ms, 1, 65536, "SURVEY 8d alternative: one big-batch update per vector step")
    del pol


def _c3(rec, rng, PPO):
    # ---- C3: PPO 1024 envs ----
    T, N, mb, K = 128, 1024, 8192, 10
    with contextlib.redirect_stdout(sys.stderr):
        pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev, mode="fast")
    data = [(rng.standard_normal((N, 8), dtype=np.float32), rng.integers(0, 4, (N, 1)).astype(np.float32), rng.standard_normal(N).astype(np.float32),
             rng.standard_normal((N, 8), dtype=np.float32), rng.random(N) < 1 / 300, -np.abs(rng.standard_normal((N, 1))).astype(np.float32) * 0.1 - 1.3)
            for _ in range(T)]

    def fill_ppo():
        for o, a, r, o2, d, lp in data:
            pol.add(o, a, r, o2, d, lp, d)

    def ppo_learn():
        pol.buffer._index, pol.buffer._size, pol.buffer.n_envs = 0, T * N, N          # rollout already resident (learn() clears it)
        pol.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
    fill_ppo()
    ms = timeit(ppo_learn, 5, warm=1, median=True)
    rec("C3 PPO 1024 envs x 128, mb 8192, K 10", ms, K * (T * N // mb), T * N, "GAE scan + 160 minibatch updates in one launch; transitions/s = rollout rows consumed")
    adv_ms = timeit(lambda: pol.compute_gae(0.99, 0.95), 10)
    rec("C3 PPO critic(obs), critic(obs') + GAE", adv_ms, 1, T * N, "2 batched value inferences + frl_gae over [128, 1024]")
    del pol


def _c4(rec, rng, Rainbow):
    # ---- C4: Rainbow ----
    trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
    with contextlib.redirect_stdout(sys.stderr):
        pol = Rainbow([8, 4], False, 1e-3, 1e6, dev, trick=trick, gamma=0.99, batch_size=256, mode="fast")
    for _ in range(60):
        pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)), rng.random(512) < 0.01)
    ms = timeit(lambda: pol.learn(256, 0.99, 0.01), 30, warm=3)
    rec("C4 Rainbow B=256, PER cap 1e6", ms, 1, 256, "PER stratified sample + fused C51/Dueling/Noisy learn + ordered priority update")
    add_ms = timeit(lambda: pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)),
                                    rng.random(512) < 0.01), 10)
    rec("C4 Rainbow add of 512 envs (n-step + PER)", add_ms, 1, 512, "host n-step fold + H2D + tree max + 512 ordered leaf updates")
    del pol


def _c5(rec, rng, MAPPO):
    # ---- C5: MAPPO ----
    # the switch set the reference's default policy_name 'MAPPO' forces (MAPPO.py:616-622)
    MAPPO_TRICK = {'adv_norm': True, 'ObsNorm': True, 'reward_norm': False, 'reward_scaling': True, 'orthogonal_init': True,
                   'adam_eps': True, 'lr_decay': False, 'ValueClip': True, 'huber_loss': True, 'LayerNorm': True, 'feature_norm': True}
    H, E, K5 = 256, 512, 15
    ids = ["agent_%d" % i for i in range(3)]
    with contextlib.redirect_stdout(sys.stderr):
        pol = MAPPO({k: [18, 5] for k in ids}, True, 1e-3, 1e-3, H * E, dev, dict(MAPPO_TRICK), mode="fast")
    step = {k: (rng.standard_normal((E, 18), dtype=np.float32), rng.uniform(-1, 1, (E, 5)).astype(np.float32), rng.standard_normal(E).astype(np.float32),
                rng.standard_normal((E, 18), dtype=np.float32), np.zeros(E, bool), (-np.abs(rng.standard_normal((E, 5))) * 0.1 - 0.9).astype(np.float32))
            for k in ids}
    for t in range(H):
        trunc = np.full(E, (t % 25) == 24)
        pol.add({k: step[k][0] for k in ids}, {k: step[k][1] for k in ids}, {k: step[k][2] for k in ids}, {k: step[k][3] for k in ids},
                {k: step[k][4] for k in ids}, {k: step[k][5] for k in ids}, {k: trunc for k in ids})

    def mappo_learn():
        for b in pol.buffers.values():
            b._index, b._size, b.n_envs = 0, H * E, E
        pol.learn(H * E, 0.95, 0.95, 0.2, K5, 0.01, 10.0)
    ms = timeit(mappo_learn, 3, warm=1, median=True)
    rec("C5 MAPPO 3 agents, 512 envs x 256, K 15", ms, 3 * K5, H * E, "joint GAE + adv-norm + 15 full-batch updates per agent (minibatch = horizon x envs)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default="C1,C2,C3,C4,C5")
    args = ap.parse_args()
    rows = run(tuple(args.only.split(",")))
    out = {"device": torch.cuda.get_device_name(0), "rows": rows, "umma": not os.environ.get("FREERL_B200_NO_UMMA")}
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
