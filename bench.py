#!/usr/bin/env python
"""bench.py — env-steps/s of the SAC training hot path (BASELINE.json config[1]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on host cores

Workload (config.workload = "sac_halfcheetah_256env_1Mreplay"): SAC with HalfCheetah-v4 dims (obs 17, act 6,
hidden 128-128, twin critic), 256 vectorised synthetic envs per GPU, a pre-filled 1 000 000-transition device
replay per GPU, batch 256, update-to-data ratio 1 (the reference trains once per env step, SAC_file/SAC.py:571-572).
One "step" = VSTEPS (8) vector steps (so that the driver's 20 timed steps last about two seconds); a vector step = policy
inference for 256 envs, 256 transitions added to the replay and 256 sequential learn() updates (ONE persistent kernel launch).  `value` keeps all inputs resident in HBM (synthetic env arrays come
from a device pool); `e2e` drives the public Python API with HOST (numpy) buffers: select_action(host obs) -> host
synthetic env -> add(host arrays) -> learn(…, n_updates=256) -> D2H read of the loss, with the H2D / D2H copies inside
the timed region.  Replay rows are sampled uniformly from 176 MB (> the 126 MB L2), so batch gathers are HBM traffic.
Multi-GPU: weak scaling, one process per GPU, env workers + replay sharded per GPU, no data-path collective; the
replicas are kept one policy by a parameter all-reduce (NCCL) after every vector step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS, ACT, NENV, CAP, BATCH = 17, 6, 256, 1_000_000, 256
VSTEPS = 8                                # vector steps per bench step
GAMMA, TAU = 0.99, 0.01
PARAMS = 19596 + 39426                    # SAC actor + twin critic
ALGO_BYTES_PER_LEARN = BATCH * 168 + PARAMS * 36 + 16       # sampled rows + param/Adam/target traffic
ALGO_FLOPS_PER_LEARN = 0.70e6 * BATCH


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class SyntheticVecEnv:
    """obs ~ N(0,1) fp32 [N, obs_dim], reward ~ N(0,1), p_term = 0, truncation every 1000 steps.
    Transitions come from a pre-generated host pool so the host cost per step is a memcpy, like a fast simulator."""

    def __init__(self, n, seed, pool=64):
        r = np.random.default_rng(seed)
        self.obs_pool = r.standard_normal((pool, n, OBS), dtype=np.float32)
        self.rew_pool = r.standard_normal((pool, n)).astype(np.float32)
        self.n, self.t, self.pool = n, 0, pool
        self.obs = self.obs_pool[0]

    def step(self, action):
        self.t += 1
        nxt = self.obs_pool[self.t % self.pool]
        rew = self.rew_pool[self.t % self.pool]
        term = np.zeros(self.n, bool)
        trunc = np.full(self.n, self.t % 1000 == 0)
        return nxt, rew, term, trunc


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NCU_CAPTURE = "profiles/r6_bench_sac_learn_ncu_full.json"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE bench launch (256 learns) from the committed `ncu --set full`
    capture of this same command and kernel build (NCU_CAPTURE; a profiler cannot run inside the timed bench, so the
    number is a stored measurement and `traffic_source` says which); None when no capture is committed."""
    p = os.path.join(ROOT, NCU_CAPTURE)
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------
def cpu_reference_arm(steps, warmup, transitions_per_step=4, fill=CAP, threads=None, device="cpu", choice=True):
    """The reference's own CPU implementation of the path (oracle port: numpy ring replay with
    np.random.choice(len, 256, replace=False) + PyTorch-CPU SAC learn), all host threads.  A step = a bounded
    sample: `transitions_per_step` single-env iterations of select_action + env + add + learn.
    device="cuda": the SAME torch code with the networks on the GPU (the like-for-like "stock PyTorch on this B200" arm:
    numpy replay on the host, batch copied H2D per learn, as the reference does with --device cuda); choice=False replaces
    np.random.choice(1e6, 256, replace=False) — a full permutation per sample, 60 % of the CPU arm — by rng.integers."""
    import torch
    from collections import OrderedDict
    from oracle import algos, buffers
    # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host core, so set the pool explicitly and keep
    # whichever of {all cores, 1 thread} runs the learn faster (the 256x128 matmuls do not always scale with threads)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(0)
    np.random.seed(0)
    rng = np.random.default_rng(0)

    def lin(i, o):
        l = torch.nn.Linear(i, o)
        return l.weight.detach().clone(), l.bias.detach().clone()
    actor, critic = OrderedDict(), OrderedDict()
    for n_, (i, o) in zip(("l1", "l2", "mean_layer"), ((OBS, 128), (128, 128), (128, ACT))):
        actor[n_ + ".weight"], actor[n_ + ".bias"] = lin(i, o)
    actor["log_std"] = torch.zeros(1, ACT)
    actor.move_to_end("log_std", last=False)
    for k in range(6):
        i, o = ((OBS + ACT, 128), (128, 128), (128, 1))[k % 3]
        critic["l%d.weight" % (k + 1)], critic["l%d.bias" % (k + 1)] = lin(i, o)
    tdev = torch.device(device)
    if tdev.type == "cuda":
        actor = OrderedDict((k, v.to(tdev)) for k, v in actor.items())
        critic = OrderedDict((k, v.to(tdev)) for k, v in critic.items())
    orc = algos.SACOracle(actor, critic, 1e-3, 1e-3, act_dim=ACT)
    buf = buffers.RingReplay(CAP, OBS, ACT)
    n = int(fill)
    buf.obs[:n] = rng.standard_normal((n, OBS), dtype=np.float32)
    buf.actions[:n] = rng.uniform(-1, 1, (n, ACT))
    buf.rewards[:n] = rng.standard_normal(n)
    buf.next_obs[:n] = rng.standard_normal((n, OBS), dtype=np.float32)
    buf._size, buf._index = n, n % CAP
    obs = rng.standard_normal(OBS).astype(np.float32)

    def one_step():
        nonlocal obs
        for _ in range(transitions_per_step):
            with torch.no_grad():
                a, _ = algos.sac_actor(orc.actor, torch.as_tensor(obs).reshape(1, -1).to(tdev), torch.randn(1, ACT, device=tdev))
            nxt = rng.standard_normal(OBS).astype(np.float32)
            buf.add(obs, a.cpu().numpy()[0], float(rng.standard_normal()), nxt, False)
            obs = nxt
            idx = buffers.uniform_indices(len(buf), BATCH) if choice else rng.integers(0, len(buf), BATCH)
            batch = tuple(torch.from_numpy(x).to(tdev) for x in buf.sample(idx))
            orc.learn(batch, torch.randn(BATCH, ACT, device=tdev), torch.randn(BATCH, ACT, device=tdev), GAMMA, TAU)
    if not threads and (os.cpu_count() or 1) > 1 and tdev.type == "cpu":
        def probe():
            t0 = time.perf_counter()
            for _ in range(2):
                one_step()
            return time.perf_counter() - t0
        probe()
        t_all = probe()
        torch.set_num_threads(1)
        if probe() > t_all:
            torch.set_num_threads(os.cpu_count())
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    return steps * transitions_per_step / dt, dt * 1e3 / steps, torch.get_num_threads(), transitions_per_step


# --------------------------------------------------------------------------------------------------
def _timed_learns(learn, learns, dev, world, dist):
    """every learn timed on its own with CUDA events (max over ranks per learn); returns (median, all) in ms — single learns of these
    long cooperative launches occasionally run 10-30 % slow, which an average of two or three would carry into the record"""
    import torch
    each = []
    for _ in range(learns):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        learn()
        e1.record()
        torch.cuda.synchronize()
        each.append(e0.elapsed_time(e1))
    t = torch.tensor(each, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    each = [float(x) for x in t.tolist()]
    return float(np.median(each)), each


def ppo_dp_block(dev, rank, world, dist, learns=5):
    """BASELINE config 3 on every rank: PPO (LunarLander dims: obs 8, 4 actions), 1024 envs x 128 steps per GPU, minibatch
    8192 rows per GPU, K = 10 epochs = 160 optimiser steps per learn.  At world > 1 the replicas train data-parallel
    (PPO.enable_data_parallel: the flat gradient is summed across ranks between the in-kernel reduction and the clip /
    optimiser stages of every step), weak scaling: the union minibatch is world x 8192 rows.  Timed with CUDA events,
    max over ranks."""
    import contextlib
    import torch
    from freerl_b200.PPO import PPO
    T, N, mb, K = 128, 1024, 8192, 10
    rng = np.random.default_rng(100 + rank)
    with contextlib.redirect_stdout(sys.stderr):
        pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev, mode="fast")
    if world > 1:
        pol.enable_data_parallel()
    data = [(rng.standard_normal((N, 8), dtype=np.float32), rng.integers(0, 4, (N, 1)).astype(np.float32), rng.standard_normal(N).astype(np.float32),
             rng.standard_normal((N, 8), dtype=np.float32), rng.random(N) < 1 / 300, -np.abs(rng.standard_normal((N, 1))).astype(np.float32) * 0.1 - 1.3)
            for _ in range(T)]
    for o, a, r, o2, d, lp in data:
        pol.add(o, a, r, o2, d, lp, d)

    def learn():
        pol.buffer._index, pol.buffer._size, pol.buffer.n_envs = 0, T * N, N      # the rollout stays resident (learn() clears it)
        pol.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
    learn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, ms_each = _timed_learns(learn, learns, dev, world, dist)
    updates = K * (T * N // mb)
    # end to end through the public API with HOST buffers: T vector steps of select_action(host obs) + add(host arrays), then learn()
    torch.cuda.synchronize()
    e0.record()
    for o, a, r, o2, d, lp in data:
        act, logp = pol.select_action(o)                          # H2D obs, kernel, D2H actions + log-probs
        pol.add(o, act.reshape(N, 1), r, o2, d, logp.reshape(N, 1), d)
    pol.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
    loss = float(pol.last_metrics[-1, 0].item())
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e.item())
    return {"workload": "ppo_lunarlander_1024env_x128_mb8192_k10", "ms_per_learn": ms, "ms_per_learn_each": ms_each, "updates_per_learn": updates,
            "e2e_env_steps_per_sec": world * T * N / ms_e2e * 1e3, "e2e_ms_per_rollout_and_learn": ms_e2e, "e2e_last_loss": loss,
            "updates_per_sec": updates / ms * 1e3, "env_steps_per_sec": world * T * N / ms * 1e3, "us_per_update": ms * 1e3 / updates,
            "scaling": "weak", "collective": "none (1 rank)" if world == 1 else getattr(pol, "dp_collective", "nccl all_reduce of net.g per optimiser step"),
            "loss_finite": bool(torch.isfinite(pol.last_metrics).all().item())}


def rainbow_block(dev, rank, world, dist, vsteps=6):
    """BASELINE config 4 on every rank: Rainbow (PER + Noisy + C51 + N-step + Double + Dueling), LunarLander dims, 512 envs per GPU,
    PER capacity 1e6 per GPU (its own sum-tree), batch 256, one learn per env step (512 per vector step).  Multi-GPU = replicas with
    per-GPU env / PER shards and a parameter average per vector step (ReplicaSyncMixin), weak scaling."""
    import contextlib
    import torch
    from freerl_b200.DQN_with_tricks import DQN as Rainbow
    NE = 512
    trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
    rng = np.random.default_rng(200 + rank)
    with contextlib.redirect_stdout(sys.stderr):
        pol = Rainbow([8, 4], False, 1e-3, 1e6, dev, trick=trick, gamma=0.99, batch_size=256, mode="fast")
    if world > 1:
        pol.enable_replica_sync()
    obs = rng.standard_normal((NE, 8))

    def vstep():
        nonlocal obs
        act = pol.select_action(obs)
        nxt = rng.standard_normal((NE, 8))
        pol.add(obs, np.asarray(act).reshape(NE, 1), rng.standard_normal(NE), nxt, rng.random(NE) < 0.01)
        obs = nxt
        for _ in range(NE):
            pol.learn(256, 0.99, 0.01)
        pol.sync_replicas()
    for _ in range(8):                      # fill past the n-step windows and the first batch
        act = pol.select_action(obs)
        nxt = rng.standard_normal((NE, 8))
        pol.add(obs, np.asarray(act).reshape(NE, 1), rng.standard_normal(NE), nxt, rng.random(NE) < 0.01)
        obs = nxt
    vstep()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(vsteps):
        vstep()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / vsteps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    return {"workload": "rainbow_lunarlander_512env_per1e6_b256_utd1", "ms_per_vector_step": ms, "env_steps_per_sec": world * NE / ms * 1e3,
            "updates_per_sec": world * NE / ms * 1e3, "scaling": "weak", "e2e": "host observations / transitions in, actions out, every step",
            "parallelism": ("replicas, env + PER shard per GPU, parameter average per vector step: " + getattr(pol, "replica_collective", "?")) if world > 1 else "1 rank"}


def mappo_block(dev, rank, world, dist, learns=3):
    """BASELINE config 5 on every rank: MAPPO, simple_spread dims (3 agents, obs 18, act 5), 512 envs x 256 steps per GPU, full-batch
    minibatch (131 072 rows), K = 15.  Multi-GPU = synchronous data parallel over the in-kernel peer exchange, weak scaling."""
    import contextlib
    import torch
    from freerl_b200.MAPPO import MAPPO
    H, E, K5 = 256, 512, 15
    trick = {'adv_norm': True, 'ObsNorm': True, 'reward_norm': False, 'reward_scaling': True, 'orthogonal_init': True,
             'adam_eps': True, 'lr_decay': False, 'ValueClip': True, 'huber_loss': True, 'LayerNorm': True, 'feature_norm': True}
    ids = ["agent_%d" % i for i in range(3)]
    rng = np.random.default_rng(300 + rank)
    with contextlib.redirect_stdout(sys.stderr):
        pol = MAPPO({k: [18, 5] for k in ids}, True, 1e-3, 1e-3, H * E, dev, dict(trick), mode="fast")
    if world > 1:
        pol.enable_data_parallel()
    step = {k: (rng.standard_normal((E, 18), dtype=np.float32), rng.uniform(-1, 1, (E, 5)).astype(np.float32), rng.standard_normal(E).astype(np.float32),
                rng.standard_normal((E, 18), dtype=np.float32), np.zeros(E, bool), (-np.abs(rng.standard_normal((E, 5))) * 0.1 - 0.9).astype(np.float32))
            for k in ids}
    for t in range(H):
        trunc = np.full(E, (t % 25) == 24)
        pol.add({k: step[k][0] for k in ids}, {k: step[k][1] for k in ids}, {k: step[k][2] for k in ids}, {k: step[k][3] for k in ids},
                {k: step[k][4] for k in ids}, {k: step[k][5] for k in ids}, {k: trunc for k in ids})

    def learn():
        for b in pol.buffers.values():
            b._index, b._size, b.n_envs = 0, H * E, E
        pol.learn(H * E, 0.95, 0.95, 0.2, K5, 0.01, 10.0)
    learn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, ms_each = _timed_learns(learn, learns, dev, world, dist)
    return {"workload": "mappo_simple_spread_3agents_512env_x256_fullbatch_k15", "ms_per_learn": ms, "ms_per_learn_each": ms_each, "updates_per_learn": 3 * K5,
            "env_steps_per_sec": world * H * E / ms * 1e3, "scaling": "weak",
            "collective": "none (1 rank)" if world == 1 else getattr(pol, "dp_collective", "?"),
            "loss_finite": bool(torch.isfinite(pol.last_metrics).all().item())}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)      # 30 x 84 ms: a 2.5 s timed region
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip reference_cuda / extra.configs / ppo_dp (profiling runs)")
    args = ap.parse_args()
    # stdout carries ONE JSON line: keep the real stdout aside and point fd 1 at stderr, so prints from libraries (NCCL's
    # version banner, reference-style constructors) cannot get in front of it
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "sac_halfcheetah_256env_1Mreplay", "obs_dim": OBS, "act_dim": ACT, "hidden": [128, 128],
              "envs_per_gpu": NENV, "replay_capacity_per_gpu": CAP, "batch": BATCH, "updates_per_env_step": 1,
              "vector_steps_per_step": VSTEPS,
              "l2_note": "batches are uniform random rows of a 176 MB replay (> 126 MB L2)",
              "parallelism": "dp%d (env+replay shards per GPU, parameter average per vector step)" % max(world, 1)}

    if args.impl == "reference":
        if rank != 0:
            return
        w = max(args.warmup, 1)
        val, ms, cores, tps = cpu_reference_arm(args.steps, w, transitions_per_step=16)
        line = json.dumps({
            "impl": "reference", "metric": "env_steps_per_sec", "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d single-env iterations (select_action + add + np.random.choice(1e6,256) + SAC learn B=256) per step" % tps},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(line, file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    from freerl_b200.SAC import SAC
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the reference-style constructor prints; stdout carries ONE JSON line
        pol = SAC([OBS, ACT], True, 1e-3, 1e-3, CAP, dev, trick={}, mode="fast")
    # pre-fill the per-GPU replay shard on the device
    g = torch.Generator(device=dev)
    g.manual_seed(99 + rank)
    chunk = 250_000
    for _ in range(CAP // chunk):
        pol.buffer.add_device(torch.randn((chunk, OBS), device=dev, generator=g), torch.rand((chunk, ACT), device=dev, generator=g) * 2 - 1,
                              torch.randn(chunk, device=dev, generator=g), torch.randn((chunk, OBS), device=dev, generator=g),
                              (torch.rand(chunk, device=dev, generator=g) < 0.001).float())
    POOL = 64
    obs_pool = torch.randn((POOL, NENV, OBS), device=dev, generator=g)
    rew_pool = torch.randn((POOL, NENV), device=dev, generator=g)
    zeros = torch.zeros(NENV, device=dev)
    from freerl_b200 import _common, _lib
    if world > 1:
        pol.enable_replica_sync()             # broadcast rank 0's parameters; sync_replicas() averages them afterwards
        config["parallelism"] += ": " + pol.replica_collective

    def sync_params():
        pol.sync_replicas()                   # no-op at world 1

    learn_ev = []

    def device_step(t, timed):
        for v in range(VSTEPS):
            device_vstep(t * VSTEPS + v, timed)

    def device_vstep(t, timed):
        obs = obs_pool[t % POOL]
        act = _common.infer(pol.agent._actor, obs, _lib.INFER_SAC_SAMPLE, dev, ACT, seed=pol._seed, counter=t)
        nxt, rew = obs_pool[(t + 1) % POOL], rew_pool[(t + 1) % POOL]
        pol.buffer.add_device(obs, act, rew, nxt, zeros)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        pol.learn(BATCH, GAMMA, TAU, n_updates=NENV)
        if timed:
            e1.record()
            learn_ev.append((e0, e1))
        sync_params()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for t in range(K):
            fn(W + t, True)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident value ------------------------------------------------------------------------
    for t in range(W):
        device_step(t, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = int(_lib.lib().frl_launch_count())
    ms_total = timed_region(device_step)
    launches = int(_lib.lib().frl_launch_count()) - launches0               # counted by the library at its launch sites
    clocks = sampler.stop() if rank == 0 else None
    learn_ms = float(np.mean([a.elapsed_time(b) for a, b in learn_ev]))      # one launch = NENV sequential learns
    value = world * NENV * VSTEPS * K / (ms_total / 1e3)

    # ---- end-to-end through the public API with host buffers ------------------------------------------
    env = SyntheticVecEnv(NENV, 7 + rank)
    state = {"obs": env.obs, "loss": 0.0}

    def host_step(t, timed):
        for v in range(VSTEPS):
            host_vstep()

    def host_vstep():
        obs = state["obs"]
        action = pol.select_action(obs)                              # H2D obs, kernel, D2H actions
        nxt, rew, term, trunc = env.step(action)
        pol.add(obs, action, rew, nxt, term)                         # H2D packed transitions
        pol.learn(BATCH, GAMMA, TAU, n_updates=NENV)
        state["loss"] = float(pol.last_metrics[NENV - 1, 0].item())  # D2H read of the step's result
        state["obs"] = nxt
        sync_params()
    for t in range(W):
        host_step(t, False)
    ms_e2e = timed_region(host_step)
    e2e = world * NENV * VSTEPS * K / (ms_e2e / 1e3)
    h2d = VSTEPS * (NENV * OBS * 4 + NENV * pol.buffer.row_floats * 4)
    d2h = VSTEPS * (NENV * ACT * 4 + 4)

    # ---- on-policy data parallel (config 3 shape) at this N: every rank runs it, rank 0 reports ------------------------
    ppo_dp = None
    multi = {}
    if not args.no_extras:
        try:
            ppo_dp = ppo_dp_block(dev, rank, world, dist)
        except Exception as e:                                   # the headline line must survive a failure of an extra
            ppo_dp = {"error": "%s: %s" % (type(e).__name__, e)}
        for name, fn in (("rainbow_replicas", rainbow_block), ("mappo_dp", mappo_block)):      # configs 4 and 5 at this N
            try:
                multi[name] = fn(dev, rank, world, dist)
            except Exception as e:
                multi[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_LEARN * NENV / (learn_ms / 1e3) / 1e9
    out = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "updates_per_sec": world * NENV * VSTEPS * K / (ms_total / 1e3),
        "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "last_loss": state["loss"]},
        "gpu_launches": launches,   # frl_launch_count() delta over the timed region (per vector step: policy-infer, replay add_batch, fused 256-update learn)
        "clocks": clocks,
        "roofline": {"kernel": "frl_fx_kernel<AcFx> (fused SAC learn x%d per launch)" % NENV, "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_LEARN * NENV, "launch_ms": learn_ms,
                     "traffic": ncu_traffic(), "traffic_source": NCU_CAPTURE + " (stored ncu --set full capture of this command; not measured in this run)",
                     "achieved_tflops_fp32": ALGO_FLOPS_PER_LEARN * NENV / (learn_ms / 1e3) / 1e12,
                     "note": "the fused update is FLOP/latency-bound at B=256 (SURVEY §7.3-1): params/Adam/targets are L2-resident"},
    }
    if not args.no_cpu_baseline and world == 1:
        val, ms, cores, tps = cpu_reference_arm(150, 1)
        out["cpu_baseline"] = {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                               "sample": "150 steps x %d single-env iterations of the oracle port (np.random.choice over the full 1e6 replay + SAC learn B=256)" % tps}
    if ppo_dp is not None:
        out["ppo_dp"] = ppo_dp
    out.update(multi)
    if not args.no_extras and world == 1:
        extra = {}
        try:       # stock PyTorch on this same B200: the oracle port's torch code with the networks on cuda
            v1, _, _, tps = cpu_reference_arm(12, 2, transitions_per_step=16, device="cuda")
            v2, _, _, _ = cpu_reference_arm(12, 2, transitions_per_step=16, device="cuda", choice=False)
            out["reference_cuda"] = {"value": v1, "value_without_np_random_choice": v2, "unit": "env-steps/s", "kind": "port",
                                     "sample": "12 steps x %d single-env iterations; numpy replay on the host, networks + learn on cuda:0" % tps}
        except Exception as e:
            out["reference_cuda"] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:       # the other BASELINE configurations (learn() only), in a fresh process: tools/configbench.py, one GPU, after this one is idle
            del pol, obs_pool, rew_pool
            torch.cuda.empty_cache()
            tmp = os.path.join(ROOT, "gpurun_out", "bench_configs.json")
            os.makedirs(os.path.dirname(tmp), exist_ok=True)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "configbench.py"), "--only", "C1,C3,C4,C5", "--json", tmp],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=600,
                               env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local))))
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-300:])
            extra["configs"] = json.load(open(tmp))["rows"]
        except Exception as e:
            extra["configs_error"] = "%s: %s" % (type(e).__name__, e)
        out["extra"] = extra
    print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
