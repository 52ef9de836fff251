"""Data-parallel PPO over NCCL on two real GPUs (skipped on a one-GPU box): the flat-gradient all-reduce between the
in-kernel reduce and optimiser stages must leave both replicas bit-identical and equal to the single-process oracle on the
union minibatches (same check as tests/test_dp_gloo.py, through the CUDA kernels instead of the host emulation)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_dp_gloo import ROOT, WORKER

NCCL_WORKER = (WORKER.replace('dist.init_process_group("gloo")',
                              'torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))\n'
                              'dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))')
               .replace('torch.device("cpu")', 'torch.device("cuda", int(os.environ["LOCAL_RANK"]))'))


@pytest.mark.gpu
@pytest.mark.parametrize("peer", ["1", "0"])
def test_ppo_data_parallel_nccl(tmp_path, peer):
    """peer = 1: gradients summed inside the one persistent launch over peer-mapped NVLink memory (the default on CUDA);
    peer = 0: host-driven launch / dist.all_reduce(net.g) / launch per optimiser step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    (tmp_path / "worker.py").write_text(NCCL_WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path), FREERL_B200_DP_PEER=peer)
    env.pop("FREERL_B200_LIB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541" if peer == "1" else "29542", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), "replicas diverged: " + k
    from oracle import algos
    from parity_util import net_from_golden
    g = np.load(os.path.join(ROOT, "tests", "golden", "ppo_cont.npz"))
    orc = algos.PPOOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, True)
    data = tuple(torch.from_numpy(g["data/" + k]) for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
    H = 128
    advs, vts = [], []
    for rk in range(2):
        a, v = orc.advantages(tuple(x[rk * H:(rk + 1) * H] for x in data), 0.99, 0.95)
        advs.append(a); vts.append(v)
    adv, vt = torch.cat(advs), torch.cat(vts)
    rng = np.random.default_rng(5)
    for perm in [rng.permutation(H) for _ in range(2)]:
        for s in range(0, H, 32):
            loc = perm[s:s + 32]
            orc.minibatch(data, adv, vt, np.concatenate([loc, loc + H]), 0.2, 0.01)
    for k, v in orc.actor.items():
        np.testing.assert_allclose(r0[k], v.detach().numpy(), rtol=2e-4, atol=2e-5, err_msg=k)
    for k, v in orc.critic.items():
        np.testing.assert_allclose(r0["critic." + k], v.detach().numpy(), rtol=2e-4, atol=2e-5, err_msg=k)


REPLICA_WORKER = r'''
import contextlib, io, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["FRL_ROOT"], "tests"))
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from freerl_b200.SAC import SAC
from test_learning_sanity import PointEnv
torch.manual_seed(rank); np.random.seed(rank)
dev = torch.device("cuda", rank)
with contextlib.redirect_stdout(io.StringIO()):
    pol = SAC([2, 1], True, 1e-3, 1e-3, 100_000, dev, trick={}, mode="fast")
pol.enable_replica_sync()
assert pol._rs_peers is not None, pol.replica_collective      # the in-kernel peer-memory average, not the all-reduce fallback
env = PointEnv(64, 100 + rank)              # every rank steps its OWN env shard and fills its OWN replay shard
obs, rets = env.obs(), []
for step in range(500):
    act = pol.select_action(obs) if step >= 20 else np.random.uniform(-1, 1, (64, 1)).astype(np.float32)
    nxt, r, term, trunc = env.step(act)
    pol.add(obs, act, r, nxt, term)
    obs = env.obs()
    rets.append(float(r.mean()))
    if step >= 20:
        pol.learn(256, 0.95, 0.01, n_updates=16)
        pol.sync_replicas()                 # parameter average per vector step, like bench.py / train_vec
assert float(pol._rs_peers.status.item()) == 0.0             # no peer ever timed out
sd = {k: v.detach().cpu().numpy() for k, v in pol.agent.actor.state_dict().items()}
np.savez(os.path.join(os.environ["FRL_OUT"], "replica%d.npz" % rank), rets=np.array(rets), **sd)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_sac_replicas_learn_point_env_nccl(tmp_path):
    """replicas with sharded envs / replay and a parameter average per vector step are a different algorithm from
    the single-process reference (local Adam moments) — so beyond bit-level checks, the N = 2 run must LEARN like the N = 1 run of
    tests/test_learning_sanity.py (same thresholds) and leave both replicas with identical parameters."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    (tmp_path / "worker.py").write_text(REPLICA_WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path))
    env.pop("FREERL_B200_LIB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r0, r1 = np.load(tmp_path / "replica0.npz"), np.load(tmp_path / "replica1.npz")
    for k in r0.files:
        if k != "rets":
            assert np.array_equal(r0[k], r1[k]), "replicas differ after the final average: " + k
    for rk, rr in enumerate((r0, r1)):
        rets = rr["rets"]
        random_phase, late = rets[:20].mean(), rets[-100:].mean()
        print("rank %d: random phase %.3f, last 100 steps %.3f" % (rk, random_phase, late))
        assert late > random_phase + 0.3 and late > -0.2, (rk, random_phase, late)


AVG_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"])
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from freerl_b200 import _common
sizes = [36868, 39426, 3, 1, 130]                   # parameter-block sized tensors, odd tails, a scalar
g = torch.Generator().manual_seed(100 + rank)
ts = [torch.randn(n, generator=g).to(dev) for n in sizes]
peers = _common.replica_peers(dist, None, dev, ts)
assert peers is not None
outs = {}
for it in range(3):                                 # three syncs: both halves of the double buffer and the epoch hand-off
    before = [t.cpu().numpy().copy() for t in ts]
    _common.replica_average(peers, ts, dev)
    torch.cuda.synchronize()
    for i, t in enumerate(ts):
        outs["in%d_%d" % (it, i)] = before[i]
        outs["out%d_%d" % (it, i)] = t.cpu().numpy().copy()
    for t in ts:                                    # drift apart again before the next sync
        t.add_(torch.randn(t.shape, generator=g).to(dev) * 0.1)
assert float(peers.status.item()) == 0.0
np.savez(os.path.join(os.environ["FRL_OUT"], "avg%d.npz" % rank), **outs)
peers.close(dist)                                   # unmap the peers' blocks, free the own one (barriers inside)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_replica_average_peer_memory_nccl(tmp_path):
    """frl_replica_average on two GPUs: every rank ends with (x_0 + x_1) * 0.5 in rank order — bit-identical across the ranks and equal
    to the host computation, over three successive syncs (both buffer halves)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    (tmp_path / "worker.py").write_text(AVG_WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path))
    env.pop("FREERL_B200_LIB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29544", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r0, r1 = np.load(tmp_path / "avg0.npz"), np.load(tmp_path / "avg1.npz")
    for it in range(3):
        for i in range(5):
            want = (r0["in%d_%d" % (it, i)] + r1["in%d_%d" % (it, i)]) * np.float32(0.5)
            assert np.array_equal(r0["out%d_%d" % (it, i)], want), (it, i)
            assert np.array_equal(r1["out%d_%d" % (it, i)], want), (it, i)
