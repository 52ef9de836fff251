"""Data-parallel PPO over torch.distributed (gloo, world_size 2, CPU emulation of the kernels): two ranks, each with half
of every minibatch, must end with the parameters a single process gets on the union minibatches."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["FRL_ROOT"], "tests"))
from parity_util import load_into, net_from_golden
from freerl_b200.PPO import PPO
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(os.path.join(os.environ["FRL_ROOT"], "tests", "golden", "ppo_cont.npz"))
H = 256 // world
pol = PPO([8, 2], True, 1e-3, 1e-3, H, torch.device("cpu"))
load_into(pol.agent.actor, net_from_golden(g, "init/actor/"))
load_into(pol.agent.critic, net_from_golden(g, "init/critic/"))
pol.enable_data_parallel()
d = [g["data/" + k] for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done")]
sl = slice(rank * H, (rank + 1) * H)                      # this rank's contiguous time shard (its own "env")
for t in range(H):
    i = rank * H + t
    pol.add(d[0][i], d[1][i], float(d[2][i, 0]), d[3][i], bool(d[4][i, 0]), d[5][i], bool(d[6][i, 0]))
rng = np.random.default_rng(5)
perms = [rng.permutation(H) for _ in range(2)]           # same local permutation on both ranks
pol.learn(32, 0.99, 0.95, 0.2, 2, 0.01, permutations=perms)
if True:
    sd = {k: v.cpu().numpy() for k, v in pol.agent.actor.state_dict().items()}
    sd.update({"critic." + k: v.cpu().numpy() for k, v in pol.agent.critic.state_dict().items()})
    np.savez(os.path.join(os.environ["FRL_OUT"], "rank%d.npz" % rank), **sd)
dist.destroy_process_group()
'''


def test_ppo_data_parallel_gloo(tmp_path, emul):
    (tmp_path / "worker.py").write_text(WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path), FREERL_B200_LIB=os.environ["FREERL_B200_LIB"], OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), "replicas diverged: " + k           # bit-identical replicas
    # single-process oracle on the union minibatches (per-rank GAE shards are independent time segments)
    from collections import OrderedDict
    from oracle import algos
    from parity_util import net_from_golden
    g = np.load(os.path.join(ROOT, "tests", "golden", "ppo_cont.npz"))
    orc = algos.PPOOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, True)
    data = tuple(torch.from_numpy(g["data/" + k]) for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
    H = 128
    advs, vts = [], []
    for rk in range(2):
        shard = tuple(x[rk * H:(rk + 1) * H] for x in data)
        a, v = orc.advantages(shard, 0.99, 0.95)
        advs.append(a); vts.append(v)
    adv, vt = torch.cat(advs), torch.cat(vts)
    rng = np.random.default_rng(5)
    perms = [rng.permutation(H) for _ in range(2)]
    for perm in perms:
        for s in range(0, H, 32):
            loc = perm[s:s + 32]
            orc.minibatch(data, adv, vt, np.concatenate([loc, loc + H]), 0.2, 0.01)
    for k, v in orc.actor.items():
        np.testing.assert_allclose(r0[k], v.detach().numpy(), rtol=2e-4, atol=2e-5, err_msg=k)
    for k, v in orc.critic.items():
        np.testing.assert_allclose(r0["critic." + k], v.detach().numpy(), rtol=2e-4, atol=2e-5, err_msg=k)


REPLICA_WORKER = r'''
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"])
from freerl_b200.SAC import SAC
dist.init_process_group("gloo")
rank = dist.get_rank()
torch.manual_seed(10 + rank); np.random.seed(10 + rank)           # different initial replicas on purpose
pol = SAC([5, 2], True, 1e-3, 1e-3, 512, torch.device("cpu"), trick={}, mode="fast")
rng = np.random.default_rng(rank)                                  # different replay shards
pol.add(rng.standard_normal((200, 5)), rng.uniform(-1, 1, (200, 2)), rng.standard_normal(200), rng.standard_normal((200, 5)), rng.random(200) < 0.1)
pol.enable_replica_sync()
p0 = pol.agent._critic.p.clone()
pol.learn(32, 0.99, 0.01, n_updates=3)
local = {n: getattr(pol.agent, n).p.clone() for n in ("_actor", "_critic", "_actor_t", "_critic_t")}
la_local = pol.alphas.state[0].clone()
pol.sync_replicas()
np.savez(os.path.join(os.environ["FRL_OUT"], "rep%d.npz" % rank), start=p0.numpy(), la_local=la_local.numpy(), la=pol.alphas.state[0].numpy(),
         **{"local" + n: v.numpy() for n, v in local.items()}, **{"synced" + n: getattr(pol.agent, n).p.numpy() for n in local},
         mirror=pol.agent._critic.pt.numpy())
dist.destroy_process_group()
'''


def test_offpolicy_replica_sync_gloo(tmp_path, emul):
    """SAC replicas with different shards: enable_replica_sync() starts both from rank 0's parameters, local learns diverge,
    sync_replicas() leaves both with the mean of the two (parameters, targets, log_alpha) and refreshed mirrors."""
    (tmp_path / "worker.py").write_text(REPLICA_WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path), FREERL_B200_LIB=os.environ["FREERL_B200_LIB"], OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    a, b = np.load(tmp_path / "rep0.npz"), np.load(tmp_path / "rep1.npz")
    assert np.array_equal(a["start"], b["start"])                                   # broadcast from rank 0
    for n in ("_actor", "_critic", "_actor_t", "_critic_t"):
        assert not np.array_equal(a["local" + n], b["local" + n])                   # shards differ -> replicas diverged
        assert np.array_equal(a["synced" + n], b["synced" + n])
        np.testing.assert_allclose(a["synced" + n], (a["local" + n] + b["local" + n]) / 2, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(a["la"], (a["la_local"] + b["la_local"]) / 2, rtol=1e-6)
    assert np.array_equal(a["mirror"], b["mirror"])


MAPPO_WORKER = r'''
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["FRL_ROOT"], "tests"))
from freerl_b200.MAPPO import MAPPO
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
trick = {'adv_norm': True, 'ObsNorm': True, 'reward_norm': False, 'reward_scaling': True, 'orthogonal_init': True,
         'adam_eps': True, 'lr_decay': False, 'ValueClip': True, 'huber_loss': True, 'LayerNorm': True, 'feature_norm': True}
ids = ["agent_%d" % i for i in range(2)]
H, E = 8, 2
torch.manual_seed(3 + rank)                      # different initial replicas: enable_data_parallel must broadcast rank 0's
pol = MAPPO({k: [6, 3] for k in ids}, True, 1e-3, 1e-3, H * E, torch.device("cpu"), dict(trick))
pol.enable_data_parallel()
rng = np.random.default_rng(40 + rank)           # every rank steps its own envs
for t in range(H):
    obs = {k: rng.standard_normal((E, 6), dtype=np.float32) for k in ids}
    act = {k: rng.uniform(-1, 1, (E, 3)).astype(np.float32) for k in ids}
    lp = {k: (-np.abs(rng.standard_normal((E, 3))) - 0.5).astype(np.float32) for k in ids}
    rew = {k: rng.standard_normal(E).astype(np.float32) * (1 + 2 * rank) for k in ids}
    nobs = {k: rng.standard_normal((E, 6), dtype=np.float32) for k in ids}
    pol.add(obs, act, rew, nobs, {k: np.zeros(E, bool) for k in ids}, lp, {k: np.full(E, t == H - 1) for k in ids})
pol.trick['adv_norm'] = False
raw, _, _ = pol.compute_advantages(0.95, 0.95)
pol.trick['adv_norm'] = True
adv, _, _ = pol.compute_advantages(0.95, 0.95)
prng = np.random.default_rng(9)
perms = {k: [prng.permutation(H * E) for _ in range(2)] for k in ids}      # same local permutations on both ranks
pol.learn(8, 0.95, 0.95, 0.2, 2, 0.01, 10.0, permutations=perms)
out = {"raw": raw.cpu().numpy(), "adv": adv.cpu().numpy()}
for k in ids:
    out["p." + k] = pol.agents[k]._net.p.cpu().numpy()
np.savez(os.path.join(os.environ["FRL_OUT"], "mappo_rank%d.npz" % rank), **out)
dist.destroy_process_group()
'''


def test_mappo_data_parallel_adv_norm_gloo(tmp_path, emul):
    """MAPPO under data parallel with the adv_norm trick (the class default): the advantages are normalised with the statistics of the
    UNION of the ranks' rollouts (all-reduced sum / sum of squares / count), and the replicas stay bit-identical through the learn."""
    (tmp_path / "worker.py").write_text(MAPPO_WORKER)
    env = dict(os.environ, FRL_ROOT=ROOT, FRL_OUT=str(tmp_path), FREERL_B200_LIB=os.environ["FREERL_B200_LIB"], OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r0, r1 = np.load(tmp_path / "mappo_rank0.npz"), np.load(tmp_path / "mappo_rank1.npz")
    union = torch.from_numpy(np.concatenate([r0["raw"], r1["raw"]]))
    want = ((union - union.mean()) / (union.std() + 1e-8)).numpy()          # MAPPO.py:386-388 over the whole (union) rollout
    n = r0["raw"].shape[0]
    np.testing.assert_allclose(r0["adv"], want[:n], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r1["adv"], want[n:], rtol=1e-5, atol=1e-6)
    assert abs(float(r0["adv"].mean())) > 1e-3                               # a per-shard normalisation would have centred each shard
    for k in r0.files:
        if k.startswith("p."):
            assert np.array_equal(r0[k], r1[k]), "replicas diverged: " + k
