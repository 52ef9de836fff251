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
