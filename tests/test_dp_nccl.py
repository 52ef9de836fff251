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
