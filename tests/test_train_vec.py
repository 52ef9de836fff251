"""freerl_b200.train_vec: the vectorised off-policy loop (N envs on the host, batched inference / add / fused learns on the device)
runs end to end for SAC, TD3 and DQN on the synthetic shape-only envs (on the host emulation: it only composes entry points that
have their own GPU parity tests)."""
import numpy as np
import pytest
import torch


def _run(device, tmp_path):
    from freerl_b200 import train_vec
    common = ["--n_envs", "8", "--total_steps", "160", "--random_steps", "40", "--start_steps", "64", "--batch_size", "16",
              "--buffer_size", "500", "--log_every", "0", "--device", str(device)]
    r = train_vec.main(["--algo", "SAC", "--env_name", "Pendulum-v1", "--obs_norm", "--save_dir", str(tmp_path / "sac")] + common)
    assert r["steps"] == 160 and r["learns"] == 104 and (tmp_path / "sac" / "SAC.pt").exists()       # UTD 1 after start_steps
    assert len(r["policy"].buffer) == 160 and np.isfinite(r["policy"].last_metrics.cpu().numpy()[:, :2]).all()
    r = train_vec.main(["--algo", "TD3", "--env_name", "Pendulum-v1", "--updates_per_step", "0.5"] + common)
    assert r["learns"] == 52 and r["policy"].total_it == 52
    r = train_vec.main(["--algo", "DQN", "--env_name", "CartPole-v1"] + common)
    assert r["learns"] == 104 and len(r["returns"]) >= 0
    r = train_vec.main(["--algo", "RAINBOW", "--env_name", "CartPole-v1", "--updates_per_step", "0.25"] + common)
    assert r["learns"] == 26 and len(r["policy"].buffer) > 100          # the n-step windows hold back the newest steps of every env
    for env in ("Pendulum-v1", "CartPole-v1"):                                             # Gaussian and Categorical heads
        r = train_vec.main(["--algo", "PPO", "--env_name", env, "--n_envs", "8", "--horizon", "16", "--total_steps", "256", "--minibatch_size", "32",
                            "--K_epochs", "2", "--log_every", "0", "--device", str(device)])
        assert r["steps"] == 256 and r["learns"] == 2 and r["policy"].agent.step == 2 * 2 * 4
    r = train_vec.main(["--algo", "MAPPO", "--env_name", "simple_spread_v3", "--n_agents", "3", "--n_envs", "4", "--horizon", "30", "--total_steps", "240",
                        "--minibatch_size", "60", "--K_epochs", "2", "--log_every", "0", "--device", str(device)])
    assert r["steps"] == 240 and r["learns"] == 2 and len(r["returns"]) == 4 * 2          # max_cycles 25: two finished episodes per env
    with pytest.raises(ValueError, match="action space"):
        train_vec.main(["--algo", "DQN", "--env_name", "Pendulum-v1"] + common)


def test_train_vec_emulated(emul, tmp_path):
    _run(torch.device("cpu"), tmp_path)
    ms = np.load(tmp_path / "sac" / "SAC_running_mean_std.npy")          # --obs_norm statistics in the reference's layout (SAC.py:583-584)
    assert ms.shape == (2, 3) and np.isfinite(ms).all()


@pytest.mark.gpu
def test_train_vec_gpu(tmp_path):
    """the same loops through the CUDA kernels (SAC / TD3 / DQN / Rainbow / PPO both heads / MAPPO) on the device"""
    _run(torch.device("cuda"), tmp_path)


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FRL_ROOT"])
from freerl_b200 import train_vec
r = train_vec.main(["--algo", "SAC", "--env_name", "Pendulum-v1", "--n_envs", "4", "--total_steps", "96", "--random_steps", "16", "--start_steps", "32",
                    "--batch_size", "8", "--buffer_size", "300", "--log_every", "0", "--device", "cpu"])
pol = r["policy"]
q = train_vec.main(["--algo", "PPO", "--env_name", "Pendulum-v1", "--n_envs", "4", "--horizon", "16", "--total_steps", "128", "--minibatch_size", "16",
                    "--K_epochs", "2", "--log_every", "0", "--device", "cpu"])
rb = train_vec.main(["--algo", "RAINBOW", "--env_name", "CartPole-v1", "--n_envs", "4", "--total_steps", "96", "--random_steps", "16", "--start_steps", "48",
                     "--batch_size", "8", "--buffer_size", "300", "--updates_per_step", "0.25", "--log_every", "0", "--device", "cpu"])
mp = train_vec.main(["--algo", "MAPPO", "--env_name", "simple_spread_v3", "--n_agents", "3", "--n_envs", "2", "--horizon", "30", "--total_steps", "120",
                     "--minibatch_size", "30", "--K_epochs", "2", "--log_every", "0", "--device", "cpu"])
np.savez(os.path.join(os.environ["FRL_OUT"], "tv%d.npz" % r["rank"]), actor=pol.agent._actor.p.numpy(), critic=pol.agent._critic.p.numpy(),
         first_obs=pol.buffer.obs[0].numpy(), learns=r["learns"], world=r["world"], ppo=q["policy"].agent._net.p.numpy(),
         ppo_steps=q["policy"].agent.step, rainbow=rb["policy"].agent.online.p.numpy(), rainbow_target=rb["policy"].agent.target.p.numpy(),
         rainbow_learns=rb["learns"], rainbow_tree_total=float(rb["policy"].buffer.tree.total_priority) if hasattr(rb["policy"].buffer, "tree") else 0.0,
         mappo=np.concatenate([ag._net.p.numpy() for ag in mp["policy"].agents.values()]), mappo_learns=mp["learns"])
dist.destroy_process_group()
'''


def test_train_vec_two_processes_gloo(tmp_path, emul):
    """torchrun x 2 (gloo, host emulation): each rank steps its own envs into its own replay shard (different data), the replicas
    end every vector step with the same averaged parameters."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "worker.py").write_text(WORKER)
    env = dict(os.environ, FRL_ROOT=root, FRL_OUT=str(tmp_path), FREERL_B200_LIB=os.environ["FREERL_B200_LIB"], OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(tmp_path / "worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    a, b = np.load(tmp_path / "tv0.npz"), np.load(tmp_path / "tv1.npz")
    assert int(a["world"]) == 2 and int(a["learns"]) == int(b["learns"]) == 68
    assert not np.array_equal(a["first_obs"], b["first_obs"])                       # different env shards
    assert np.array_equal(a["actor"], b["actor"]) and np.array_equal(a["critic"], b["critic"])
    # PPO: gradient all-reduce inside every minibatch step -> bit-identical replicas although the rollouts differ
    assert int(a["ppo_steps"]) == 2 * 2 * 4 and np.array_equal(a["ppo"], b["ppo"])
    # Rainbow (BASELINE config 4): replicas with their own env / PER shards (sum-trees differ), parameters averaged per vector step
    assert int(a["rainbow_learns"]) == int(b["rainbow_learns"]) > 0
    assert np.array_equal(a["rainbow"], b["rainbow"]) and np.array_equal(a["rainbow_target"], b["rainbow_target"])
    # MAPPO (config 5): synchronous data parallel over the ranks' env shards -> bit-identical replicas of every agent
    assert int(a["mappo_learns"]) == 2 and np.array_equal(a["mappo"], b["mappo"])
