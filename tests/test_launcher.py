"""The reference train scripts run UNCHANGED on top of freerl_b200 through the launcher (only where the reference tree
is present — the build container; skipped on the GPU box)."""
import os

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "DQN_file")), reason="reference tree not present")


def test_reference_dqn_script_unchanged(tmp_path, emul):
    from freerl_b200 import launcher
    from freerl_b200.DQN import DQN
    ns = launcher.run_reference_script(os.path.join(REF, "DQN_file", "DQN.py"),
                                       ["--env_name", "CartPole-v1", "--max_episodes", "3", "--start_steps", "40", "--batch_size", "32",
                                        "--buffer_size", "2000", "--device", "cpu"], results_root=str(tmp_path))
    assert isinstance(ns["policy"], DQN) and ns["episode_num"] == 3
    assert ns["policy"].agent.step > 0                                   # learn() ran on the fused kernel
    model_dir = ns["model_dir"]
    assert os.path.exists(os.path.join(model_dir, "DQN.pt")) and os.path.exists(os.path.join(model_dir, "DQN_seed_0.npy"))
    sd = torch.load(os.path.join(model_dir, "DQN.pt"))
    assert list(sd.keys()) == ["l1.weight", "l1.bias", "l2.weight", "l2.bias"] and sd["l1.weight"].shape == (128, 4)


def test_reference_sac_and_ppo_scripts_unchanged(tmp_path, emul):
    from freerl_b200 import launcher
    ns = launcher.run_reference_script(os.path.join(REF, "SAC_file", "SAC.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "1", "--start_steps", "60", "--random_steps", "20",
                                        "--batch_size", "32", "--buffer_size", "1000", "--device", "cpu"], results_root=str(tmp_path))
    assert ns["policy"].agent.critic_step > 0 and np.isfinite(float(ns["policy"].alphas.alpha))
    ns = launcher.run_reference_script(os.path.join(REF, "PPO_file", "PPO.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "2", "--horizon", "128", "--minibatch_size", "32",
                                        "--K_epochs", "2", "--device", "cpu"], results_root=str(tmp_path))
    assert ns["policy"].agent.step == 3 * 2 * 4 or ns["policy"].agent.step > 0
