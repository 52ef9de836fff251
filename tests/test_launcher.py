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


def test_reference_multi_agent_scripts_unchanged(tmp_path, emul):
    """MADDPG.py / MATD3_simple.py (off-policy, dict-in / dict-out PettingZoo loop) and MAPPO.py / IPPO.py / HAPPO.py (on-policy)
    run unchanged on the synthetic MPE shim; learn() runs on the fused kernels (optimiser step counters move)."""
    from freerl_b200 import launcher
    off = ["--env_name", "simple_spread_v3", "--N", "3", "--max_episodes", "4", "--start_steps", "60", "--batch_size", "32",
           "--buffer_size", "2000", "--device", "cpu"]
    ns = launcher.run_reference_script(os.path.join(REF, "MADDPG_file", "MADDPG.py"), off, results_root=str(tmp_path))
    pol = ns["policy"]
    assert type(pol).__module__ == "freerl_b200.MADDPG" and all(a.critic_step > 0 for a in pol.agents.values())
    assert os.path.exists(os.path.join(ns["model_dir"], "MADDPG.pth"))
    ns = launcher.run_reference_script(os.path.join(REF, "MADDPG_file", "MATD3_simple.py"), off, results_root=str(tmp_path))
    pol = ns["policy"]
    assert type(pol).__module__ == "freerl_b200.MATD3_simple" and pol.total_it > 0
    on = ["--env_name", "simple_spread_v3", "--N", "3", "--max_episodes", "6", "--horizon", "50", "--minibatch_size", "25",
          "--K_epochs", "2", "--device", "cpu"]
    for sub, cls in (("MAPPO.py", "MAPPO"), ("IPPO.py", "IPPO"), ("HAPPO.py", "HAPPO")):
        extra = ["--continuous_actions", "True"] if sub != "MAPPO.py" else []
        ns = launcher.run_reference_script(os.path.join(REF, "MAPPO_file", sub), on + extra, results_root=str(tmp_path))
        pol = ns["policy"]
        assert type(pol).__module__ == "freerl_b200." + cls and all(a.step > 0 for a in pol.agents.values()), sub
    # MAPPO_discrete.py: shared networks + episode ReplayBuffer; `horizon` counts EPISODES.  The script's DEFAULT switches (policy_name
    # MAPPO: ObsNorm, reward_scaling, adv_norm, orthogonal init, adam_eps, ValueClip + huber_loss, LayerNorm, feature_norm) and the
    # all-False set (MAPPO_simple) both run
    for extra in ([], ["--policy_name", "MAPPO_simple"]):
        ns = launcher.run_reference_script(os.path.join(REF, "MAPPO_file", "MAPPO_discrete.py"),
                                           ["--env_name", "simple_spread_v3", "--N", "3", "--max_episodes", "4", "--horizon", "2", "--minibatch_size", "1",
                                            "--K_epochs", "2", "--device", "cpu"] + extra, results_root=str(tmp_path))
        pol = ns["policy"]
        assert type(pol).__module__ == "freerl_b200.MAPPO_discrete" and pol.agent.step == 2 * 2 * 2 * 2      # 2 learns x 2 epochs x 2 minibatches x 2 steps
        assert bool(pol.trick["LayerNorm"]) == (not extra) and np.isfinite(pol.last_metrics.cpu().numpy()).all()
        assert os.path.exists(os.path.join(ns["model_dir"], "MAPPO_discrete.pth"))


def test_reference_single_agent_siblings_unchanged(tmp_path, emul):
    """TD3.py, Rainbow DQN_with_tricks.py (default trick set), DDPG_simple.py, MADDPG_simple.py and PPO_advance/PPO.py unchanged."""
    from freerl_b200 import launcher
    ns = launcher.run_reference_script(os.path.join(REF, "TD3_file", "TD3.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "1", "--start_steps", "60", "--batch_size", "32",
                                        "--buffer_size", "1000", "--device", "cpu"], results_root=str(tmp_path))
    assert ns["policy"].total_it > 0 and ns["policy"].agent.actor_step == ns["policy"].total_it // 2
    ns = launcher.run_reference_script(os.path.join(REF, "DQN_file", "DQN_with_tricks.py"),
                                       ["--env_name", "CartPole-v1", "--max_episodes", "2", "--start_steps", "40", "--batch_size", "16",
                                        "--buffer_size", "512", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__name__ == "_RainbowDQN" and ns["policy"].agent.step > 0
    ns = launcher.run_reference_script(os.path.join(REF, "DDPG_file", "DDPG_simple.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "1", "--start_steps", "60", "--batch_size", "32",
                                        "--buffer_size", "1000", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__module__ == "freerl_b200.DDPG_simple" and ns["policy"].agent.critic_step > 0
    ns = launcher.run_reference_script(os.path.join(REF, "MADDPG_file", "MADDPG_simple.py"),
                                       ["--env_name", "simple_spread_v3", "--N", "3", "--max_episodes", "4", "--start_steps", "60",
                                        "--batch_size", "32", "--buffer_size", "2000", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__module__ == "freerl_b200.MADDPG_simple"
    ns = launcher.run_reference_script(os.path.join(REF, "PPO_advance", "PPO.py"),
                                       ["--env_name", "CartPole-v1", "--max_episodes", "3", "--horizon", "64", "--minibatch_size", "32",
                                        "--K_epochs", "2", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__module__ == "freerl_b200.PPO_advance" and ns["policy"].agent.step > 0


def test_reference_ppo_with_tricks_script_unchanged(tmp_path, emul):
    """PPO_file/PPO_with_tricks.py: upstream its learn() raises (np.zeros with a torch dtype, :302); with the class rebound to ours the
    UNCHANGED script trains (default trick dict: everything off)."""
    from freerl_b200 import launcher
    ns = launcher.run_reference_script(os.path.join(REF, "PPO_file", "PPO_with_tricks.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "2", "--horizon", "128", "--minibatch_size", "32",
                                        "--K_epochs", "2", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__module__ == "freerl_b200.PPO_with_tricks" and ns["policy"].agent.step > 0
    for d, f, mod in (("PPO_advance", "PPO_with_tricks.py", "freerl_b200.PPO_with_tricks"), ("PPO_advance", "PPO_cc.py", "freerl_b200.PPO_advance")):
        ns = launcher.run_reference_script(os.path.join(REF, d, f),
                                           ["--env_name", "Pendulum-v1", "--max_episodes", "2", "--horizon", "128", "--minibatch_size", "32",
                                            "--K_epochs", "2", "--device", "cpu"], results_root=str(tmp_path))
        assert type(ns["policy"]).__module__ == mod and ns["policy"].agent.step > 0, f


def test_sac_add_discrete_continuous_is_sac(tmp_path, emul):
    """SAC_file/SAC_add_discrete.py with a continuous action space is SAC.py statement for statement: the unmodified reference class,
    driven exactly like oracle/make_golden.py::gen_sac, reproduces tests/golden/sac.npz bit for bit — which is what justifies
    freerl_b200.SAC_add_discrete.SAC delegating to the SAC kernel path.  Then the unchanged script runs through the launcher."""
    import tempfile
    from oracle import make_golden as mg, refload
    m = refload.load("SAC_file", "SAC_add_discrete")
    m.is_continue = True          # the actor branch reads the script's module-level global (SAC_add_discrete.py:328)
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    old_out, mg.OUT = mg.OUT, tempfile.mkdtemp(dir=str(tmp_path))
    try:
        mg.gen_offpolicy("sacd", lambda: m.SAC([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=trick),
                         lambda p, B: p.learn(B, 0.99, 0.01), 3, 64, 17, 6, 2,
                         {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                          "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target},
                         extra=lambda p: {"final/log_alpha": np.array(p.alphas.log_alpha.item(), np.float64)})
        got = dict(np.load(os.path.join(mg.OUT, "sacd.npz")))
    finally:
        mg.OUT = old_out
    want = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "sac.npz")))
    want = {k: v for k, v in want.items() if not k.startswith("sel/")}       # the stochastic select_action record is written for the name "sac" only
    assert sorted(got) == sorted(want) and all(np.array_equal(got[k], want[k]) for k in want)
    from freerl_b200 import launcher
    ns = launcher.run_reference_script(os.path.join(REF, "SAC_file", "SAC_add_discrete.py"),
                                       ["--env_name", "Pendulum-v1", "--max_episodes", "1", "--start_steps", "60", "--random_steps", "20",
                                        "--batch_size", "32", "--buffer_size", "2000", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__module__ == "freerl_b200.SAC_add_discrete" and ns["policy"].agent.critic_step > 0
    # the discrete `hands_on` branch: the same unchanged script on a discrete env trains through frl_sacd_learn
    ns = launcher.run_reference_script(os.path.join(REF, "SAC_file", "SAC_add_discrete.py"),
                                       ["--env_name", "CartPole-v1", "--max_episodes", "2", "--start_steps", "40", "--random_steps", "20",
                                        "--batch_size", "32", "--buffer_size", "2000", "--device", "cpu"], results_root=str(tmp_path))
    assert type(ns["policy"]).__name__ == "_DiscreteSAC" and ns["policy"].agent.critic_step > 0
