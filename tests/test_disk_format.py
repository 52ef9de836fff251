"""On-disk formats: a model saved by freerl_b200 loads with the UNMODIFIED reference class's own ``load()`` and acts
the same, and the other way round — which is what keeps the reference's ``evaluate.py`` / ``MA_evaluate.py`` tools usable on our runs.
Needs the reference tree (build container only); the product side runs on the host emulation here."""
import os

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "DQN_file")), reason="reference tree not present")

SAC_TRICK = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
TD3_TRICK = {"Batch_ObsNorm": False}
DDPG_SUP = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": False}
PPO_TRICKS = {"adv_norm": True, "ObsNorm": False, "reward_norm": False, "reward_scaling": False, "orthogonal_init": True, "adam_eps": True,
              "lr_decay": False, "tanh": False, "Batch_ObsNorm": False}
CASES = [
    # name, reference dir, module, class, our module, dim_info, is_continue, ctor args after (dim_info, is_continue), kwargs, file
    ("dqn", "DQN_file", "DQN", "DQN", "freerl_b200.DQN", [4, 3], False, (1e-3, 100), {}, "DQN.pt"),
    ("sac", "SAC_file", "SAC", "SAC", "freerl_b200.SAC", [5, 2], True, (1e-3, 1e-3, 100), {"trick": SAC_TRICK}, "SAC.pt"),
    ("ppo_cont", "PPO_file", "PPO", "PPO", "freerl_b200.PPO", [6, 2], True, (1e-3, 1e-3, 32), {"trick": {"adv_norm": False}}, "PPO.pt"),
    ("ppo_disc", "PPO_file", "PPO", "PPO", "freerl_b200.PPO", [6, 3], False, (1e-3, 1e-3, 32), {"trick": {"adv_norm": False}}, "PPO.pt"),
    ("ppo_advance", "PPO_advance", "PPO", "PPO", "freerl_b200.PPO_advance", [6, 2], True, (1e-3, 1e-3, 32), {"trick": {"adv_norm": False}}, "PPO.pt"),
    ("td3", "TD3_file", "TD3", "TD3", "freerl_b200.TD3", [5, 2], True, (1e-3, 1e-3, 100),
     {"trick": None, "realize": {"clip_double": True, "policy_noise": True, "twin_delay": True}}, "TD3.pt"),
    ("ddpg", "DDPG_file", "DDPG", "DDPG", "freerl_b200.DDPG", [5, 2], True, (1e-3, 1e-3, 100), {"trick": None, "supplement": DDPG_SUP}, "DDPG.pt"),
    ("ddpg_simple", "DDPG_file", "DDPG_simple", "DDPG", "freerl_b200.DDPG_simple", [5, 2], True, (1e-3, 1e-3, 100), {"trick": None}, "DDPG.pt"),
    ("ppo_tricks", "PPO_file", "PPO_with_tricks", "PPO", "freerl_b200.PPO_with_tricks", [6, 2], True, (1e-3, 1e-3, 32),
     {"trick": PPO_TRICKS, "beta": False}, "PPO.pt"),
    ("ppo_tricks_beta", "PPO_file", "PPO_with_tricks", "PPO", "freerl_b200.PPO_with_tricks", [6, 2], True, (1e-3, 1e-3, 32),
     {"trick": PPO_TRICKS, "beta": True}, "PPO.pt"),          # alpha_layer / beta_layer keys: two row slices of one device layer
    ("ppo_tricks_disc", "PPO_file", "PPO_with_tricks", "PPO", "freerl_b200.PPO_with_tricks", [6, 3], False, (1e-3, 1e-3, 32),
     {"trick": PPO_TRICKS, "beta": False}, "PPO.pt"),
]


def _acts(policy, obs):
    return np.stack([np.asarray(policy.evaluate_action(o), dtype=np.float64).reshape(-1) for o in obs])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_saved_models_interoperate_with_the_reference(case, tmp_path, emul):
    import importlib
    from oracle import refload
    name, rdir, rmod, cls, ours_mod, dim_info, is_continue, args, kw, fname = case
    ref = getattr(refload.load(rdir, rmod), cls)
    ours = getattr(importlib.import_module(ours_mod), cls)
    dev = torch.device("cpu")
    obs = np.random.default_rng(1).standard_normal((6, dim_info[0])).astype(np.float32)
    # ours -> disk -> reference.load()
    torch.manual_seed(5)
    a = ours(dim_info, is_continue, *args, dev, **kw)
    d1 = tmp_path / "ours"; d1.mkdir()
    a.save(str(d1))
    assert os.path.exists(d1 / fname)
    r = ref.load(dim_info, is_continue, str(d1), **kw)
    np.testing.assert_allclose(_acts(r, obs), _acts(a, obs), rtol=1e-5, atol=2e-6)
    # reference -> disk -> ours.load()
    torch.manual_seed(6)
    r2 = ref(dim_info, is_continue, *args, dev, **kw)
    d2 = tmp_path / "ref"; d2.mkdir()
    r2.save(str(d2))
    b = ours.load(dim_info, is_continue, str(d2), device=dev, **kw)
    np.testing.assert_allclose(_acts(b, obs), _acts(r2, obs), rtol=1e-5, atol=2e-6)
    # same keys, shapes and dtypes on disk
    s1, s2 = torch.load(d1 / fname), torch.load(d2 / fname)
    assert list(s1) == list(s2) and all(s1[k].shape == s2[k].shape and s1[k].dtype == s2[k].dtype for k in s1)


def test_maddpg_checkpoint_interoperates_with_the_reference(tmp_path, emul):
    """MADDPG.pth = {agent_id: actor state_dict} (MADDPG_file/MADDPG.py:240-253), both directions."""
    from freerl_b200.MADDPG import MADDPG
    from oracle import refload
    ref = refload.load("MADDPG_file", "MADDPG").MADDPG
    sup = {"weight_decay": False, "OUNoise": False, "ObsNorm": False, "net_init": False, "Batch_ObsNorm": False}
    dim_info = {"agent_%d" % i: [7, 3] for i in range(3)}
    dev = torch.device("cpu")
    rng = np.random.default_rng(2)
    obs = [{k: rng.standard_normal(7).astype(np.float32) for k in dim_info} for _ in range(4)]
    acts = lambda pol: np.stack([np.concatenate([np.asarray(v, np.float64).reshape(-1) for v in pol.evaluate_action(o).values()]) for o in obs])
    torch.manual_seed(3)
    a = MADDPG(dim_info, True, 1e-3, 1e-3, 100, dev, trick=None, supplement=dict(sup))
    d1 = tmp_path / "ours"; d1.mkdir()
    a.save(str(d1))
    r = ref.load(dim_info, True, str(d1), trick=None, supplement=dict(sup))
    np.testing.assert_allclose(acts(r), acts(a), rtol=1e-5, atol=2e-6)
    torch.manual_seed(4)
    r2 = ref(dim_info, True, 1e-3, 1e-3, 100, dev, trick=None, supplement=dict(sup))
    d2 = tmp_path / "ref"; d2.mkdir()
    r2.save(str(d2))
    b = MADDPG.load(dim_info, True, str(d2), trick=None, supplement=dict(sup), device=dev)
    np.testing.assert_allclose(acts(b), acts(r2), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("algo", ["MAPPO", "IPPO", "HAPPO"])
def test_mappo_family_checkpoints_interoperate_with_the_reference(algo, tmp_path, emul):
    """MAPPO.pth / IPPO.pth / HAPPO.pth = {agent_id: actor state_dict} (e.g. MAPPO_file/MAPPO.py:494-507), both directions."""
    import importlib
    from oracle import refload
    from oracle.make_golden_marl import MAPPO_TRICK
    refmod = refload.load("MAPPO_file", algo)
    ref = getattr(refmod, algo)

    def fresh():      # IPPO / HAPPO number their critics with a class-level counter (IPPO.py:131-134): one policy per process upstream
        if hasattr(refmod.Critic, "id_num"):
            refmod.Critic.id_num = 0
    ours = getattr(importlib.import_module("freerl_b200." + algo), algo)
    dim_info = {"agent_%d" % i: [9, 3] for i in range(3)}
    dev = torch.device("cpu")
    rng = np.random.default_rng(4)
    obs = [{k: rng.standard_normal(9).astype(np.float32) for k in dim_info} for _ in range(4)]
    acts = lambda pol: np.stack([np.concatenate([np.asarray(v, np.float64).reshape(-1) for v in pol.evaluate_action(o).values()]) for o in obs])
    torch.manual_seed(7)
    a = ours(dim_info, True, 1e-3, 1e-3, 32, dev, dict(MAPPO_TRICK))
    d1 = tmp_path / "ours"; d1.mkdir()
    a.save(str(d1))
    fresh()
    r = ref.load(dim_info, True, str(d1), trick=dict(MAPPO_TRICK))
    np.testing.assert_allclose(acts(r), acts(a), rtol=1e-5, atol=2e-6)
    torch.manual_seed(8)
    fresh()
    r2 = ref(dim_info, True, 1e-3, 1e-3, 32, dev, dict(MAPPO_TRICK))
    d2 = tmp_path / "ref"; d2.mkdir()
    r2.save(str(d2))
    b = ours.load(dim_info, True, str(d2), trick=dict(MAPPO_TRICK), device=dev)
    np.testing.assert_allclose(acts(b), acts(r2), rtol=1e-5, atol=2e-6)


def test_rainbow_checkpoint_schema_matches_the_reference(tmp_path, emul):
    """Rainbow DQN.pt (DQN_file/DQN_with_tricks.py:297-298): NoisyLinear parameters AND the weight_epsilon / bias_epsilon buffers.
    Upstream ``DQN.load`` cannot run (it omits gamma / batch_size), so the file written by us is loaded with a strict
    ``load_state_dict`` into a reference network built by the reference constructor, and the reverse through ours."""
    from freerl_b200.DQN_with_tricks import DQN
    from oracle import refload
    ref = refload.load("DQN_file", "DQN_with_tricks").DQN
    trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
    dev = torch.device("cpu")
    torch.manual_seed(9)
    a = DQN([8, 4], False, 1e-3, 64, dev, trick=dict(trick), gamma=0.99, batch_size=32)
    d1 = tmp_path / "ours"; d1.mkdir()
    a.save(str(d1))
    r = ref([8, 4], False, 1e-3, 64, dev, trick=dict(trick), gamma=0.99, batch_size=32)
    sd = torch.load(d1 / "DQN.pt")
    r.agent.Qnet.load_state_dict(sd, strict=True)
    assert list(sd) == list(r.agent.Qnet.state_dict())
    obs = np.random.default_rng(3).standard_normal((5, 8)).astype(np.float32)
    assert [int(a.evaluate_action(o)) for o in obs] == [int(r.evaluate_action(o)) for o in obs]
    d2 = tmp_path / "ref"; d2.mkdir()
    r.save(str(d2))
    b = DQN.load([8, 4], False, str(d2), trick=dict(trick), device=dev, gamma=0.99, batch_size=32)
    assert [int(b.evaluate_action(o)) for o in obs] == [int(r.evaluate_action(o)) for o in obs]


@pytest.mark.parametrize("tricks", ["simple", "full"])
def test_mappo_discrete_checkpoint_schema_matches_the_reference(tmp_path, emul, tricks):
    """MAPPO_discrete.pth (MAPPO_file/MAPPO_discrete.py:389-393) = the shared actor's state dict.  Upstream ``MAPPO.load`` cannot run (it
    builds ``MAPPO(...)`` without a buffer and dereferences ``buffer.batch_size``), so the file written by us is loaded with a strict
    ``load_state_dict`` into a reference policy built by the reference constructor, and the reverse through ours; greedy actions agree
    (``full``: per-row LayerNorm in the acting network, hidden layers only — the discrete actor drops its normalised input)."""
    import sys
    from freerl_b200.MAPPO_discrete import MAPPO, ReplayBuffer
    from oracle import refload
    from oracle.make_golden_mappo_discrete import TRICKS
    refmod = refload.load("MAPPO_file", "MAPPO_discrete")
    ref_buffer = sys.modules["Buffer"].ReplayBuffer
    dim_info = {"agent_%d" % i: [9, 4] for i in range(3)}
    trick = dict(TRICKS[tricks])
    dev = torch.device("cpu")
    obs = np.random.default_rng(2).standard_normal((5, 3, 9)).astype(np.float32)
    torch.manual_seed(11)
    a = MAPPO(dim_info, False, 1e-3, 5e-4, 4, dev, dict(trick), ReplayBuffer(3, 9, 27, 5, 4, dev))
    d1 = tmp_path / "ours"; d1.mkdir()
    a.save(str(d1))
    r = refmod.MAPPO(dim_info, False, 1e-3, 5e-4, 4, dev, dict(trick), ref_buffer(N=3, obs_dim=9, state_dim=27, episode_limit=5, batch_size=4, device="cpu"))
    sd = torch.load(d1 / "MAPPO_discrete.pth")
    r.agent.actor.load_state_dict(sd, strict=True)
    assert list(sd) == list(r.agent.actor.state_dict())
    assert [a.evaluate_action(o).tolist() for o in obs] == [r.evaluate_action(o).tolist() for o in obs]
    d2 = tmp_path / "ref"; d2.mkdir()
    r.save(str(d2))
    b = MAPPO.load(dim_info, False, str(d2), trick=dict(trick), device=dev)
    assert [b.evaluate_action(o).tolist() for o in obs] == [r.evaluate_action(o).tolist() for o in obs]
