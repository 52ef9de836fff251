"""Resume checkpoints (freerl_b200/checkpoint.py): train, checkpoint, keep training; a second policy built with the same arguments,
restored from the file and trained with the same calls must end with BIT-IDENTICAL parameters, optimiser state, replay and RNG
— for the off-policy, distributional / PER / n-step, on-policy and multi-agent families."""
import contextlib
import io

import numpy as np
import pytest
import torch

from freerl_b200.checkpoint import load_checkpoint, save_checkpoint, state_of


def _quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def _same(a, b):
    sa, sb = state_of(a), state_of(b)
    assert sa.keys() == sb.keys()
    for k in sa:
        if isinstance(sa[k], torch.Tensor):
            assert torch.equal(sa[k].cpu(), sb[k].cpu()), k
        elif isinstance(sa[k], np.ndarray):
            assert np.array_equal(sa[k], sb[k]), k
        elif not hasattr(sa[k], "maxlen"):
            assert sa[k] == sb[k] or (sa[k] != sa[k] and sb[k] != sb[k]), k


def _case(make, feed, learn, tmp_path, rounds=2):
    torch.manual_seed(1); np.random.seed(1)
    a = _quiet(make)
    rng = np.random.default_rng(0)
    feed(a, rng)
    for _ in range(rounds):
        learn(a)
    path = str(tmp_path / "ck.pt")
    save_checkpoint(a, path)
    rng_state = rng.bit_generator.state
    feed(a, rng); learn(a); learn(a)
    torch.manual_seed(99); np.random.seed(99)                   # different init and RNG: everything must come from the file
    b = _quiet(make)
    load_checkpoint(b, path)
    rng2 = np.random.default_rng(0); rng2.bit_generator.state = rng_state
    feed(b, rng2); learn(b); learn(b)
    _same(a, b)


def _run(device, tmp_path):
    from freerl_b200.DQN_with_tricks import DQN as Rainbow
    from freerl_b200.MADDPG import MADDPG
    from freerl_b200.PPO import PPO
    from freerl_b200.SAC import SAC
    from freerl_b200.TD3 import TD3

    def feed_ac(p, rng, n=80, od=5, ad=2):
        p.add(rng.standard_normal((n, od)), rng.uniform(-1, 1, (n, ad)), rng.standard_normal(n), rng.standard_normal((n, od)), rng.random(n) < 0.1)
    for mode in ("parity", "fast"):
        _case(lambda: SAC([5, 2], True, 1e-3, 1e-3, 256, device, trick={"Batch_ObsNorm": True}, mode=mode), feed_ac,
              lambda p: p.learn(32, 0.99, 0.01, n_updates=2), tmp_path)
    _case(lambda: TD3([5, 2], True, 1e-3, 1e-3, 256, device, trick=None, realize={"clip_double": True, "policy_noise": True, "twin_delay": True}),
          feed_ac, lambda p: p.learn(32, 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0), tmp_path, rounds=3)
    trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}

    def feed_rb(p, rng):
        for _ in range(5):
            p.add(rng.standard_normal((16, 4)), rng.integers(0, 3, (16, 1)), rng.standard_normal(16), rng.standard_normal((16, 4)), rng.random(16) < 0.1)
    _case(lambda: Rainbow([4, 3], False, 1e-3, 300, device, trick=trick, gamma=0.99, batch_size=16), feed_rb,
          lambda p: p.learn(16, 0.99, 0.01), tmp_path)

    def feed_ppo(p, rng):
        for _ in range(64 - len(p.buffer)):
            o = rng.standard_normal(6).astype(np.float32)
            act, lp = p.select_action(o)
            p.add(o, act, float(rng.standard_normal()), rng.standard_normal(6).astype(np.float32), bool(rng.random() < 0.05), lp, bool(rng.random() < 0.1))
    _case(lambda: PPO([6, 2], True, 1e-3, 1e-3, 64, device), feed_ppo, lambda p: (p.learn(32, 0.99, 0.95, 0.2, 2, 0.01), feed_ppo(p, np.random.default_rng(p.agent.step))),
          tmp_path, rounds=1)
    ids = ["a", "b"]
    sup = {"weight_decay": True, "OUNoise": False, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": True}

    def feed_ma(p, rng, n=60):
        p.add({k: rng.standard_normal((n, 4)) for k in ids}, {k: rng.uniform(-1, 1, (n, 2)) for k in ids}, {k: rng.standard_normal(n) for k in ids},
              {k: rng.standard_normal((n, 4)) for k in ids}, {k: rng.random(n) < 0.1 for k in ids})
    _case(lambda: MADDPG({k: [4, 2] for k in ids}, True, 1e-3, 1e-3, 256, device, None, sup), feed_ma, lambda p: p.learn(16, 0.95, 0.01), tmp_path)

    # MAPPO_discrete (shared nets, episode ReplayBuffer, the script's default switches): the checkpoint is taken in the MIDDLE of a
    # rollout — three of four episodes stored — so the host-side episode arrays and the episode counter are part of the state
    from freerl_b200.MAPPO_discrete import MAPPO as MAPPOd, ReplayBuffer
    from oracle.make_golden_mappo_discrete import TRICKS          # the switch sets only
    Nn, OD, AD, T, B = 2, 6, 3, 4, 4

    def episode(p, rng):
        for t in range(T):
            obs = [rng.standard_normal(OD).astype(np.float32) for _ in range(Nn)]
            a, lp = p.select_action(obs)
            p.add(obs, a, {k: float(rng.standard_normal()) for k in ids}, None, {k: t == T - 1 for k in ids}, lp, None, t)
        p.buffer.store_last_value(T, p.get_value(rng.standard_normal(Nn * OD).astype(np.float32)))

    def feed_md(p, rng):
        while p.buffer.episode_num < B - 1:
            episode(p, rng)

    def learn_md(p):
        rng = np.random.default_rng(p.agent.step)
        episode(p, rng)
        p.learn(2, 0.95, 0.95, 0.2, 2, 0.01, 10.0)
        feed_md(p, rng)
    _case(lambda: MAPPOd({k: [OD, AD] for k in ids}, False, 1e-3, 5e-4, B, device, dict(TRICKS["full"]),
                         ReplayBuffer(N=Nn, obs_dim=OD, state_dim=Nn * OD, episode_limit=T, batch_size=B, device=device)),
          feed_md, learn_md, tmp_path, rounds=1)


def test_checkpoint_resume_emulated(emul, tmp_path):
    _run(torch.device("cpu"), tmp_path)


@pytest.mark.gpu
def test_checkpoint_resume_gpu(tmp_path):
    _run(torch.device("cuda"), tmp_path)


def test_checkpoint_rejects_other_class(emul, tmp_path):
    from freerl_b200.DDPG import DDPG
    from freerl_b200.SAC import SAC
    a = _quiet(lambda: SAC([5, 2], True, 1e-3, 1e-3, 64, torch.device("cpu"), trick={}))
    save_checkpoint(a, str(tmp_path / "x.pt"))
    b = _quiet(lambda: DDPG([5, 2], True, 1e-3, 1e-3, 64, torch.device("cpu")))
    with pytest.raises(ValueError):
        load_checkpoint(b, str(tmp_path / "x.pt"))


def _run_more(device, tmp_path):
    """the classes added after the first slice: MAPPO (both heads), IPPO, HAPPO, discrete SAC with its per-learn staging replay,
    PPO_with_tricks (Batch_ObsNorm, Beta head), MATD3"""
    import importlib
    ids = ["a", "b"]
    mt = {"adv_norm": True, "ObsNorm": False, "reward_norm": False, "reward_scaling": False, "orthogonal_init": True, "adam_eps": True,
          "lr_decay": False, "ValueClip": True, "huber_loss": True, "LayerNorm": True, "feature_norm": True}

    def feed_on(p, rng):
        while len(p.buffers["a"]) < 32:
            obs = {k: rng.standard_normal(5).astype(np.float32) for k in ids}
            a, lp = p.select_action(obs)
            p.add(obs, a, {k: float(rng.standard_normal()) for k in ids}, {k: rng.standard_normal(5).astype(np.float32) for k in ids},
                  {k: bool(rng.random() < 0.05) for k in ids}, lp, {k: bool(rng.random() < 0.1) for k in ids})

    def learn_on(p):
        p.learn(16, 0.95, 0.95, 0.2, 2, 0.01, 10.0)
        feed_on(p, np.random.default_rng(sum(a.step for a in p.agents.values())))
    for mod, cls, cont in (("MAPPO", "MAPPO", True), ("MAPPO", "MAPPO", False), ("IPPO", "IPPO", True), ("HAPPO", "HAPPO", True)):
        C = getattr(importlib.import_module("freerl_b200." + mod), cls)
        _case(lambda: C({k: [5, 3] for k in ids}, cont, 1e-3, 5e-4, 32, device, dict(mt)), feed_on, learn_on, tmp_path, rounds=1)

    from freerl_b200.SAC_add_discrete import SAC as SACd

    def feed_sd(p, rng, n=80):
        p.add(rng.standard_normal((n, 4)), rng.integers(0, 3, (n, 1)), rng.standard_normal(n), rng.standard_normal((n, 4)), rng.random(n) < 0.1)
    _case(lambda: SACd([4, 3], False, 1e-3, 1e-3, 256, device, trick={"Batch_ObsNorm": True}), feed_sd, lambda p: p.learn(32, 0.99, 0.01), tmp_path)

    from freerl_b200.PPO_with_tricks import PPO as PPOt
    pt = {"adv_norm": True, "ObsNorm": False, "reward_norm": False, "reward_scaling": False, "orthogonal_init": True, "adam_eps": True,
          "lr_decay": False, "tanh": False, "Batch_ObsNorm": True}

    def feed_pt(p, rng):
        for _ in range(64 - len(p.buffer)):
            o = rng.standard_normal(6).astype(np.float32)
            act, lp = p.select_action(o)
            p.add(o, act, float(rng.standard_normal()), rng.standard_normal(6).astype(np.float32), bool(rng.random() < 0.05), lp, bool(rng.random() < 0.1))
    for kw in ({}, {"beta": True}):
        _case(lambda: PPOt([6, 2], True, 1e-3, 1e-3, 64, device, trick=dict(pt), **kw), feed_pt,
              lambda p: (p.learn(32, 0.99, 0.95, 0.2, 2, 0.01), feed_pt(p, np.random.default_rng(p.agent.step))), tmp_path, rounds=1)

    from freerl_b200.MATD3_simple import MATD3

    def feed_ma(p, rng, n=60):
        p.add({k: rng.standard_normal((n, 4)) for k in ids}, {k: rng.uniform(-1, 1, (n, 2)) for k in ids}, {k: rng.standard_normal(n) for k in ids},
              {k: rng.standard_normal((n, 4)) for k in ids}, {k: rng.random(n) < 0.1 for k in ids})
    _case(lambda: MATD3({k: [4, 2] for k in ids}, True, 1e-3, 1e-3, 256, device, trick=None), feed_ma,
          lambda p: p.learn(16, 0.95, 0.01, 0.1, 0.5, 1.0, 2, 1.0), tmp_path, rounds=3)


def test_checkpoint_resume_more_classes_emulated(emul, tmp_path):
    _run_more(torch.device("cpu"), tmp_path)


@pytest.mark.gpu
def test_checkpoint_resume_more_classes_gpu(tmp_path):
    _run_more(torch.device("cuda"), tmp_path)
