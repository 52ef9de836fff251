"""Generate golden fixtures by running the UNMODIFIED reference (``/root/reference``) in this container.

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together
with this script.  Each fixture records everything a replay needs: initial parameters, the sampled
batches (drawn from the reference's own RNG call sites), every random tensor in reference call order,
the losses the reference computed and the parameters / targets after k learns.

RNG capture: before each ``learn()`` the numpy-legacy and torch CPU generator states are saved; after it
they are restored and the draws are re-made HERE in the order documented in SURVEY.md §8c
(``np.random.choice`` then ``torch.randn`` per rsample / randn_like), then the post-learn states are put
back.  ``tests/test_oracle_golden.py`` proves that order right by replaying through ``oracle.algos``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sd_np(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def fill(policy, n, obs_dim, act_dim, rng, discrete_actions=None, offset=None):
    for _ in range(n):
        o = rng.standard_normal(obs_dim).astype(np.float32)
        if offset is not None:          # Batch_ObsNorm fixtures: observations with a non-zero mean, like real envs
            o = o + offset
        a = rng.integers(0, discrete_actions) if discrete_actions else rng.uniform(-1, 1, act_dim).astype(np.float32)
        r = float(rng.standard_normal())
        o2 = rng.standard_normal(obs_dim).astype(np.float32)
        if offset is not None:
            o2 = o2 + offset
        d = bool(rng.random() < 0.1)
        policy.add(o, a, r, o2, d)


class LossTap:
    """Record the scalar passed to Agent.update_* without changing behaviour."""

    def __init__(self, agent, names):
        self.log = []
        for n in names:
            orig = getattr(agent, n)

            def wrapped(*losses, _orig=orig, _n=n):
                self.log.append((_n, [float(l.item()) for l in losses]))
                return _orig(*losses)
            setattr(agent, n, wrapped)


def rng_snapshot():
    return np.random.get_state(), torch.get_rng_state()


def rng_restore(s):
    np.random.set_state(s[0])
    torch.set_rng_state(s[1])


def gen_offpolicy(name, make_policy, learn_call, n_learn, B, obs_dim, act_dim, noise_draws, nets, discrete=None,
                  extra=None, seed=3, bon=False, n_fill=400):
    np.random.seed(seed)
    torch.manual_seed(seed)
    policy = make_policy()
    rng = np.random.default_rng(seed)
    offset = rng.uniform(1.0, 3.0, obs_dim).astype(np.float32) if bon else None
    fill(policy, n_fill, obs_dim, act_dim, rng, discrete, offset)
    tap = LossTap(policy.agent, [n for n in ("update_critic", "update_actor", "update_Qnet") if hasattr(policy.agent, n)])
    rec = {}
    for nm, getter in nets.items():
        rec.update(sd_np(getter(policy), "init/%s/" % nm))
    for it in range(n_learn):
        before = rng_snapshot()
        learn_call(policy, B)
        after = rng_snapshot()
        rng_restore(before)
        idx = np.random.choice(len(policy.buffer), B, replace=False)
        batch = policy.buffer.sample(idx)
        rec["idx/%d" % it] = idx
        for k, t in zip(("obs", "act", "rew", "nobs", "done"), batch):
            rec["batch/%d/%s" % (it, k)] = t.numpy().copy()
        for j in range(noise_draws):
            rec["noise/%d/%d" % (it, j)] = torch.randn(B, act_dim).numpy().copy()
        rng_restore(after)
    for nm, getter in nets.items():
        rec.update(sd_np(getter(policy), "final/%s/" % nm))
    for i, (n, vals) in enumerate(tap.log):
        rec["loss/%03d/%s" % (i, n)] = np.array(vals, np.float64)
    if extra:
        rec.update(extra(policy))
    if bon:     # Batch_ObsNorm running statistics after the learns + one normalised select_action
        ms = policy.batch_size_obs_norm.running_ms
        rec["final/norm/mean"], rec["final/norm/std"] = ms.mean.numpy().copy(), ms.std.numpy().copy()
        rec["final/norm/n"] = np.array(ms.n)
        o = rng.standard_normal(obs_dim).astype(np.float32) + offset
        rec["act/obs"] = o
        if name.startswith("ddpg"):
            rec["act/action"] = np.asarray(policy.select_action(o))
    if name == "sac":     # stochastic select_action (SAC.py:192-198: tanh(rsample)) after the learns, with the generator state it starts from
        o = rng.standard_normal((8, obs_dim)).astype(np.float32)
        rec["sel/obs"] = o
        rec["sel/rng_state"] = torch.get_rng_state().numpy().copy()
        rec["sel/action"] = np.stack([np.asarray(policy.select_action(o[i])) for i in range(8)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok:", len(rec), "arrays;", [(n, v) for n, v in tap.log[:3]])


def gen_dqn():
    m = refload.load("DQN_file", "DQN")
    gen_offpolicy("dqn", lambda: m.DQN([4, 2], False, 1e-3, 1000, torch.device("cpu")),
                  lambda p, B: p.learn(B, 0.99, 0.01), 3, 64, 4, 1, 0,
                  {"q": lambda p: p.agent.Qnet, "q_target": lambda p: p.agent.Qnet_target}, discrete=2)


def gen_sac():
    m = refload.load("SAC_file", "SAC")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    gen_offpolicy("sac", lambda: m.SAC([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=trick),
                  lambda p, B: p.learn(B, 0.99, 0.01), 3, 64, 17, 6, 2,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                   "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target},
                  extra=lambda p: {"final/log_alpha": np.array(p.alphas.log_alpha.item(), np.float64)})


def gen_sac_b256():
    """The bench's batch shape: SAC, B = 256, ten chained learns (the other fixtures are B = 64, k <= 4)."""
    m = refload.load("SAC_file", "SAC")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    gen_offpolicy("sac_b256", lambda: m.SAC([17, 6], True, 1e-3, 1e-3, 2000, torch.device("cpu"), trick=trick),
                  lambda p, B: p.learn(B, 0.99, 0.01), 10, 256, 17, 6, 2,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic, "actor_target": lambda p: p.agent.actor_target,
                   "critic_target": lambda p: p.agent.critic_target},
                  extra=lambda p: {"final/log_alpha": p.alphas.log_alpha.detach().numpy().copy()}, seed=13, n_fill=1200)


def gen_sac_discrete():
    """``SAC_add_discrete.py`` with ``is_continue=False`` (the ``hands_on`` branch): obs 8, 4 actions, B = 64, three learns, plus
    stochastic ``select_action`` calls (Categorical sampling) with the generator state they start from."""
    m = refload.load("SAC_file", "SAC_add_discrete")
    m.is_continue = False       # learn() reads the module global its __main__ block defines (SAC_add_discrete.py:325 <- :494)
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}

    def extra(p):
        rng = np.random.default_rng(77)
        o = rng.standard_normal((8, 8)).astype(np.float32)
        rec = {"final/log_alpha": np.array(p.alphas.log_alpha.item(), np.float64), "sel/obs": o,
               "sel/rng_state": torch.get_rng_state().numpy().copy()}
        rec["sel/action"] = np.stack([np.asarray(p.select_action(o[i])) for i in range(8)])
        rec["sel/greedy"] = np.stack([np.asarray(p.evaluate_action(torch.as_tensor(o[i]).reshape(1, -1))) for i in range(8)])
        return rec
    gen_offpolicy("sac_discrete", lambda: m.SAC([8, 4], False, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=trick),
                  lambda p, B: p.learn(B, 0.99, 0.01), 3, 64, 8, 1, 0,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic, "actor_target": lambda p: p.agent.actor_target,
                   "critic_target": lambda p: p.agent.critic_target}, discrete=4, extra=extra, seed=23)


def gen_sac_bon():
    """SAC with trick Batch_ObsNorm (SAC.py:181-182, 215-217)."""
    m = refload.load("SAC_file", "SAC")
    trick = {"ObsNorm": False, "Batch_ObsNorm": True, "OUNoise": True, "GaussNoise": False}
    gen_offpolicy("sac_bon", lambda: m.SAC([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=trick),
                  lambda p, B: p.learn(B, 0.99, 0.01), 4, 64, 17, 6, 2,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                   "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target},
                  extra=lambda p: {"final/log_alpha": np.array(p.alphas.log_alpha.item(), np.float64)}, seed=21, bon=True)


def gen_ddpg_bon():
    """DDPG with its DEFAULT supplements (DDPG.py:446): weight_decay, net_init, Batch_ObsNorm all on."""
    m = refload.load("DDPG_file", "DDPG")
    sup = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": True}
    gen_offpolicy("ddpg_bon", lambda: m.DDPG([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=None, supplement=sup),
                  lambda p, B: p.learn(B, 0.99, 0.01), 4, 64, 17, 6, 0,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                   "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target},
                  seed=22, bon=True)


def gen_td3():
    m = refload.load("TD3_file", "TD3")
    realize = {"clip_double": True, "policy_noise": True, "twin_delay": True}
    gen_offpolicy("td3", lambda: m.TD3([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=None, realize=realize),
                  lambda p, B: p.learn(B, 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0), 4, 64, 17, 6, 1,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                   "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target})


def gen_ddpg():
    m = refload.load("DDPG_file", "DDPG")
    sup = {"weight_decay": True, "OUNoise": True, "ObsNorm": False, "net_init": True, "Batch_ObsNorm": False}
    gen_offpolicy("ddpg", lambda: m.DDPG([17, 6], True, 1e-3, 1e-3, 1000, torch.device("cpu"), trick=None, supplement=sup),
                  lambda p, B: p.learn(B, 0.99, 0.01), 3, 64, 17, 6, 0,
                  {"actor": lambda p: p.agent.actor, "critic": lambda p: p.agent.critic,
                   "actor_target": lambda p: p.agent.actor_target, "critic_target": lambda p: p.agent.critic_target})


def gen_ppo(is_continue):
    m = refload.load("PPO_file", "PPO")
    seed, horizon, mb, K = 5, 256, 64, 2
    obs_dim, act_dim = 8, (2 if is_continue else 4)
    np.random.seed(seed)
    torch.manual_seed(seed)
    policy = m.PPO([obs_dim, act_dim], is_continue, 1e-3, 1e-3, horizon, torch.device("cpu"))
    rng = np.random.default_rng(seed)
    rec = {}
    rec.update(sd_np(policy.agent.actor, "init/actor/"))
    rec.update(sd_np(policy.agent.critic, "init/critic/"))
    obs = rng.standard_normal(obs_dim).astype(np.float32)
    rec["rng/before_rollout"] = torch.get_rng_state().numpy().copy()     # CPU generator state the 256 select_action calls start from
    for t in range(horizon):
        a, logp = policy.select_action(obs)
        o2 = rng.standard_normal(obs_dim).astype(np.float32)
        term = bool(rng.random() < 0.02)
        trunc = (t % 50) == 49
        policy.add(obs, a, float(rng.standard_normal()), o2, term, logp, term or trunc)
        obs = o2
    data = policy.buffer.all()
    for k, t in zip(("obs", "act", "rew", "nobs", "done", "logp", "adv_done"), data):
        rec["data/" + k] = t.numpy().copy()
    tap = LossTap(policy.agent, ["update_ac_"])
    before = rng_snapshot()
    policy.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
    after = rng_snapshot()
    rng_restore(before)
    for k in range(K):
        rec["perm/%d" % k] = np.random.permutation(horizon)
    rng_restore(after)
    rec.update(sd_np(policy.agent.actor, "final/actor/"))
    rec.update(sd_np(policy.agent.critic, "final/critic/"))
    rec["losses"] = np.array([v for _, v in tap.log], np.float64)
    name = "ppo_cont" if is_continue else "ppo_disc"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok", rec["losses"][:2])


def gen_buffers():
    """Ring / sum-tree / PER / n-step KATs from ``DQN_file/Buffer.py`` (incl. a non-power-of-two capacity)."""
    m = refload.load("DQN_file", "Buffer")
    rec = {}
    rng = np.random.default_rng(11)
    for cap in (5, 8, 37, 100):
        np.random.seed(cap)
        per = m.PER_Buffer(cap, 3, 1, torch.device("cpu"))
        n_add = int(cap * 1.6)
        tr = [(rng.standard_normal(3), rng.integers(0, 4), float(rng.standard_normal()), rng.standard_normal(3),
               bool(rng.random() < 0.2)) for _ in range(n_add)]
        rec["per%d/obs" % cap] = np.array([t[0] for t in tr])
        rec["per%d/act" % cap] = np.array([t[1] for t in tr], np.float64)
        rec["per%d/rew" % cap] = np.array([t[2] for t in tr])
        rec["per%d/nobs" % cap] = np.array([t[3] for t in tr])
        rec["per%d/done" % cap] = np.array([t[4] for t in tr])
        half = n_add // 2
        for t in tr[:half]:
            per.add(*t)
        B = min(4, len(per))
        st = np.random.get_state()
        idx, w = per.sample(B)
        rec["per%d/s1_idx" % cap], rec["per%d/s1_w" % cap] = idx, w.numpy()
        np.random.set_state(st)
        seg = per.sumtree.sum() / B
        rec["per%d/s1_u" % cap] = np.array([np.random.uniform(seg * i, seg * (i + 1)) for i in range(B)])
        td = rng.standard_normal((B, 1)).astype(np.float32)
        rec["per%d/td1" % cap] = td
        per.update_priorities(idx, td)
        rec["per%d/tree_mid" % cap] = per.sumtree.tree.copy()
        for t in tr[half:]:
            per.add(*t)
        rec["per%d/tree_end" % cap] = per.sumtree.tree.copy()
        rec["per%d/index_end" % cap] = np.array([per.buffer._index, per.buffer._size])
        rec["per%d/beta_end" % cap] = np.array(per.beta)
        B = min(6, len(per))
        st = np.random.get_state()
        idx, w = per.sample(B)
        rec["per%d/s2_idx" % cap], rec["per%d/s2_w" % cap] = idx, w.numpy()
        np.random.set_state(st)
        seg = per.sumtree.sum() / B
        rec["per%d/s2_u" % cap] = np.array([np.random.uniform(seg * i, seg * (i + 1)) for i in range(B)])
        smp = per.buffer.sample(idx)
        for k, t in zip(("obs", "act", "rew", "nobs", "done"), smp):
            rec["per%d/s2_%s" % (cap, k)] = t.numpy().copy()
    # n-step fold (n=3, gamma .9) with dones sprinkled in
    nb = m.N_Step_PER_Buffer(16, 2, 1, torch.device("cpu"), gamma=0.9)
    tr = [(rng.standard_normal(2), rng.integers(0, 3), float(rng.standard_normal()), rng.standard_normal(2),
           bool(rng.random() < 0.3)) for _ in range(12)]
    for t in tr:
        nb.add(*t)
    rec["nstep/obs_in"] = np.array([t[0] for t in tr])
    rec["nstep/act_in"] = np.array([t[1] for t in tr], np.float64)
    rec["nstep/rew_in"] = np.array([t[2] for t in tr])
    rec["nstep/nobs_in"] = np.array([t[3] for t in tr])
    rec["nstep/done_in"] = np.array([t[4] for t in tr])
    b = nb.buffer
    rec["nstep/obs"], rec["nstep/act"], rec["nstep/rew"] = b.obs.copy(), b.actions.copy(), b.rewards.copy()
    rec["nstep/nobs"], rec["nstep/done"] = b.next_obs.copy(), b.dones.copy()
    rec["nstep/size"] = np.array([b._index, b._size])
    rec["nstep/tree"] = nb.sumtree.tree.copy()
    # uniform index stream KAT: np.random.choice on the legacy global generator
    np.random.seed(0)
    rec["choice/seed0_1000_8"] = np.random.choice(1000, 8, replace=False)
    rec["choice/seed0_next_50_50"] = np.random.choice(50, 50, replace=False)
    np.savez_compressed(os.path.join(OUT, "buffers.npz"), **rec)
    print("buffers ok", len(rec))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    which = sys.argv[1:] or ["buffers", "dqn", "sac", "td3", "ddpg", "ppo"]
    if "buffers" in which:
        gen_buffers()
    if "dqn" in which:
        gen_dqn()
    if "sac" in which:
        gen_sac()
    if "sac_discrete" in which:
        gen_sac_discrete()
    if "sac_b256" in which:
        gen_sac_b256()
    if "td3" in which:
        gen_td3()
    if "ddpg" in which:
        gen_ddpg()
    if "bon" in which:
        gen_sac_bon()
        gen_ddpg_bon()
    if "ppo" in which:
        gen_ppo(True)
        gen_ppo(False)
