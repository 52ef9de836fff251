"""Golden fixtures for the plain-Adam PPO of the UNMODIFIED reference ``PPO_advance/PPO.py`` (separate actor / critic
``torch.optim.Adam``, ``Categorical(probs=softmax)`` head; see oracle/make_golden.py for the conventions).

    python -m oracle.make_golden_ppo_advance     # writes tests/golden/ppo_adv_{cont,disc}.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refload  # noqa: E402
from oracle.make_golden import OUT, LossTap, rng_restore, rng_snapshot, sd_np  # noqa: E402


def gen(is_continue):
    m = refload.load("PPO_advance", "PPO")
    seed, horizon, mb, K = 9, 256, 64, 2
    obs_dim, act_dim = 8, (2 if is_continue else 4)
    np.random.seed(seed)
    torch.manual_seed(seed)
    policy = m.PPO([obs_dim, act_dim], is_continue, 1e-3, 5e-4, horizon, torch.device("cpu"), trick={'adv_norm': False})
    rng = np.random.default_rng(seed)
    rec = {}
    rec.update(sd_np(policy.agent.actor, "init/actor/"))
    rec.update(sd_np(policy.agent.critic, "init/critic/"))
    obs = rng.standard_normal(obs_dim).astype(np.float32)
    for t in range(horizon):
        a, logp = policy.select_action(obs)
        o2 = rng.standard_normal(obs_dim).astype(np.float32)
        term = bool(rng.random() < 0.02)
        trunc = (t % 50) == 49
        policy.add(obs, a, float(rng.standard_normal()), o2, term, logp, term or trunc)
        obs = o2
    for k, t in zip(("obs", "act", "rew", "nobs", "done", "logp", "adv_done"), policy.buffer.all()):
        rec["data/" + k] = t.numpy().copy()
    tap = LossTap(policy.agent, ["update_actor", "update_critic"])
    before = rng_snapshot()
    policy.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
    after = rng_snapshot()
    rng_restore(before)
    for k in range(K):
        rec["perm/%d" % k] = np.random.permutation(horizon)
    rng_restore(after)
    rec.update(sd_np(policy.agent.actor, "final/actor/"))
    rec.update(sd_np(policy.agent.critic, "final/critic/"))
    log = [v[0] for _, v in tap.log]
    rec["losses"] = np.array(list(zip(log[0::2], log[1::2])), np.float64)          # (actor, critic) per minibatch
    rec["act/obs"] = rng.standard_normal((8, obs_dim)).astype(np.float32)
    rec["act/eval"] = np.array([policy.evaluate_action(o) for o in rec["act/obs"]])
    name = "ppo_adv_cont" if is_continue else "ppo_adv_disc"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok", rec["losses"][:2])


if __name__ == "__main__":
    torch.set_num_threads(1)
    gen(True)
    gen(False)
