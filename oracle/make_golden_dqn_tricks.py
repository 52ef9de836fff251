"""Golden fixtures for the non-distributional trick combinations of the UNMODIFIED reference
``DQN_file/DQN_with_tricks.py`` (see oracle/make_golden.py for the conventions).

    python -m oracle.make_golden_dqn_tricks     # writes tests/golden/dqn_tricks_{double,dueling,d3per}.npz

RNG per learn(): uniform replay -> one ``np.random.choice(total, B, replace=False)``; PER -> B draws of ``np.random.uniform``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refload  # noqa: E402
from oracle.make_golden import OUT, LossTap, fill, rng_restore, rng_snapshot, sd_np  # noqa: E402

OFF = {"Double": False, "Dueling": False, "PER": False, "Noisy": False, "N_Step": False, "Categorical": False}
CASES = {"double": dict(Double=True), "dueling": dict(Dueling=True), "d3per": dict(Double=True, Dueling=True, PER=True, N_Step=True)}


def one(name, on):
    m = refload.load("DQN_file", "DQN_with_tricks")
    trick = dict(OFF, **on)
    seed, B, obs_dim, nA = 11, 32, 8, 4
    np.random.seed(seed)
    torch.manual_seed(seed)
    pol = m.DQN([obs_dim, nA], False, 1e-3, 500, torch.device("cpu"), trick=trick, gamma=0.99, batch_size=B)
    rng = np.random.default_rng(seed)
    fill(pol, 300, obs_dim, 1, rng, discrete_actions=nA)
    rec = {"trick": np.array([int(trick[k]) for k in sorted(trick)])}
    rec.update(sd_np(pol.agent.Qnet, "init/q/"))
    b = pol.buffer.buffer if trick["PER"] else pol.buffer
    n = b._size
    rec["init/index"] = np.array([b._index, b._size])
    rec["buf/obs"], rec["buf/act"], rec["buf/rew"] = b.obs[:n].copy(), b.actions[:n].copy(), b.rewards[:n].copy()
    rec["buf/nobs"], rec["buf/done"] = b.next_obs[:n].copy(), b.dones[:n].copy()
    if trick["PER"]:
        rec["init/tree"] = pol.buffer.sumtree.tree.copy()
    tap = LossTap(pol.agent, ["update_Qnet"])
    for it in range(4):
        before = rng_snapshot()
        pol.learn(B, 0.99, 0.01)
        after = rng_snapshot()
        rng_restore(before)
        if trick["PER"]:
            rec["u/%d" % it] = np.array([np.random.random_sample() for _ in range(B)])
        else:
            rec["idx/%d" % it] = np.random.choice(n, B, replace=False)
        rng_restore(after)
        if trick["PER"]:
            rec["tree/%d" % it] = pol.buffer.sumtree.tree.copy()
        rec.update(sd_np(pol.agent.Qnet, "after/%d/q/" % it))
    rec.update(sd_np(pol.agent.Qnet, "final/q/"))
    rec.update(sd_np(pol.agent.Qnet_target, "final/q_target/"))
    rec["losses"] = np.array([v[0] for _, v in tap.log])
    rec["gamma_used"] = np.array(pol.buffer.n_step_gamma if trick["N_Step"] else 0.99)
    rec["act/obs"] = rng.standard_normal((16, obs_dim)).astype(np.float32)
    rec["act/action"] = np.array([int(pol.select_action(o)) for o in rec["act/obs"]])
    np.savez_compressed(os.path.join(OUT, "dqn_tricks_%s.npz" % name), **rec)
    print(name, "ok", rec["losses"])


if __name__ == "__main__":
    torch.set_num_threads(1)
    for k, v in CASES.items():
        one(k, v)
