"""Measured fp32 summation-order bound for the loss tolerances.

A learn step is mathematically invariant under a permutation of the batch rows; in fp32 a permutation only changes the ORDER of the
batch sums (loss means, dW = dY^T X, bias gradients).  Running the oracle — the checker itself, on the CPU — on row-permuted copies of
the reference-generated batches therefore measures how far two equally valid fp32 evaluations of the SAME reference arithmetic sit
apart; a CUDA kernel (another summation order again) cannot be asked to sit closer to the oracle than the oracle sits to itself.
The spreads measured here (committed: ``profiles/r4_reorder_bound.json``, written by ``python tests/test_reorder_bound.py``) are what
the parity tolerances are held against: 1e-5 on critic / value losses, 2e-5 on actor and surrogate losses (``tests/test_parity_ac.py``,
``tests/test_parity_ppo.py``).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import algos  # noqa: E402
from parity_util import golden_batch, net_from_golden  # noqa: E402


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def sac_spread(g, n_learn, n_perm, seed=0):
    """max relative loss difference between the oracle on the golden batches and on row-permuted copies, per learn"""
    def run(perms):
        orc = algos.SACOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), 1e-3, 1e-3, act_dim=6)
        out = []
        for it in range(n_learn):
            p = perms[it]
            batch = tuple(x[p] for x in golden_batch(g, it))
            r = orc.learn(batch, torch.from_numpy(g["noise/%d/0" % it])[p], torch.from_numpy(g["noise/%d/1" % it])[p], 0.99, 0.01)
            out.append((r["critic_loss"], r["actor_loss"]))
        return np.array(out)
    B = golden_batch(g, 0)[0].shape[0]
    base = run([torch.arange(B)] * n_learn)
    gen = torch.Generator().manual_seed(seed)
    worst = np.zeros((n_learn, 2))
    for _ in range(n_perm):
        got = run([torch.randperm(B, generator=gen) for _ in range(n_learn)])
        worst = np.maximum(worst, np.abs(got - base) / np.maximum(np.abs(base), 1e-12))
    return base, worst


def _load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def test_sac_reorder_spread_b64():
    """B = 64, three chained learns: the spread exists (fp32 is order-dependent) and stays inside the parity tolerances, which are
    therefore not tighter than the checker's own ambiguity and not more than ~an order of magnitude looser"""
    base, worst = sac_spread(_load("sac"), 3, 12)
    assert worst.max() > 0.0
    assert worst[:, 0].max() < 1e-5 and worst[:, 1].max() < 2e-5, worst


def test_sac_reorder_spread_b256():
    base, worst = sac_spread(_load("sac_b256"), 4, 6)
    assert worst[:, 0].max() < 1e-5 and worst[:, 1].max() < 2e-5, worst


if __name__ == "__main__":
    rec = {"what": "max relative loss difference, oracle vs oracle on row-permuted batches (fp32 summation order only)",
           "tolerances_in_tests": {"critic_loss": 1e-5, "actor_loss": 2e-5}}
    for name, n_learn, n_perm in (("sac", 3, 64), ("sac_b256", 10, 32)):
        base, worst = sac_spread(_load(name), n_learn, n_perm)
        rec[name] = {"learns": n_learn, "permutations": n_perm, "critic_loss_spread_per_learn": worst[:, 0].tolist(),
                     "actor_loss_spread_per_learn": worst[:, 1].tolist(), "critic_max": float(worst[:, 0].max()),
                     "actor_max": float(worst[:, 1].max())}
        print(name, "critic max %.2e actor max %.2e" % (worst[:, 0].max(), worst[:, 1].max()))
    with open(os.path.join(ROOT, "profiles", "r4_reorder_bound.json"), "w") as f:
        json.dump(rec, f, indent=1)
