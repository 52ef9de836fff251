"""Golden fixtures for ``PPO_file/PPO_with_tricks.py`` with tricks adv_norm + orthogonal_init + adam_eps + lr_decay.

    python -m oracle.make_golden_ppo_tricks     # writes tests/golden/ppo_tricks_{cont,disc}.npz and ppo_tricks_tanh_{cont,disc}.npz (the `tanh` switch on)

The reference file is run UNMODIFIED except for one call it cannot execute: ``np.zeros(self.horizon, dtype=torch.float32)``
(``:302``) raises ``TypeError`` under every NumPy.  The module's ``np`` name is rebound to a proxy whose ``zeros`` maps a torch dtype
to the NumPy dtype of the same name (float32 — what the line obviously means: the array is cast to float32 two lines later anyway);
everything else is forwarded to numpy.  Two rollouts / learns with ``lr_decay(10, 100)`` in between.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refload  # noqa: E402
from oracle.make_golden import OUT, LossTap, rng_restore, rng_snapshot, sd_np  # noqa: E402

TRICK = {'adv_norm': True, 'ObsNorm': False, 'Batch_ObsNorm': False, 'reward_norm': False, 'reward_scaling': False,
         'lr_decay': True, 'orthogonal_init': True, 'adam_eps': True, 'tanh': False}


class _NpProxy:
    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def zeros(shape, dtype=float):
        if isinstance(dtype, torch.dtype):
            dtype = getattr(np, str(dtype).split(".")[-1])
        return np.zeros(shape, dtype=dtype)


def gen(is_continue, tanh=False, bon=False, beta=False):
    m = refload.load("PPO_file", "PPO_with_tricks")
    m.np = _NpProxy()
    seed, horizon, mb, K = 13, 256, 64, 2
    obs_dim, act_dim = 8, (2 if is_continue else 4)
    np.random.seed(seed)
    torch.manual_seed(seed)
    policy = m.PPO([obs_dim, act_dim], is_continue, 1e-3, 5e-4, horizon, torch.device("cpu"), trick=dict(TRICK, tanh=tanh, Batch_ObsNorm=bon), beta=beta)
    rng = np.random.default_rng(seed)
    offset = rng.uniform(1.0, 3.0, obs_dim).astype(np.float32) if bon else 0.0      # Batch_ObsNorm fixture: observations with a non-zero mean
    rec = {}
    rec.update(sd_np(policy.agent.actor, "init/actor/"))
    rec.update(sd_np(policy.agent.critic, "init/critic/"))
    tap = LossTap(policy.agent, ["update_actor", "update_critic"])
    obs = rng.standard_normal(obs_dim).astype(np.float32) + offset
    for r in range(2):
        if bon or beta:
            rec["rng%d/before_rollout" % r] = torch.get_rng_state().numpy().copy()
        for t in range(horizon):
            a, logp = policy.select_action(obs)
            o2 = rng.standard_normal(obs_dim).astype(np.float32) + offset
            term = bool(rng.random() < 0.02)
            trunc = (t % 50) == 49
            policy.add(obs, a, float(rng.standard_normal()), o2, term, logp, term or trunc)
            obs = o2
        for k, t in zip(("obs", "act", "rew", "nobs", "done", "logp", "adv_done"), policy.buffer.all()):
            rec["data%d/%s" % (r, k)] = t.numpy().copy()
        before = rng_snapshot()
        policy.learn(mb, 0.99, 0.95, 0.2, K, 0.01)
        after = rng_snapshot()
        rng_restore(before)
        for k in range(K):
            rec["perm%d/%d" % (r, k)] = np.random.permutation(horizon)
        rng_restore(after)
        policy.lr_decay(10, max_episodes=100)
        if bon:
            ms = policy.batch_size_obs_norm.running_ms
            rec["after%d/norm/mean" % r], rec["after%d/norm/std" % r] = ms.mean.numpy().copy(), ms.std.numpy().copy()
        rec.update(sd_np(policy.agent.actor, "after%d/actor/" % r))
        rec.update(sd_np(policy.agent.critic, "after%d/critic/" % r))
    log = [v[0] for _, v in tap.log]
    rec["losses"] = np.array(list(zip(log[0::2], log[1::2])), np.float64)          # (actor, critic) per minibatch, both learns
    name = ("ppo_tricks_beta_" if beta else "ppo_tricks_bon_" if bon else "ppo_tricks_tanh_" if tanh else "ppo_tricks_") + ("cont" if is_continue else "disc")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok", rec["losses"][:2], rec["losses"].shape)


if __name__ == "__main__":
    torch.set_num_threads(1)
    if "beta" in sys.argv[1:]:
        gen(True, beta=True)
        sys.exit(0)
    if "bon" in sys.argv[1:]:
        gen(True, bon=True)
        gen(False, bon=True)
        sys.exit(0)
    gen(True)
    gen(False)
    gen(True, tanh=True)
    gen(False, tanh=True)
