"""Stochastic ``select_action`` against the REFERENCE's own numbers: the goldens hold the actions and
log-probabilities the unmodified reference returned during its rollouts (``data/act``, ``data/logp``) and the CPU generator state
those calls started from (``rng/before_rollout``; SAC: ``sel/*`` after its learns).
  * emulated (CPU): parity mode draws by itself — the generator is restored and the calls must reproduce the reference values,
    which also pins the claim that parity mode consumes torch's generator in the reference's order;
  * GPU: the same draws are made on the host in that order and injected (``noise=``), the kernels compute action and log-prob.
SAC ``tanh(rsample)`` SAC_file/SAC.py:192-198, PPO PPO_file/PPO.py:168-182, MAPPO MAPPO_file/MAPPO.py:305-320."""
import numpy as np
import pytest
import torch

from parity_util import load_into, net_from_golden

TOL = dict(rtol=1e-5, atol=2e-6)


def _restore(g, key):
    torch.set_rng_state(torch.from_numpy(g[key].copy()))


def _ppo(golden, device, name, is_continue, inject):
    from freerl_b200.PPO import PPO
    g = golden(name)
    ad = 2 if is_continue else 4
    pol = PPO([8, ad], is_continue, 1e-3, 1e-3, 256, device)
    load_into(pol.agent.actor, net_from_golden(g, "init/actor/"))
    load_into(pol.agent.critic, net_from_golden(g, "init/critic/"))
    obs, act, logp = g["data/obs"], g["data/act"], g["data/logp"]
    _restore(g, "rng/before_rollout")
    for t in range(256):
        if inject:      # what dist.sample() consumes: Normal -> empty(shape).normal_(); Categorical -> multinomial's exponential_(1)
            z = torch.empty((1, ad)).normal_() if is_continue else torch.empty((1, ad)).exponential_(1)
            a, lp = pol.select_action(obs[t], noise=z)
        else:
            a, lp = pol.select_action(obs[t])
        if is_continue:
            np.testing.assert_allclose(a, act[t], err_msg="action, step %d" % t, **TOL)
        else:
            assert int(a) == int(act[t].reshape(-1)[0]), ("action, step %d" % t, a, act[t])
        np.testing.assert_allclose(np.asarray(lp).reshape(-1), logp[t].reshape(-1), err_msg="log-prob, step %d" % t, rtol=2e-5, atol=2e-6)


def _sac(golden, device, inject):
    from freerl_b200.SAC import SAC
    g = golden("sac")
    trick = {"ObsNorm": False, "Batch_ObsNorm": False, "OUNoise": True, "GaussNoise": False}
    pol = SAC([17, 6], True, 1e-3, 1e-3, 1000, device, trick=trick)
    load_into(pol.agent.actor, net_from_golden(g, "final/actor/"))           # the reference sampled after its three learns
    _restore(g, "sel/rng_state")
    for i in range(8):
        if inject:
            a = pol.select_action(g["sel/obs"][i], noise=torch.empty((1, 6)).normal_())
        else:
            a = pol.select_action(g["sel/obs"][i])
        np.testing.assert_allclose(a, g["sel/action"][i], err_msg="row %d" % i, **TOL)


def _mappo(golden, device, name, is_continue, inject):
    from freerl_b200.MAPPO import MAPPO
    from oracle.make_golden_marl import MAPPO_TRICK          # the trick dict only
    g = golden(name)
    ids = ["agent_%d" % i for i in range(3)]
    pol = MAPPO({k: [18, 5] for k in ids}, is_continue, 1e-3, 1e-3, 64, device, dict(MAPPO_TRICK))
    for k in ids:
        load_into(pol.agents[k].actor, net_from_golden(g, "init/%s/actor/" % k))
        load_into(pol.agents[k].critic, net_from_golden(g, "init/%s/critic/" % k))
    _restore(g, "rng/before_rollout")
    for t in range(64):
        obs = {k: g["data/%s/obs" % k][t] for k in ids}
        if inject:      # the reference loops over the agents in dict order, one draw per agent (MAPPO.py:305-320)
            z = {k: (torch.empty((1, 5)).normal_() if is_continue else torch.empty((1, 5)).exponential_(1)) for k in ids}
            a, lp = pol.select_action(obs, noise=z)
        else:
            a, lp = pol.select_action(obs)
        for k in ids:
            if is_continue:
                np.testing.assert_allclose(a[k], g["data/%s/act" % k][t], err_msg="%s action, step %d" % (k, t), **TOL)
            else:
                assert int(a[k]) == int(g["data/%s/act" % k][t].reshape(-1)[0]), (k, t)
            np.testing.assert_allclose(np.asarray(lp[k]).reshape(-1), g["data/%s/logp" % k][t].reshape(-1),
                                       err_msg="%s log-prob, step %d" % (k, t), rtol=2e-5, atol=2e-6)


def test_select_action_values_emulated(golden, emul):
    dev = torch.device("cpu")
    _ppo(golden, dev, "ppo_cont", True, False)
    _ppo(golden, dev, "ppo_disc", False, False)
    _sac(golden, dev, False)
    _mappo(golden, dev, "mappo", True, False)
    _mappo(golden, dev, "mappo_disc", False, False)


@pytest.mark.gpu
def test_select_action_values_gpu(golden):
    dev = torch.device("cuda")
    _ppo(golden, dev, "ppo_cont", True, True)
    _ppo(golden, dev, "ppo_disc", False, True)
    _sac(golden, dev, True)
    _mappo(golden, dev, "mappo", True, True)
    _mappo(golden, dev, "mappo_disc", False, True)
