"""The oracle (oracle/) replayed against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  This pins the checker; it runs on CPU."""
from collections import OrderedDict

import numpy as np
import torch

from oracle import algos, buffers

TOL = dict(rtol=2e-5, atol=2e-6)


def net(g, prefix):
    out = OrderedDict()
    for k in g.files:
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(g[k].copy())
    assert out, prefix
    return out


def batch(g, it):
    return tuple(torch.from_numpy(g["batch/%d/%s" % (it, k)]) for k in ("obs", "act", "rew", "nobs", "done"))


def losses(g, name):
    return [g[k] for k in sorted(g.files) if k.startswith("loss/") and k.endswith(name)]


def assert_net(a, g, prefix):
    for k, v in a.items():
        np.testing.assert_allclose(v.detach().numpy(), g[prefix + k], err_msg=prefix + k, **TOL)


def test_dqn(golden):
    g = golden("dqn")
    o = algos.DQNOracle(net(g, "init/q/"), 1e-3)
    ls = losses(g, "update_Qnet")
    for it in range(3):
        r = o.learn(batch(g, it), 0.99, 0.01)
        np.testing.assert_allclose(r["loss"], ls[it][0], rtol=1e-6)
    assert_net(o.q, g, "final/q/")
    assert_net(o.q_target, g, "final/q_target/")


def test_sac(golden):
    g = golden("sac")
    o = algos.SACOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3, act_dim=6)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    for it in range(3):
        r = o.learn(batch(g, it), torch.from_numpy(g["noise/%d/0" % it]), torch.from_numpy(g["noise/%d/1" % it]), 0.99, 0.01)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        np.testing.assert_allclose(r["actor_loss"], la[it][0], rtol=1e-5)
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)
    np.testing.assert_allclose(o.log_alpha.item(), float(g["final/log_alpha"]), rtol=1e-6)


def test_td3(golden):
    g = golden("td3")
    o = algos.TD3Oracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    na = 0
    for it in range(4):
        r = o.learn(batch(g, it), torch.from_numpy(g["noise/%d/0" % it]), 0.99, 0.01, 0.1, 0.5, 1.0, 2, 1.0)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        if "actor_loss" in r:
            np.testing.assert_allclose(r["actor_loss"], la[na][0], rtol=1e-5)
            na += 1
    assert na == 2
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)


def test_ddpg(golden):
    g = golden("ddpg")
    o = algos.DDPGOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3, weight_decay=True)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    for it in range(3):
        r = o.learn(batch(g, it), 0.99, 0.01)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        np.testing.assert_allclose(r["actor_loss"], la[it][0], rtol=1e-5)
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)


def test_sac_batch_obs_norm(golden):
    g = golden("sac_bon")
    bon = algos.BatchObsNorm(17)
    o = algos.SACOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3, act_dim=6, obs_norm=bon)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    for it in range(4):
        r = o.learn(batch(g, it), torch.from_numpy(g["noise/%d/0" % it]), torch.from_numpy(g["noise/%d/1" % it]), 0.99, 0.01)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        np.testing.assert_allclose(r["actor_loss"], la[it][0], rtol=1e-5)
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)
    assert bon.n == int(g["final/norm/n"]) == 4
    np.testing.assert_array_equal(bon.mean.numpy(), g["final/norm/mean"])
    np.testing.assert_array_equal(bon.std.numpy(), g["final/norm/std"])


def test_ddpg_batch_obs_norm(golden):
    g = golden("ddpg_bon")
    bon = algos.BatchObsNorm(17)
    o = algos.DDPGOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3, weight_decay=True, obs_norm=bon)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    for it in range(4):
        r = o.learn(batch(g, it), 0.99, 0.01)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        np.testing.assert_allclose(r["actor_loss"], la[it][0], rtol=1e-5)
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)
    np.testing.assert_array_equal(bon.std.numpy(), g["final/norm/std"])
    # select_action with update=False (DDPG.py:166-173)
    a = algos.tanh_actor(o.actor, bon(torch.from_numpy(g["act/obs"]).reshape(1, -1), update=False))
    np.testing.assert_allclose(a.detach().numpy()[0], g["act/action"], rtol=1e-5, atol=1e-6)


def _ppo(golden, name, is_continue):
    g = golden(name)
    o = algos.PPOOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, is_continue)
    data = tuple(torch.from_numpy(g["data/" + k]) for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
    perms = [g["perm/%d" % k] for k in range(2)]
    r = o.learn(data, perms, 64, 0.99, 0.95, 0.2, 0.01)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=2e-5, atol=1e-6)
    assert_net(o.actor, g, "final/actor/")
    assert_net(o.critic, g, "final/critic/")


def test_ppo_continuous(golden):
    _ppo(golden, "ppo_cont", True)


def test_ppo_discrete(golden):
    _ppo(golden, "ppo_disc", False)


def test_ddpg_simple(golden):
    """DDPG_file/DDPG_simple.py = DDPG without supplements, replayed through DDPGOracle(weight_decay=False)"""
    g = golden("ddpg_simple")
    o = algos.DDPGOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 1e-3, weight_decay=False)
    lc, la = losses(g, "update_critic"), losses(g, "update_actor")
    for it in range(3):
        r = o.learn(batch(g, it), 0.99, 0.01)
        np.testing.assert_allclose(r["critic_loss"], lc[it][0], rtol=1e-6)
        np.testing.assert_allclose(r["actor_loss"], la[it][0], rtol=1e-5)
    for n in ("actor", "critic", "actor_target", "critic_target"):
        assert_net(getattr(o, n), g, "final/%s/" % n)


def _ppo_advance(golden, name, is_continue):
    """PPO_advance/PPO.py (separate Adams, probs head) replayed through oracle.algos.PPOAdvanceOracle"""
    g = golden(name)
    o = algos.PPOAdvanceOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 5e-4, is_continue)
    data = tuple(torch.from_numpy(g["data/" + k]) for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
    r = o.learn(data, [g["perm/%d" % k] for k in range(2)], 64, 0.99, 0.95, 0.2, 0.01)
    np.testing.assert_allclose(np.array(r["losses"]), g["losses"], rtol=2e-5, atol=1e-6)
    assert_net(o.actor, g, "final/actor/")
    assert_net(o.critic, g, "final/critic/")


def test_ppo_advance_continuous(golden):
    _ppo_advance(golden, "ppo_adv_cont", True)


def test_ppo_advance_discrete(golden):
    _ppo_advance(golden, "ppo_adv_disc", False)


def test_buffers_per_and_tree(golden):
    g = golden("buffers")
    for cap in (5, 8, 37, 100):
        p = "per%d/" % cap
        per = buffers.PrioritizedReplay(cap, 3, 1)
        n_add = g[p + "obs"].shape[0]
        tr = [(g[p + "obs"][i], g[p + "act"][i], g[p + "rew"][i], g[p + "nobs"][i], g[p + "done"][i]) for i in range(n_add)]
        np.random.seed(cap)
        half = n_add // 2
        for t in tr[:half]:
            per.add(*t)
        B = min(4, len(per))
        idx, w = per.sample(B)
        assert np.array_equal(idx, g[p + "s1_idx"])
        assert np.array_equal(w, g[p + "s1_w"])
        per.update_priorities(idx, g[p + "td1"])
        assert np.array_equal(per.sumtree.tree, g[p + "tree_mid"])            # bit-exact float64
        for t in tr[half:]:
            per.add(*t)
        assert np.array_equal(per.sumtree.tree, g[p + "tree_end"])
        assert [per.buffer._index, per.buffer._size] == list(g[p + "index_end"])
        assert float(per.beta) == float(g[p + "beta_end"])
        B = min(6, len(per))
        idx, w = per.sample(B)
        assert np.array_equal(idx, g[p + "s2_idx"]) and np.array_equal(w, g[p + "s2_w"])
        for k, t in zip(("obs", "act", "rew", "nobs", "done"), per.buffer.sample(idx)):
            assert np.array_equal(t, g[p + "s2_" + k])


def test_buffers_nstep(golden):
    g = golden("buffers")
    nb = buffers.NStepPrioritizedReplay(16, 2, 1, gamma=0.9)
    for i in range(g["nstep/obs_in"].shape[0]):
        nb.add(g["nstep/obs_in"][i], g["nstep/act_in"][i], float(g["nstep/rew_in"][i]), g["nstep/nobs_in"][i],
               bool(g["nstep/done_in"][i]))
    b = nb.buffer
    assert np.array_equal(b.obs, g["nstep/obs"]) and np.array_equal(b.actions, g["nstep/act"])
    assert np.array_equal(b.rewards, g["nstep/rew"]) and np.array_equal(b.next_obs, g["nstep/nobs"])
    assert np.array_equal(b.dones, g["nstep/done"])
    assert [b._index, b._size] == list(g["nstep/size"])
    assert np.array_equal(nb.sumtree.tree, g["nstep/tree"])


def test_choice_stream(golden):
    g = golden("buffers")
    np.random.seed(0)
    assert np.array_equal(buffers.uniform_indices(1000, 8), g["choice/seed0_1000_8"])
    assert np.array_equal(buffers.uniform_indices(50, 50), g["choice/seed0_next_50_50"])


def _ppo_tricks(golden, name, is_continue, tanh=False):
    """PPO_file/PPO_with_tricks.py (adv_norm + orthogonal_init + adam_eps + lr_decay; np.zeros call patched at generation time) replayed
    through oracle.algos.PPOTricksOracle: two learns with lr_decay(10, 100) after each"""
    g = golden(name)
    o = algos.PPOTricksOracle(net(g, "init/actor/"), net(g, "init/critic/"), 1e-3, 5e-4, is_continue, adam_eps=True, adv_norm=True, tanh=tanh)
    losses_ = []
    for r in range(2):
        data = tuple(torch.from_numpy(g["data%d/%s" % (r, k)]) for k in ("obs", "act", "rew", "nobs", "done", "logp", "adv_done"))
        losses_ += o.learn(data, [g["perm%d/%d" % (r, k)] for k in range(2)], 64, 0.99, 0.95, 0.2, 0.01)["losses"]
        o.lr_decay(10, 100)
        assert_net(o.actor, g, "after%d/actor/" % r)
        assert_net(o.critic, g, "after%d/critic/" % r)
    np.testing.assert_allclose(np.array(losses_), g["losses"], rtol=2e-5, atol=1e-6)


def test_ppo_with_tricks_continuous(golden):
    _ppo_tricks(golden, "ppo_tricks_cont", True)


def test_ppo_with_tricks_discrete(golden):
    _ppo_tricks(golden, "ppo_tricks_disc", False)


def test_ppo_with_tricks_tanh(golden):
    _ppo_tricks(golden, "ppo_tricks_tanh_cont", True, tanh=True)
    _ppo_tricks(golden, "ppo_tricks_tanh_disc", False, tanh=True)       # Actor_discrete keeps ReLU, the critic switches
