"""``MAPPO_file/MAPPO_discrete.py`` (shared networks, episode ``ReplayBuffer``, joint clip, double optimiser step):
the oracle against fixtures generated from the unmodified reference, then ``freerl_b200.MAPPO_discrete`` (fused kernels) against both."""
import numpy as np
import pytest
import torch

from oracle.make_golden_mappo_discrete import AD, B, HP, K, MB, N, OD, T, TRICKS
from oracle.mappo_discrete import MAPPODiscreteOracle
from parity_util import assert_module_close, load_into, net_from_golden

IDS = ["agent_%d" % i for i in range(N)]


def _batch(g):
    return {k: torch.from_numpy(g["buf/" + k].copy()) for k in ("obs_n", "s", "v_n", "a_n", "a_logprob_n", "r_n", "done_n")}


def _oracle(g, name):
    orc = MAPPODiscreteOracle(net_from_golden(g, "init/actor/"), net_from_golden(g, "init/critic/"), N, OD, AD, 1e-3, TRICKS[name])
    r = orc.learn(_batch(g), MB, HP["gamma"], HP["lmbda"], HP["clip_param"], K, HP["entropy_coefficient"], HP["huber_delta"])
    return orc, r


@pytest.mark.parametrize("name", ["simple", "clip", "full"])
def test_oracle_vs_reference_fixture(golden, name):
    """pins the restatement: losses 1e-6, final parameters 1e-6 of the reference's, Adam stepped twice per minibatch"""
    g = golden("mappo_discrete_" + name)
    orc, r = _oracle(g, name)
    tot = np.array([a + c for a, c in r["losses"]])
    np.testing.assert_allclose(tot, g["losses"], rtol=1e-6, atol=1e-6)
    for kind, mod in (("actor", orc.actor), ("critic", orc.critic)):
        for k, v in net_from_golden(g, "final/%s/" % kind).items():
            np.testing.assert_allclose(mod.state_dict()[k].numpy(), v.numpy(), rtol=1e-6, atol=1e-7, err_msg=kind + "/" + k)
    assert int(orc.opt.state[orc.params[0]]["step"]) == int(g["final/adam_step"]) == 2 * len(g["losses"])


def _run(golden, device, name, parity_draws):
    from freerl_b200.MAPPO_discrete import MAPPO, ReplayBuffer
    g = golden("mappo_discrete_" + name)
    dim_info = {k: [OD, AD] for k in IDS}
    buf = ReplayBuffer(N=N, obs_dim=OD, state_dim=N * OD, episode_limit=T, batch_size=B, device=device)
    pol = MAPPO(dim_info, False, 1e-3, 5e-4, B, device, dict(TRICKS[name]), buf)
    load_into(pol.agent.actor, net_from_golden(g, "init/actor/"))
    load_into(pol.agent.critic, net_from_golden(g, "init/critic/"))
    # the rollout, stored through add(): select_action reproduces the reference's own draws (actions exactly, log-probs to 2e-5) and
    # get_value the values it stored
    torch.set_rng_state(torch.from_numpy(g["rng/before_rollout"].copy()))
    obs_n, r_n, done_n = g["buf/obs_n"], g["buf/r_n"], g["buf/done_n"]
    i = 0
    for ep in range(B):
        for t in range(T):
            if parity_draws:
                a, lp = pol.select_action(list(obs_n[ep, t]))
            else:           # GPU: the reference's host draws, injected (Categorical.sample -> multinomial's exponential_(1) over [N, A])
                a, lp = pol.select_action(list(obs_n[ep, t]), noise=torch.empty((N, AD)).exponential_(1))
            assert a.tolist() == g["data/act"][i].tolist(), ("action", ep, t)
            np.testing.assert_allclose(lp, g["data/logp"][i], rtol=2e-5, atol=2e-6)
            pol.add(list(obs_n[ep, t]), g["data/act"][i], {k: float(r_n[ep, t, j]) for j, k in enumerate(IDS)}, None,
                    {k: bool(done_n[ep, t, j]) for j, k in enumerate(IDS)}, g["data/logp"][i], None, t)
            i += 1
        buf.store_last_value(T, pol.get_value(g["data/last_state"][ep]))
    assert buf.episode_num == B
    np.testing.assert_allclose(buf.buffer["v_n"], g["buf/v_n"], rtol=1e-5, atol=2e-6)
    for k in ("obs_n", "s", "a_n", "a_logprob_n", "r_n", "done_n"):
        np.testing.assert_array_equal(buf.buffer[k], g["buf/" + k], err_msg=k)
    buf.buffer["v_n"][...] = g["buf/v_n"]            # learn from exactly the reference's rollout
    orc, r = _oracle(g, name)
    pol.learn(MB, HP["gamma"], HP["lmbda"], HP["clip_param"], K, HP["entropy_coefficient"], HP["huber_delta"])
    assert buf.episode_num == 0 and pol.agent.step == int(g["final/adam_step"])
    # advantages: the product keeps time-major rows (t, b, n); the oracle / reference [b, t, n]
    tm = lambda x: x.reshape(T, B, N).permute(1, 0, 2).cpu().numpy()
    np.testing.assert_allclose(tm(pol.last_adv), r["adv"].numpy(), rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(tm(pol.last_v_target), r["v_target"].numpy(), rtol=1e-5, atol=2e-6)
    m = pol.last_metrics.cpu().numpy()
    ref = np.array(r["losses"])
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 0] + m[:, 1], g["losses"], rtol=1e-5, atol=2e-6)
    tol = dict(rtol=5e-5, atol=5e-6)
    assert_module_close(pol.agent.actor, orc.actor.state_dict(), "actor vs oracle", tol)
    assert_module_close(pol.agent.critic, orc.critic.state_dict(), "critic vs oracle", tol)
    assert_module_close(pol.agent.actor, net_from_golden(g, "final/actor/"), "actor vs reference", tol)
    assert_module_close(pol.agent.critic, net_from_golden(g, "final/critic/"), "critic vs reference", tol)
    assert pol.evaluate_action(g["eval/obs"]).tolist() == g["eval/action"].tolist()
    np.testing.assert_allclose(pol.get_value(g["eval/obs"].flatten()), g["eval/value"], rtol=1e-4, atol=1e-5)
    # on-disk format (MAPPO_discrete.py:389-393): the actor's state dict under MAPPO_discrete.pth, loadable by the reference's module
    import os, tempfile
    with tempfile.TemporaryDirectory() as d:
        pol.save(d)
        data = torch.load(os.path.join(d, "MAPPO_discrete.pth"))
        assert list(data) == [k[len("final/actor/"):] for k in g.files if k.startswith("final/actor/")]
        back = MAPPO.load(dim_info, False, d, trick=dict(TRICKS[name]), device=device)
        assert back.evaluate_action(g["eval/obs"]).tolist() == g["eval/action"].tolist()


@pytest.mark.parametrize("name", ["simple", "clip", "full"])
def test_mappo_discrete_shared_emulated(golden, emul, name):
    _run(golden, torch.device("cpu"), name, True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["simple", "clip", "full"])
def test_mappo_discrete_shared_gpu(golden, name):
    _run(golden, torch.device("cuda"), name, False)


def _big(device, name="clip", seed=5, dims=None):
    """MPE-sized learn (25 steps, 3 agents, 32 episodes, minibatches of 16 episodes = 1200 rows) on a synthetic rollout, product vs
    oracle.  "clip": the tensor-core path on a GPU; "full" (the script's default switches): group mode with 75-row LayerNorm groups —
    nine full 8-row tiles and one of three rows — and the scalar huber / ValueClip loss"""
    from freerl_b200.MAPPO_discrete import MAPPO, ReplayBuffer
    N, OD, AD, T_, B_, MB_ = dims or (3, 18, 5, 25, 32, 16)          # agents, obs, actions, episode_limit, episodes, minibatch (episodes)
    IDS = ["agent_%d" % i for i in range(N)]
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    trick = dict(TRICKS[name])
    buf = ReplayBuffer(N=N, obs_dim=OD, state_dim=N * OD, episode_limit=T_, batch_size=B_, device=device)
    pol = MAPPO({k: [OD, AD] for k in IDS}, False, 1e-3, 5e-4, B_, device, trick, buf)
    sd = lambda mod: {k: v.detach().cpu().clone() for k, v in mod.state_dict().items()}
    orc = MAPPODiscreteOracle(sd(pol.agent.actor), sd(pol.agent.critic), N, OD, AD, 1e-3, trick)
    b = buf.buffer
    b["obs_n"][...] = rng.standard_normal(b["obs_n"].shape)
    b["s"][...] = b["obs_n"].reshape(B_, T_, N * OD)
    b["v_n"][...] = 0.5 * rng.standard_normal(b["v_n"].shape)
    b["a_n"][...] = rng.integers(0, AD, b["a_n"].shape)
    b["a_logprob_n"][...] = np.log(1.0 / AD) + 0.1 * rng.standard_normal(b["a_n"].shape)
    b["r_n"][...] = rng.standard_normal(b["r_n"].shape)
    b["done_n"][...] = rng.random(b["done_n"].shape) < 0.05
    buf.episode_num = B_
    batch = {k: torch.from_numpy(v.copy()) for k, v in b.items()}
    r = orc.learn(batch, MB_, 0.95, 0.95, 0.2, 2, 0.01, 1.0)
    pol.learn(MB_, 0.95, 0.95, 0.2, 2, 0.01, 1.0)          # huber_delta 1: both huber branches occur
    m = pol.last_metrics.cpu().numpy()
    ref = np.array(r["losses"])
    np.testing.assert_allclose(m[:, 0], ref[:, 0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m[:, 1], ref[:, 1], rtol=1e-5, atol=2e-6)
    tol = dict(rtol=1e-4, atol=1e-5)
    assert_module_close(pol.agent.actor, orc.actor.state_dict(), "actor vs oracle", tol)
    assert_module_close(pol.agent.critic, orc.critic.state_dict(), "critic vs oracle", tol)


@pytest.mark.parametrize("name", ["clip", "full"])
def test_mappo_discrete_mpe_size_emulated(emul, name):
    _big(torch.device("cpu"), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["clip", "full"])
def test_mappo_discrete_mpe_size_gpu(name):
    _big(torch.device("cuda"), name)


# odd shapes: two agents / 10-row groups (one full tile + two rows) with a ragged last minibatch; five agents with a 150-wide joint
# observation (wider than the tensor-core path takes) and 35-row groups; a group of exactly one tile
ODD = [(2, 7, 3, 5, 6, 4), (5, 30, 4, 7, 8, 3), (2, 6, 2, 4, 5, 5)]


@pytest.mark.parametrize("dims", ODD, ids=["n2_g10", "n5_g35", "n2_g8"])
@pytest.mark.parametrize("name", ["clip", "full"])
def test_mappo_discrete_odd_shapes_emulated(emul, name, dims):
    _big(torch.device("cpu"), name, seed=9, dims=dims)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", ODD, ids=["n2_g10", "n5_g35", "n2_g8"])
@pytest.mark.parametrize("name", ["clip", "full"])
def test_mappo_discrete_odd_shapes_gpu(name, dims):
    _big(torch.device("cuda"), name, seed=9, dims=dims)


def test_unreproduced_switches_raise(emul):
    from freerl_b200.MAPPO_discrete import MAPPO, ReplayBuffer
    dev = torch.device("cpu")
    buf = ReplayBuffer(N=N, obs_dim=OD, state_dim=N * OD, episode_limit=T, batch_size=B, device=dev)
    with pytest.raises(NotImplementedError):
        MAPPO({k: [OD, AD] for k in IDS}, True, 1e-3, 5e-4, B, dev, dict(TRICKS["simple"]), buf)
