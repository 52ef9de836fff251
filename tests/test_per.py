"""Device sum-tree / PER / n-step buffers: bit-exact against fixtures produced by the reference's DQN_file/Buffer.py
and against the numpy oracle on larger random cases (incl. the non-power-of-two heap layout)."""
import numpy as np
import pytest
import torch

from oracle import buffers as ob


def _golden_per(golden, device):
    from freerl_b200.per import N_Step_PER_Buffer, PER_Buffer
    g = golden("buffers")
    for cap in (5, 8, 37, 100):
        p = "per%d/" % cap
        per = PER_Buffer(cap, 3, 1, device)
        n_add = g[p + "obs"].shape[0]
        tr = [(g[p + "obs"][i], g[p + "act"][i], g[p + "rew"][i], g[p + "nobs"][i], g[p + "done"][i]) for i in range(n_add)]
        half = n_add // 2
        for t in tr[:half]:
            per.add(*t)
        B = min(4, len(per))
        idx, w, pri = per.sample_device(B, u=_u_of(g, p + "s1_u", per, B))
        assert np.array_equal(idx.cpu().numpy(), g[p + "s1_idx"])
        np.testing.assert_allclose(w.cpu().numpy(), g[p + "s1_w"], rtol=1e-6)
        per.update_priorities(g[p + "s1_idx"], g[p + "td1"])
        assert np.array_equal(per.sumtree.tree.cpu().numpy(), g[p + "tree_mid"])     # bit-exact float64 heap
        for t in tr[half:]:
            per.add(*t)
        assert np.array_equal(per.sumtree.tree.cpu().numpy(), g[p + "tree_end"])
        assert [per.buffer._index, per.buffer._size] == list(g[p + "index_end"])
        assert float(per.beta) == float(g[p + "beta_end"])
        B = min(6, len(per))
        idx, w, pri = per.sample_device(B, u=_u_of(g, p + "s2_u", per, B))
        assert np.array_equal(idx.cpu().numpy(), g[p + "s2_idx"])
        np.testing.assert_allclose(w.cpu().numpy(), g[p + "s2_w"], rtol=1e-6)
        for k, t in zip(("obs", "act", "rew", "nobs", "done"), per.buffer.sample(idx)):
            assert np.array_equal(t.cpu().numpy(), g[p + "s2_" + k])
    nb = N_Step_PER_Buffer(16, 2, 1, device, gamma=0.9)
    for i in range(g["nstep/obs_in"].shape[0]):
        nb.add(g["nstep/obs_in"][i], g["nstep/act_in"][i], float(g["nstep/rew_in"][i]), g["nstep/nobs_in"][i], bool(g["nstep/done_in"][i]))
    assert [nb.buffer._index, nb.buffer._size] == list(g["nstep/size"])
    assert np.array_equal(nb.sumtree.tree.cpu().numpy(), g["nstep/tree"])
    n = int(g["nstep/size"][1])
    got = nb.buffer.sample(np.arange(n))
    assert np.array_equal(got[0].cpu().numpy(), g["nstep/obs"][:n].astype(np.float32))
    assert np.array_equal(got[2].cpu().numpy().reshape(-1), g["nstep/rew"][:n].astype(np.float32))
    assert np.array_equal(got[3].cpu().numpy(), g["nstep/nobs"][:n].astype(np.float32))
    assert np.array_equal(got[4].cpu().numpy().reshape(-1), g["nstep/done"][:n].astype(np.float32))


def _u_of(g, key, per, B):
    """recover the unit uniforms the reference consumed: s = a + (b-a)*u with a = seg*i"""
    seg = per.sumtree.sum() / B
    s = g[key]
    return np.array([(s[i] - seg * i) / (seg * (i + 1) - seg * i) for i in range(B)])


def _random_vs_oracle(device, cap, B, rounds):
    """legacy-RNG parity: same np.random seed on both sides -> identical indices, weights and heap."""
    from freerl_b200.per import PER_Buffer
    rng = np.random.default_rng(cap)
    ours, orc = PER_Buffer(cap, 4, 1, device), ob.PrioritizedReplay(cap, 4, 1)
    n = int(cap * 1.3)
    o, a = rng.standard_normal((n, 4)).astype(np.float32), rng.integers(0, 3, (n, 1))
    r, o2, d = rng.standard_normal(n), rng.standard_normal((n, 4)).astype(np.float32), rng.random(n) < 0.1
    for i in range(0, n, 50):
        ours.add(o[i:i + 50], a[i:i + 50], r[i:i + 50], o2[i:i + 50], d[i:i + 50])     # vectorised add of 50 rows
        for j in range(i, min(i + 50, n)):
            orc.add(o[j], a[j], r[j], o2[j], d[j])
    assert np.array_equal(ours.sumtree.tree.cpu().numpy(), orc.sumtree.tree)
    for k in range(rounds):
        np.random.seed(100 + k)
        idx_o, w_o = orc.sample(B)
        np.random.seed(100 + k)
        idx, w = ours.sample(B)
        assert np.array_equal(idx, idx_o)
        np.testing.assert_allclose(w.cpu().numpy(), w_o, rtol=1e-6)
        td = rng.standard_normal((B, 1)).astype(np.float32)
        orc.update_priorities(idx_o, td)
        ours.update_priorities(idx, td)
        assert np.array_equal(ours.sumtree.tree.cpu().numpy(), orc.sumtree.tree), "round %d" % k
        assert ours.sumtree.max() == orc.sumtree.max_leaf()


def test_per_golden_emulated(golden, emul):
    _golden_per(golden, torch.device("cpu"))


def test_per_random_emulated(emul):
    _random_vs_oracle(torch.device("cpu"), 1000, 64, 4)       # non-power-of-two capacity (rotated leaf order)
    _random_vs_oracle(torch.device("cpu"), 256, 32, 3)


@pytest.mark.gpu
def test_per_golden_gpu(golden):
    _golden_per(golden, torch.device("cuda"))


@pytest.mark.gpu
def test_per_random_gpu():
    _random_vs_oracle(torch.device("cuda"), 100000, 256, 4)   # reference-style 1e5 (non power of two), B = 256
    _random_vs_oracle(torch.device("cuda"), 1 << 14, 256, 3)


def test_vectorised_nstep_fold_matches_per_env_folds():
    """N lock-stepped envs folded at once (numpy float64) == N reference-style python folds, bit for bit."""
    from collections import deque
    from freerl_b200.per import _NStepMixin, _fold

    class W(_NStepMixin):
        pass
    rng = np.random.default_rng(0)
    w = W()
    w._init_nstep(0.9, 3)
    ref_win = [deque(maxlen=3) for _ in range(5)]
    emitted = 0
    for _ in range(9):
        o, ac = rng.standard_normal((5, 2)), rng.integers(0, 3, (5, 1))
        r, o2, d = rng.standard_normal(5), rng.standard_normal((5, 2)), rng.random(5) < 0.3
        out = w._push(o, ac, r, o2, d, 2)
        ref = []
        for e in range(5):
            ref_win[e].append((o[e], ac[e], float(r[e]), o2[e], bool(d[e])))
            if len(ref_win[e]) == 3:
                ref.append(_fold(ref_win[e], 0.9))
        if out is None:
            assert not ref
            continue
        emitted += 1
        np.testing.assert_array_equal(out[0], np.stack([x[0] for x in ref]))
        np.testing.assert_array_equal(out[2], np.array([x[2] for x in ref]))
        np.testing.assert_array_equal(out[3], np.stack([x[3] for x in ref]))
        np.testing.assert_array_equal(out[4], np.array([x[4] for x in ref]))
    assert emitted == 7


def test_per_vectorised_add_larger_than_one_launch(emul):
    """a 2500-row vectorised add (3 update launches) == 2500 sequential reference adds, bit for bit"""
    from freerl_b200.per import PER_Buffer
    rng = np.random.default_rng(3)
    cap = 3000
    ours, orc = PER_Buffer(cap, 3, 1, torch.device("cpu")), ob.PrioritizedReplay(cap, 3, 1)
    for n in (2500, 1200):                                    # the second add wraps the ring
        o, a = rng.standard_normal((n, 3)).astype(np.float32), rng.integers(0, 3, (n, 1))
        r, o2, d = rng.standard_normal(n), rng.standard_normal((n, 3)).astype(np.float32), rng.random(n) < 0.1
        ours.add(o, a, r, o2, d)
        for j in range(n):
            orc.add(o[j], a[j], r[j], o2[j], d[j])
        assert np.array_equal(ours.sumtree.tree.cpu().numpy(), orc.sumtree.tree)
        assert (ours.buffer._index, len(ours)) == (orc.buffer._index, len(orc))


def _sumtree_property(device):
    """Random capacities (leaves at one or two depths), random batches WITH duplicate leaves, priorities spanning 12 orders of
    magnitude (so fp64 sums are NOT exact and the batch order matters): the device heap equals the sequential reference
    `tree[idx] += change` walk bit for bit after every batch."""
    from freerl_b200.per import SumTree
    rng = np.random.default_rng(11)
    for cap in (1, 2, 3, 5, 6, 7, 8, 9, 31, 33, 100, 257, 1000, 4097):
        ours, orc = SumTree(cap, device), ob.SumTreeOracle(cap)
        for _ in range(4):
            B = int(rng.integers(1, min(300, 4 * cap) + 1))
            idx = rng.integers(0, cap, B)
            pri = (10.0 ** rng.uniform(-6, 6, B)).astype(np.float32)
            for i, p in zip(idx, pri):
                orc.set_leaf(int(i), float(p))                       # float(np.float32) widens exactly, like the kernel
            ours._update(torch.from_numpy(idx.astype(np.int64)).to(device), pri32=torch.from_numpy(pri).to(device))
            assert np.array_equal(ours.tree.cpu().numpy(), orc.tree), (cap, B)
        assert ours.max() == orc.max_leaf() and ours.sum() == orc.total()


def test_sumtree_ordered_update_property_emulated(emul):
    _sumtree_property(torch.device("cpu"))


@pytest.mark.gpu
def test_sumtree_ordered_update_property_gpu():
    _sumtree_property(torch.device("cuda"))


def test_sumtree_update_by_tree_index_emulated(emul):
    """``SumTree.update(tree_idx, p)`` (DQN_file/Buffer.py:157-166) next to ``add(buffer_idx, p)``: same heap as the oracle's."""
    from freerl_b200.per import SumTree
    from oracle import buffers as ob
    cap = 11
    ours, orc = SumTree(cap, torch.device("cpu")), ob.SumTreeOracle(cap)
    rng = np.random.default_rng(0)
    for _ in range(30):
        leaf, p = int(rng.integers(0, cap)), float(rng.random() * 3)
        if rng.random() < 0.5:
            ours.update(leaf + cap - 1, p)
        else:
            ours.add(leaf, p)
        orc.set_leaf(leaf, p)
    assert np.array_equal(ours.tree.cpu().numpy(), orc.tree)
    with pytest.raises(IndexError):
        ours.update(0, 1.0)          # the root is not a leaf: the reference would corrupt the heap silently
