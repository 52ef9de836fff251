"""Ring replay (frl_replay_add_batch / frl_replay_gather behind freerl_b200.Buffer) vs the oracle's numpy restatement of
SAC_file/Buffer.py:14-61 on random capacities, random add sizes (single rows, vector steps, batches larger than the capacity,
wrap-arounds) and random sample indices: stored rows, `_index`, `_size` and the five sampled tensors agree bit for bit."""
import numpy as np
import pytest
import torch

from oracle import buffers as ob


def _run(device, seed):
    from freerl_b200.Buffer import Buffer
    rng = np.random.default_rng(seed)
    for trial in range(8):
        big = trial >= 6                          # several 64-row tiles per launch, ragged last tile
        cap = int(rng.integers(200, 700)) if big else int(rng.integers(1, 90))
        od, ad = int(rng.integers(1, 20)), int(rng.integers(1, 7))
        ours, orc = Buffer(float(cap), od, ad, device), ob.RingReplay(cap, od, ad)
        for _ in range(12):
            n = int(rng.choice([1, 1, 2, 7, cap, cap + 3, 2 * cap + 1, int(rng.integers(1, 40))]))
            if big:
                n = int(rng.choice([63, 64, 65, 128, 191, 300, cap]))
            o, a = rng.standard_normal((n, od)), rng.uniform(-1, 1, (n, ad))                     # float64 like gymnasium rewards
            r, o2, d = rng.standard_normal(n), rng.standard_normal((n, od)).astype(np.float32), rng.random(n) < 0.3
            if n == 1 and rng.random() < 0.5:
                ours.add(o[0], a[0], float(r[0]), o2[0], bool(d[0]))                             # the reference's scalar call
            else:
                ours.add(o, a, r, o2, d)
            for j in range(n):
                orc.add(o[j], a[j], r[j], o2[j], d[j])
            assert (ours._index, len(ours)) == (orc._index, len(orc)), (cap, n)
            idx = rng.integers(0, len(orc), int(rng.integers(1, 400 if big else 50)))
            got, want = ours.sample(idx), orc.sample(idx)
            for g, w, name in zip(got, want, ("obs", "actions", "rewards", "next_obs", "dones")):
                w = np.asarray(w, dtype=np.float32).reshape(g.shape)
                assert g.dtype == torch.float32 and np.array_equal(g.cpu().numpy(), w), (name, cap, n)
        assert got[2].shape == (idx.size, 1) and got[4].shape == (idx.size, 1)                   # rewards / dones come back [B, 1]


def test_ring_replay_property_emulated(emul):
    _run(torch.device("cpu"), 0)


@pytest.mark.gpu
def test_ring_replay_property_gpu():
    _run(torch.device("cuda"), 1)


def test_host_index_semantics_and_empty_replay_emulated(emul):
    """numpy fancy-indexing semantics for host indices (negative = from the end of the store, out of range raises IndexError like
    ``self.obs[indices]`` in SAC_file/Buffer.py:41-45) and a clear error for learn() on an empty replay."""
    from freerl_b200.Buffer import Buffer
    from freerl_b200.DQN import DQN
    from freerl_b200.SAC import SAC
    dev = torch.device("cpu")
    cap, od, ad = 12, 3, 2
    ours, orc = Buffer(cap, od, ad, dev), ob.RingReplay(cap, od, ad)
    rng = np.random.default_rng(3)
    for _ in range(cap):
        t = (rng.standard_normal(od), rng.uniform(-1, 1, ad), float(rng.standard_normal()), rng.standard_normal(od), bool(rng.random() < 0.5))
        ours.add(*t); orc.add(*t)
    # the reference's array attributes are views of the row store
    for name, want in (("obs", orc.obs), ("actions", orc.actions), ("rewards", orc.rewards), ("next_obs", orc.next_obs), ("dones", orc.dones)):
        assert np.array_equal(getattr(ours, name).cpu().numpy(), np.asarray(want, dtype=np.float32)), name
    idx = np.array([-1, -cap, 0, cap - 1, -3])
    for g, w in zip(ours.sample(idx), orc.sample(idx)):
        assert np.array_equal(g.cpu().numpy(), np.asarray(w, dtype=np.float32).reshape(g.shape))
    for bad in ([cap], [-cap - 1], [0, 5, 10 ** 9]):
        with pytest.raises(IndexError):
            ours.sample(np.array(bad))
    assert all(t.shape[0] == 0 for t in ours.sample(np.array([], dtype=np.int64)))
    with pytest.raises(RuntimeError, match="empty replay"):
        SAC([3, 2], True, 1e-3, 1e-3, 100, dev, trick={}).learn(8, 0.99, 0.01)
    with pytest.raises(RuntimeError, match="empty replay"):
        DQN([4, 2], False, 1e-3, 100, dev).learn(8, 0.99, 0.01)
