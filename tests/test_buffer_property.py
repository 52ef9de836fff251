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
