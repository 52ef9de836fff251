"""frl_sample_uniform — the on-device stand-in for np.random.choice(N, B, replace=False) (DQN.py:97, SAC.py:213 …) in fast
mode: B distinct in-range indices per update in BOTH regimes (sparse: rejection; dense, N < 2B: partial Fisher-Yates),
reproducible per (seed, counter)."""
import numpy as np
import pytest
import torch


def _check(device):
    from freerl_b200 import _common
    for N, B, U in ((100_000, 256, 8), (600, 256, 4), (511, 256, 4), (300, 256, 4), (256, 256, 4), (257, 256, 3), (5, 5, 2), (7, 3, 2), (1, 1, 1),
                    (100_000, 20_000, 2), (9000, 9000, 2), (70_000, 65_536, 1)):      # > 8192: keyed Feistel bijection
        a = _common.make_indices("fast", N, B, U, device, 99, 3).cpu().numpy()
        assert a.shape == (U, B) and a.min() >= 0 and a.max() < N, (N, B)
        assert all(np.unique(r).size == B for r in a), (N, B)
        assert np.array_equal(a, _common.make_indices("fast", N, B, U, device, 99, 3).cpu().numpy())
        if N > 8:
            assert not np.array_equal(a, _common.make_indices("fast", N, B, U, device, 99, 4).cpu().numpy())
    big = _common.make_indices("fast", 1_000_000, 65_536, 4, device, 7, 1).cpu().numpy()
    hist = np.bincount(big.reshape(-1) // 10_000, minlength=100)
    assert ((hist - hist.mean()) ** 2 / hist.mean()).sum() < 180          # chi-square, 99 dof (the draws are without replacement)
    # dense regime is a uniform shuffle: every position sees every value about equally often
    p = _common.make_indices("fast", 8, 8, 4000, device, 5, 0).cpu().numpy()
    cnt = np.stack([np.bincount(p[:, k], minlength=8) for k in range(8)])
    assert np.abs(cnt - 500).max() < 110, cnt


def test_uniform_sampler_emulated(emul):
    _check(torch.device("cpu"))


@pytest.mark.gpu
def test_uniform_sampler_gpu():
    _check(torch.device("cuda"))
