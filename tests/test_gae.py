"""frl_gae / frl_adv_norm through the C ABI vs the oracle's float64 reverse scan (PPO_file/PPO.py:222-233) on vectorised
[T, N] rollouts: the coalesced column-tile kernel (N >= 32, with and without the shared-memory stash), the
warp-per-column kernel (N < 32), ragged T / N, and segment resets at adv_done."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import algos


def _gae_case(_lib, device, T, N, seed):
    rng = np.random.default_rng(seed)
    rew = rng.standard_normal((T, N)).astype(np.float32)
    vs = rng.standard_normal((T, N)).astype(np.float32)
    vn = rng.standard_normal((T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.03).astype(np.float32)
    adone = np.maximum(done, (rng.random((T, N)) < 0.02).astype(np.float32))
    gamma, lmbda = 0.99, 0.95
    td = (torch.from_numpy(rew) + gamma * (1.0 - torch.from_numpy(done)) * torch.from_numpy(vn) - torch.from_numpy(vs)).numpy()
    want = np.stack([algos.gae_reference(td[:, j].astype(np.float64), adone[:, j].astype(np.float64), gamma, lmbda) for j in range(N)], axis=1)
    d = lambda x: torch.from_numpy(x).to(device)
    r_, dn_, ad_, vs_, vn_ = d(rew), d(done), d(adone), d(vs), d(vn)
    adv, vt = torch.empty((T, N), device=device), torch.empty((T, N), device=device)
    _lib.check(_lib.lib().frl_gae(_lib.ptr(r_), _lib.ptr(dn_), _lib.ptr(ad_), _lib.ptr(vs_), _lib.ptr(vn_), T, N, gamma, lmbda,
                                  _lib.ptr(adv), _lib.ptr(vt), _lib.stream_ptr(device)), "frl_gae")
    np.testing.assert_allclose(adv.cpu().numpy(), want.astype(np.float32), rtol=2e-6, atol=2e-6, err_msg="T=%d N=%d" % (T, N))
    np.testing.assert_allclose(vt.cpu().numpy(), want.astype(np.float32) + vs, rtol=2e-6, atol=2e-6)


def _run(_lib, device):
    for T, N, seed in ((128, 64, 0), (50, 33, 1), (7, 32, 2), (600, 40, 3), (1, 70, 4), (257, 3, 5), (40, 1, 6), (1030, 64, 7), (129, 32, 8)):
        _gae_case(_lib, device, T, N, seed)
    # advantage normalisation: ragged / unaligned sizes vs torch (MAPPO.py:385-386: (adv - adv.mean()) / (adv.std() + 1e-8))
    rng = np.random.default_rng(9)
    for n in (3, 768, 1025, 70001):
        x = torch.from_numpy((rng.standard_normal(n + 1) * 3 + 1).astype(np.float32))
        for off in (0, 1):
            xs = x[off:off + n]
            xd = x.to(device)[off:off + n]
            out = torch.empty(n + 1, device=device)[off:off + n]
            _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(xd), n, ctypes.c_float(1e-8), _lib.ptr(out), _lib.stream_ptr(device)), "frl_adv_norm")
            want = (xs - xs.mean()) / (xs.std() + 1e-8)
            np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6, err_msg="n=%d off=%d" % (n, off))


def test_gae_adv_norm_emulated(emul):
    _run(emul, torch.device("cpu"))


@pytest.mark.gpu
def test_gae_adv_norm_gpu():
    from freerl_b200 import _lib
    _run(_lib, torch.device("cuda"))


def test_gae_round_boundaries_emulated(emul):
    """The column-tile kernel walks the horizon in rounds of 8 chunks x 16 steps aligned to the END of the rollout: horizons around the
    round size (128) and its multiples, chunk-length boundaries (T <= 128: ceil(T / 8) steps per chunk) and ragged column tiles."""
    for seed, (T, N) in enumerate(((127, 32), (128, 33), (129, 64), (255, 40), (256, 32), (257, 65), (16, 32), (17, 47), (9, 32), (8, 96),
                                   (383, 32), (640, 34))):
        _gae_case(emul, torch.device("cpu"), T, N, 100 + seed)


@pytest.mark.gpu
def test_gae_stream_round_boundaries_gpu():
    """The streamed kernel (gae_stream.cuh: N % 4 == 0, rounds of 8 warps x 4 steps aligned to the END of the rollout): horizons around the
    round size (32) and its multiples, a front round shorter than one warp's chunk, ragged column tiles, and the tile kernel for N % 4 != 0."""
    import os
    import subprocess
    import sys
    if os.environ.get("FREERL_B200_GAE_STREAM") is None:            # the library reads the switch once: run the cases in a child with it set
        env = dict(os.environ, FREERL_B200_GAE_STREAM="1")
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", __file__ + "::test_gae_stream_round_boundaries_gpu"],
                           env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return
    from freerl_b200 import _lib
    for seed, (T, N) in enumerate(((31, 32), (32, 36), (33, 64), (63, 40), (64, 32), (65, 68), (3, 32), (4, 44), (5, 32), (1, 96),
                                   (95, 32), (640, 36), (1024, 256), (127, 33), (130, 47), (40, 4800))):
        _gae_case(_lib, torch.device("cuda"), T, N, 200 + seed)


@pytest.mark.gpu
def test_adv_norm_resident_gpu():
    """The single-launch kernel that keeps the data in shared memory across the grid-wide statistics (adv_norm_resident.cuh: n % 4 == 0,
    aligned, n <= 2 M) against torch in float64, from one quad to the largest size it takes, with a mean far from zero; and the
    same input through the two-launch kernel (a 4-byte offset view) gives the same statistics."""
    from freerl_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(5)
    for n in (4, 8, 1024, 1028, 393216, 1 << 21, 4_000_000):
        x = torch.randn(n + 4, device=dev, generator=g) * 3 + 7
        out = torch.empty(n + 4, device=dev)
        for off in (0, 1):
            xs, os_ = x[off:off + n], out[off:off + n]
            _lib.check(_lib.lib().frl_adv_norm(_lib.ptr(xs), n, ctypes.c_float(1e-8), _lib.ptr(os_), _lib.stream_ptr(dev)), "frl_adv_norm")
            x64 = xs.double()
            want = ((x64 - x64.mean()) / (x64.std() + 1e-8)).float()
            err = (os_ - want).abs().max().item()
            assert err < 5e-6, (n, off, err)
