"""frl_policy_infer takes 16-row tiles for batches of >= 2048 rows (heads of <= 16 outputs): the results must be the ones the 8-row kernel
gives for the same rows (to fp32 rounding: the two GEMM tiles split K differently) — checked by running the same observations once as one big batch and once in chunks below the threshold — and must
agree with a plain torch fp32 forward.  Critic values (RAW), greedy / tanh actions, Categorical draws with injected noise, with and
without the per-row LayerNorm variants."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _run(device):
    from freerl_b200 import _common, _lib
    from freerl_b200.MAPPO import MAPPO
    from oracle.make_golden_marl import MAPPO_TRICK          # the trick dict only
    ids = ["a", "b", "c"]
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    n = 2048 + 37                                             # a ragged last tile for both tile heights
    for ln_on in (True, False):
        trick = dict(MAPPO_TRICK, LayerNorm=ln_on, feature_norm=ln_on)
        with contextlib.redirect_stdout(io.StringIO()):
            pol = MAPPO({k: [18, 5] for k in ids}, False, 1e-3, 1e-3, 8, device, trick)
        net = pol.agents["a"]._net
        obs = rng.standard_normal((n, 18)).astype(np.float32)
        joint = rng.standard_normal((n, 54)).astype(np.float32)
        noise = torch.empty((n, 5)).exponential_(1).to(device)
        cases = [("critic", joint, _lib.INFER_RAW, 1, dict(l0=3, nl=3), None),
                 ("argmax", obs, _lib.INFER_ARGMAX, 1, dict(l0=0, nl=3), None),
                 ("categorical", obs, _lib.INFER_PPO_CAT, 2, dict(l0=0, nl=3), noise)]
        variants = (1, 2, 3) if ln_on else (0,)
        for name, x, mode, cols, kw, nz in cases:
            for ln in variants:
                big = _common.infer(net, x, mode, device, cols, noise=nz, layer_norm=ln, **kw).cpu().numpy()
                parts = [_common.infer(net, x[s:s + 1000], mode, device, cols, noise=None if nz is None else nz[s:s + 1000].contiguous(),
                                       layer_norm=ln, **kw).cpu().numpy() for s in range(0, n, 1000)]
                small = np.concatenate(parts)
                # the 16-row GEMM splits K differently from the 8-row one: the same numbers to fp32 rounding, not bit for bit
                if name == "critic":
                    np.testing.assert_allclose(big, small, rtol=2e-6, atol=2e-6, err_msg="%s ln=%d" % (name, ln))
                else:               # actions: equal unless two logits tie within rounding (none expected in 2085 rows; allow one)
                    assert (big[:, 0] != small[:, 0]).sum() <= 1, (name, ln)
                    same = big[:, 0] == small[:, 0]
                    np.testing.assert_allclose(big[same], small[same], rtol=2e-6, atol=2e-6, err_msg="%s ln=%d" % (name, ln))
        # against torch (critic, input + hidden LayerNorm or none)
        sd = {k: v.detach().cpu() for k, v in pol.agents["a"].critic.state_dict().items()}
        h = torch.from_numpy(joint)
        norm = (lambda t: F.layer_norm(t, t.shape[1:])) if ln_on else (lambda t: t)
        h = norm(h)
        h = norm(F.relu(F.linear(h, sd["l1.weight"], sd["l1.bias"])))
        h = norm(F.relu(F.linear(h, sd["l2.weight"], sd["l2.bias"])))
        want = F.linear(h, sd["l3.weight"], sd["l3.bias"]).numpy()
        got = _common.infer(net, joint, _lib.INFER_RAW, device, 1, l0=3, nl=3, layer_norm=1 if ln_on else 0).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


def test_infer_16_row_tiles_emulated(emul):
    _run(torch.device("cpu"))


@pytest.mark.gpu
def test_infer_16_row_tiles_gpu():
    _run(torch.device("cuda"))
