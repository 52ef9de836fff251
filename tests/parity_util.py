"""Shared helpers for CUDA-vs-oracle parity tests (run on the GPU under -m gpu, and on the host-emulation
build of the same kernels under -m "not gpu")."""
from collections import OrderedDict

import numpy as np
import torch

TOL = dict(rtol=1e-5, atol=2e-6)


def net_from_golden(g, prefix):
    out = OrderedDict()
    for k in g.files:
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(g[k].copy())
    assert out, prefix
    return out


def golden_batch(g, it):
    return tuple(torch.from_numpy(g["batch/%d/%s" % (it, k)]) for k in ("obs", "act", "rew", "nobs", "done"))


def load_into(module, sd):
    module.load_state_dict({k: v.clone() for k, v in sd.items()})


# Adam divides by sqrt(v) + eps: a parameter whose gradient is of the order of eps = 1e-8 (cancelling contributions) gets
# an update that depends on the LAST BITS of that gradient, so a different — equally valid — fp32 summation order moves it
# by up to a percent of one lr step.  Such elements are rare (observed: 1 in 2944 for the TD3 actor on B200); everything
# else must sit inside the 1e-5 band.  An outlier may not exceed 3 % of an lr = 1e-3 step and there may be at most 0.1 %.
OUTLIER_ABS, OUTLIER_FRAC = 3e-5, 1e-3


def assert_module_close(module, sd, what, tol=TOL):
    got = module.state_dict()
    for k, v in sd.items():
        a, b = got[k].detach().cpu().numpy(), v.detach().cpu().numpy()
        bad = np.abs(a - b) > tol["atol"] + tol["rtol"] * np.abs(b)
        if bad.any() and bad.mean() <= OUTLIER_FRAC and np.abs(a - b).max() <= OUTLIER_ABS:
            continue
        np.testing.assert_allclose(a, b, err_msg="%s/%s" % (what, k), **tol)


def fill_buffer_from_batches(buf, g, n_iter):
    """Store the golden batches as consecutive rows; returns index arrays selecting each batch."""
    idxs, start = [], 0
    for it in range(n_iter):
        obs, act, rew, nobs, done = golden_batch(g, it)
        buf.add(obs.numpy(), act.numpy(), rew.numpy().reshape(-1), nobs.numpy(), done.numpy().reshape(-1))
        B = obs.shape[0]
        idxs.append(np.arange(start, start + B))
        start += B
    return idxs


def push_block_state(net, module, params, moments_m=None, moments_v=None):
    """Teacher forcing: copy an oracle's tensors INTO the device blocks of ``net`` — ``params`` (OrderedDict, the module's state-dict
    names) into ``net.p`` and, when given, the optimiser moments (lists in the same order) into ``net.m`` / ``net.v``.  The module's
    parameters are views of ``net.p``; the same (offset, shape, stride) addresses the moment blocks."""
    sd = module.state_dict()
    with torch.no_grad():
        for i, (k, src) in enumerate(params.items()):
            view = sd[k]
            off = (view.data_ptr() - net.p.data_ptr()) // 4
            for buf, t in ((net.p, src), (net.m if moments_m is not None else None, moments_m[i] if moments_m is not None else None),
                           (net.v if moments_v is not None else None, moments_v[i] if moments_v is not None else None)):
                if buf is None:
                    continue
                torch.as_strided(buf, tuple(view.shape), tuple(view.stride()), off).copy_(t.detach().reshape(view.shape).to(buf.device))
    net.sync_mirror()
