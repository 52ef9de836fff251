"""A short run for ncu over the memory-bound helper kernels at streaming sizes (replay add / gather, GAE, adv_norm).

    ncu --set full --clock-control none -k regex:"frl_for_kernel|frl_tile_kernel|frl_simple_kernel" -c 12 -o out python tools/ncu_helpers.py
"""
import contextlib, ctypes, sys
import torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
from freerl_b200.Buffer import Buffer
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(0)
with contextlib.redirect_stdout(sys.stderr):
    L = _lib.lib(); st = _lib.stream_ptr(dev)
    OBS, ACT, cap, n = 17, 6, 4_000_000, 1_000_000
    buf = Buffer(cap, OBS, ACT, dev)
    o = torch.randn((n, OBS), device=dev, generator=g); a_ = torch.rand((n, ACT), device=dev, generator=g)
    r = torch.randn(n, device=dev, generator=g); o2 = torch.randn((n, OBS), device=dev, generator=g); d = torch.zeros(n, device=dev)
    for _ in range(3):
        buf.add_device(o, a_, r, o2, d)
    idx = torch.randint(0, cap, (3, 1 << 20), device=dev, generator=g)
    for k in range(3):
        buf.sample(idx[k])
    T, N = 1024, 16384
    f = lambda: torch.randn((T, N), device=dev, generator=g)
    rw, dn, ad, vs, vn = f(), (f() > 2).float(), (f() > 1.5).float(), f(), f()
    adv, vt = torch.empty((T, N), device=dev), torch.empty((T, N), device=dev)
    for _ in range(3):
        L.frl_gae(_lib.ptr(rw), _lib.ptr(dn), _lib.ptr(ad), _lib.ptr(vs), _lib.ptr(vn), T, N, 0.99, 0.95, _lib.ptr(adv), _lib.ptr(vt), st)
    x = torch.randn(1 << 24, device=dev, generator=g); y = torch.empty_like(x)
    for _ in range(3):
        L.frl_adv_norm(_lib.ptr(x), 1 << 24, ctypes.c_float(1e-8), _lib.ptr(y), st)
    torch.cuda.synchronize()
