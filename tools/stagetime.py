"""Scratch: per-stage timestamps of one fused SAC learn launch (debug hook frl_debug_set_timing)."""
import sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
from freerl_b200.SAC import SAC
dev = torch.device('cuda')
pol = SAC([17, 6], True, 1e-3, 1e-3, int(1e5), dev, trick={}, mode='fast')
rng = np.random.default_rng(0)
n = 100000
pol.add(rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32),
        rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32), rng.random(n) < 0.01)
for _ in range(3):
    pol.learn(256, 0.99, 0.01, n_updates=4)
buf = torch.zeros(4000, dtype=torch.int64, device=dev)
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(buf.data_ptr()))
pol.learn(256, 0.99, 0.01, n_updates=3)
torch.cuda.synchronize()
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(0))
b = buf.cpu().numpy().reshape(-1, 2)
b = b[b[:, 1] > 0]
t0 = b[0, 1]
prev = t0
for i, (k, t) in enumerate(b):
    print('%4d id=%3d  t=%8.2f us  dt=%7.2f us' % (i, k, (t - t0) / 1e3, (t - prev) / 1e3))
    prev = t
