"""Scratch: clock64 trace of the first ops of one fused SAC learn (variants/v_trace.so, -DFRL_TRACE)."""
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
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.lib().frl_debug_set_trace_cta(cta)
print('trace of CTA', cta)
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(buf.data_ptr()))
pol.learn(256, 0.99, 0.01, n_updates=2)
torch.cuda.synchronize()
b = buf.cpu().numpy().reshape(-1, 2)
b = b[b[:, 1] > 0]
b = b[np.argsort(b[:, 1], kind='stable')]
starts = [i for i, (k, t) in enumerate(b) if k == 1000]
i0 = starts[1] if len(starts) > 1 else 0
t0 = b[i0, 1]; prev = t0
for i, (k, t) in enumerate(b[i0:i0 + 420]):
    print('%4d id=%4d  t=%8d clk  dt=%6d' % (i, k, t - t0, t - prev)); prev = t
