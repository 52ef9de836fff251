"""Scratch timing of the fused SAC learn (not the official bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200.SAC import SAC
dev = torch.device('cuda')
pol = SAC([17, 6], True, 1e-3, 1e-3, int(1e6), dev, trick={}, mode='fast')
rng = np.random.default_rng(0)
n = 200000
for _ in range(5):
    pol.add(rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32),
            rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32), rng.random(n) < 0.01)
print('buffer', len(pol.buffer))
for U in (1, 16, 256):
    for _ in range(3):
        pol.learn(256, 0.99, 0.01, n_updates=U)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        pol.learn(256, 0.99, 0.01, n_updates=U)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('U=%d: %.3f ms per launch, %.2f us per learn, %.0f learns/s' % (U, ms, ms * 1e3 / U, U / ms * 1e3))
print(pol.last_metrics[-1])
