"""Scratch: a short run for ncu (3 warm-up launches + a few profiled launches of the fused SAC learn)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200.SAC import SAC
dev = torch.device('cuda')
pol = SAC([17, 6], True, 1e-3, 1e-3, int(1e5), dev, trick={}, mode='fast')
rng = np.random.default_rng(0)
n = 100000
pol.add(rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32),
        rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32), rng.random(n) < 0.01)
for _ in range(6):
    pol.learn(256, 0.99, 0.01, n_updates=8)
torch.cuda.synchronize()
print(pol.last_metrics[-1])
