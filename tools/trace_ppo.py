"""clock64 trace of one tensor-core PPO update (variants/v_trace.so built with -DFRL_TRACE; FREERL_B200_LIB points at it):
ids 2xxx from thread 32 (an epilogue thread) and thread 0 (the MMA issuer) of CTA 0.   python tools/trace_ppo.py"""
import sys, ctypes
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
from freerl_b200.PPO import PPO
dev = torch.device('cuda')
rng = np.random.default_rng(0)
T, N = 32, 1024
pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev, mode="fast")
for _ in range(T):
    pol.add(rng.standard_normal((N, 8), dtype=np.float32), rng.integers(0, 4, (N, 1)).astype(np.float32), rng.standard_normal(N).astype(np.float32),
            rng.standard_normal((N, 8), dtype=np.float32), rng.random(N) < 0.01, -np.ones((N, 1), np.float32) * 1.3, rng.random(N) < 0.02)


def run():
    pol.buffer._index, pol.buffer._size, pol.buffer.n_envs = 0, T * N, N
    pol.learn(8192, 0.99, 0.95, 0.2, 1, 0.01)


run(); run()
buf = torch.zeros(4000, dtype=torch.int64, device=dev)
_lib.lib().frl_debug_set_trace_cta(0)
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
b = buf.cpu().numpy().reshape(-1, 2)
b = b[b[:, 1] > 0]
b = b[np.argsort(b[:, 1], kind='stable')]
sel = b[(b[:, 0] >= 2000) & (b[:, 0] < 2300)]
t0 = sel[0, 1]; prev = t0
for i, (k, t) in enumerate(sel[:150]):
    print('%4d id=%4d  t=%8d clk  dt=%6d' % (i, k, t - t0, t - prev)); prev = t
