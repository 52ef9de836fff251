"""Per-stage timestamps of one fused learn launch (debug hook frl_debug_set_timing): ids 100+s = stage s done, 200+s = the grid
barrier after it passed (CTA 0, thread 0; globaltimer).   python tools/stagetime.py [sac|rainbow|ppo|dqn]"""
import contextlib, ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200 import _lib
dev = torch.device('cuda')
rng = np.random.default_rng(0)
algo = sys.argv[1] if len(sys.argv) > 1 else "sac"
with contextlib.redirect_stdout(sys.stderr):
    if algo == "sac":
        from freerl_b200.SAC import SAC
        pol = SAC([17, 6], True, 1e-3, 1e-3, int(1e5), dev, trick={}, mode='fast')
        n = 100000
        pol.add(rng.standard_normal((n, 17), dtype=np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32),
                rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, 17), dtype=np.float32), rng.random(n) < 0.01)
        run = lambda k: pol.learn(256, 0.99, 0.01, n_updates=k)
    elif algo == "dqn":
        from freerl_b200.DQN import DQN
        pol = DQN([4, 2], False, 1e-3, 1e5, dev, mode="fast")
        n = 50000
        pol.add(rng.standard_normal((n, 4)), rng.integers(0, 2, (n, 1)), rng.standard_normal(n), rng.standard_normal((n, 4)), rng.random(n) < 0.01)
        run = lambda k: pol.learn(256, 0.99, 0.01, n_updates=k)
    elif algo == "rainbow":
        from freerl_b200.DQN_with_tricks import DQN
        trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
        pol = DQN([8, 4], False, 1e-3, 1e6, dev, trick=trick, gamma=0.99, batch_size=256, mode="fast")
        for _ in range(40):
            pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)), rng.random(512) < 0.01)
        run = lambda k: [pol.learn(256, 0.99, 0.01) for _ in range(k)]
    else:
        from freerl_b200.PPO import PPO
        T, N = 32, 1024
        pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev, mode="fast")
        for _ in range(T):
            pol.add(rng.standard_normal((N, 8), dtype=np.float32), rng.integers(0, 4, (N, 1)).astype(np.float32), rng.standard_normal(N).astype(np.float32),
                    rng.standard_normal((N, 8), dtype=np.float32), rng.random(N) < 0.01, -np.ones((N, 1), np.float32) * 1.3, rng.random(N) < 0.02)

        def run(k):
            pol.buffer._index, pol.buffer._size, pol.buffer.n_envs = 0, T * N, N
            pol.learn(8192, 0.99, 0.95, 0.2, 1, 0.01)
run(3)
buf = torch.zeros(4000, dtype=torch.int64, device=dev)
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(buf.data_ptr()))
run(1 if algo == "rainbow" else 2)
torch.cuda.synchronize()
_lib.lib().frl_debug_set_timing(ctypes.c_void_p(0))
b = buf.cpu().numpy().reshape(-1, 2)
b = b[b[:, 1] > 0]
t0 = b[0, 1]
prev = t0
for i, (k, t) in enumerate(b[:60]):
    print('%4d id=%3d  t=%8.2f us  dt=%7.2f us' % (i, k, (t - t0) / 1e3, (t - prev) / 1e3))
    prev = t
