"""A short run for ncu: a few warm-up launches + profiled launches of one fused learn kernel.

    ncu --set full --clock-control none --import-source on -k regex:frl_persistent -s 3 -c 1 -o out python tools/ncu_target.py [sac|rainbow|ppo]
"""
import contextlib, sys
import numpy as np, torch
sys.path.insert(0, '.')
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
        run = lambda: pol.learn(256, 0.99, 0.01, n_updates=8)
    elif algo == "rainbow":
        from freerl_b200.DQN_with_tricks import DQN
        trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
        pol = DQN([8, 4], False, 1e-3, 1e6, dev, trick=trick, gamma=0.99, batch_size=256, mode="fast")
        for _ in range(40):
            pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)), rng.random(512) < 0.01)
        run = lambda: pol.learn(256, 0.99, 0.01)
    else:
        from freerl_b200.PPO import PPO
        T, N = 32, 1024
        pol = PPO([8, 4], False, 1e-3, 1e-3, T * N, dev, mode="fast")
        for _ in range(T):
            pol.add(rng.standard_normal((N, 8), dtype=np.float32), rng.integers(0, 4, (N, 1)).astype(np.float32), rng.standard_normal(N).astype(np.float32),
                    rng.standard_normal((N, 8), dtype=np.float32), rng.random(N) < 0.01, -np.ones((N, 1), np.float32) * 1.3, rng.random(N) < 0.02)

        def run():
            pol.buffer._index, pol.buffer._size, pol.buffer.n_envs = 0, T * N, N
            pol.learn(8192, 0.99, 0.95, 0.2, 1, 0.01)           # 4 minibatch updates of 8192 rows per launch
for _ in range(6):
    run()
torch.cuda.synchronize()
