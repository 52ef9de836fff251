"""learn() time of MAPPO_discrete at the reference script's default shape (simple_spread N = 3: obs 18, 5 actions, episode_limit 25,
horizon = minibatch = 256 episodes, K_epochs 15): the all-False switch set (tensor-core path, one launch per learn) and the script's
default switches (group mode: episode-wide LayerNorm, scalar huber / ValueClip loss — a pre-pass + an update launch per minibatch).

    python tools/mappo_discrete_bench.py          # on a GPU box
"""
import contextlib, io, sys
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200.MAPPO_discrete import MAPPO, ReplayBuffer
OFF = dict(adv_norm=False, ObsNorm=False, reward_norm=False, reward_scaling=False, orthogonal_init=False, adam_eps=False, lr_decay=False,
           ValueClip=False, huber_loss=False, LayerNorm=False, feature_norm=False)
TRICKS = {"simple": OFF,
          "clip": dict(OFF, adv_norm=True, orthogonal_init=True, adam_eps=True, ValueClip=True),
          # the script's default (`--policy_name MAPPO`, MAPPO_discrete.py:520-527): every switch on except reward_norm / lr_decay
          "full": dict(OFF, adv_norm=True, ObsNorm=True, reward_scaling=True, orthogonal_init=True, adam_eps=True, ValueClip=True,
                       huber_loss=True, LayerNorm=True, feature_norm=True)}

dev = torch.device("cuda")
N, OD, AD, T, B, K = 3, 18, 5, 25, 256, 15
ids = ["agent_%d" % i for i in range(N)]
rng = np.random.default_rng(0)
for name in ("simple", "clip", "full"):
    buf = ReplayBuffer(N=N, obs_dim=OD, state_dim=N * OD, episode_limit=T, batch_size=B, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        pol = MAPPO({k: [OD, AD] for k in ids}, False, 1e-3, 5e-4, B, dev, dict(TRICKS[name]), buf)

    def fill():
        b = buf.buffer
        b["obs_n"][...] = rng.standard_normal(b["obs_n"].shape)
        b["s"][...] = b["obs_n"].reshape(B, T, N * OD)
        b["v_n"][...] = 0.5 * rng.standard_normal(b["v_n"].shape)
        b["a_n"][...] = rng.integers(0, AD, b["a_n"].shape)
        b["a_logprob_n"][...] = np.log(1.0 / AD)
        b["r_n"][...] = rng.standard_normal(b["r_n"].shape)
        b["done_n"][...] = 0.0
        buf.episode_num = B
    ts = []
    for it in range(4):
        fill()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pol.learn(B, 0.95, 0.95, 0.2, K, 0.01, 10.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    m = pol.last_metrics.cpu().numpy()
    print("%-7s learn (upload + GAE + %d updates of %d rows): %.2f ms  (first call %.2f)  finite losses: %s"
          % (name, K, B * T * N, min(ts[1:]), ts[0], bool(np.isfinite(m[:, :2]).all())))
