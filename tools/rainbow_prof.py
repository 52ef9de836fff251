"""Where one Rainbow learn() (C4: B 256, PER capacity 1e6) spends its time: wall clock per sub-step with a device sync after each."""
import contextlib, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from freerl_b200.DQN_with_tricks import DQN
dev = torch.device('cuda')
trick = {"Double": True, "Dueling": True, "PER": True, "Noisy": True, "N_Step": True, "Categorical": True}
with contextlib.redirect_stdout(sys.stderr):
    pol = DQN([8, 4], False, 1e-3, 1e6, dev, trick=trick, gamma=0.99, batch_size=256, mode="fast")
rng = np.random.default_rng(0)
for _ in range(60):
    pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)), rng.random(512) < 0.01)
for _ in range(5):
    pol.learn(256, 0.99, 0.01)
torch.cuda.synchronize()

def wall(fn, reps=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6, r

us, _ = wall(lambda: pol.learn(256, 0.99, 0.01)); print("learn() total             %8.1f us" % us)
us, s = wall(lambda: pol.buffer.sample_device(256)); print("  PER sample_device        %8.1f us" % us)
idx, w, pri = s
us, nz = wall(lambda: [pol._draw_forward_noise() for _ in range(3)]); print("  draw noise (3 forwards)  %8.1f us" % us)
us, _ = wall(lambda: pol._pack_eps(nz)); print("  pack eps + H2D           %8.1f us" % us)
us, a = wall(lambda: pol._args()); print("  build args struct        %8.1f us" % us)
err = torch.randn(256, device=dev)
us, _ = wall(lambda: pol.buffer.update_priorities(idx, err)); print("  update_priorities        %8.1f us" % us)
import ctypes
from freerl_b200 import _lib
a = pol._args(); a.indices, a.B = idx.data_ptr(), 256; a.is_weight = w.data_ptr(); a.gamma, a.tau = 0.97, 0.01
e = torch.empty(256, device=dev); a.error_out = e.data_ptr()
us, _ = wall(lambda: _lib.check(_lib.lib().frl_rainbow_learn(ctypes.byref(a), _lib.stream_ptr(dev)), "x")); print("  frl_rainbow_learn        %8.1f us" % us)
us, _ = wall(lambda: pol.add(rng.standard_normal((512, 8)), rng.integers(0, 4, (512, 1)), rng.standard_normal(512), rng.standard_normal((512, 8)), rng.random(512) < 0.01), 20)
print("add(512 envs)             %8.1f us" % us)
us, _ = wall(lambda: pol.select_action(rng.standard_normal((512, 8)).astype(np.float32)), 20); print("select_action(512 envs)   %8.1f us" % us)
# ---- select_action internals ----
x = rng.standard_normal((512, 8)).astype(np.float32)
us, xd = wall(lambda: torch.from_numpy(x).to(dev), 20); print("  act: H2D obs              %8.1f us" % us)
out = torch.empty(512, device=dev)
a = pol._args()
us, _ = wall(lambda: _lib.check(_lib.lib().frl_rainbow_act(ctypes.byref(a), _lib.ptr(xd), 512, _lib.ptr(out), _lib.stream_ptr(dev)), "x"), 20); print("  act: frl_rainbow_act      %8.1f us" % us)
us, _ = wall(lambda: out.to(torch.int64).cpu().numpy(), 20); print("  act: D2H                  %8.1f us" % us)
us, _ = wall(lambda: pol._device_eps(), 20); print("  device eps (fast mode)    %8.1f us" % us)
# ---- select_action, step by step with a sync after each step ----
import freerl_b200._common as _c
def step_by_step():
    t = []
    def lap(): torch.cuda.synchronize(); t.append(time.perf_counter())
    lap()
    xx, single = _c.as_obs_batch(x, 8); lap()
    xd2 = torch.from_numpy(np.ascontiguousarray(xx, dtype=np.float32)).to(dev); lap()
    pol.agent._last_eps["online"] = ("f", pol._device_eps()[0]); lap()
    o2 = torch.empty(512, dtype=torch.float32, device=dev); aa = pol._args(); lap()
    _lib.check(_lib.lib().frl_rainbow_act(ctypes.byref(aa), _lib.ptr(xd2), 512, _lib.ptr(o2), _lib.stream_ptr(dev)), "x"); lap()
    act = o2.to(torch.int64).cpu().numpy(); lap()
    return [(b - a) * 1e6 for a, b in zip(t, t[1:])]
for _ in range(3): r = step_by_step()
print("  act steps (us): as_obs %.0f | H2D %.0f | device eps %.0f | args %.0f | kernels %.0f | D2H %.0f" % tuple(r))
us, _ = wall(lambda: pol.select_action(x), 50); print("select_action(512) again   %8.1f us" % us)
