"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/freerl_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "freerl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    so = os.path.join(ROOT, "freerl_b200", "libfreerl_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 20, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.frl_abi_version.restype = ctypes.c_int
    from freerl_b200 import _lib
    assert lib.frl_abi_version() == _lib.ABI_VERSION
    assert lib.frl_is_emulation() == 0
    # the ctypes mirrors have the byte size of the compiled structs (also enforced by _lib.lib() at load time)
    lib.frl_struct_size.restype = ctypes.c_int
    mirrors = (_lib.Layer, _lib.Net, _lib.Replay, _lib.DqnArgs, _lib.AcArgs, _lib.InferArgs, _lib.PpoArgs, _lib.NoisyMap, _lib.RainbowArgs, _lib.ExploreArgs,
               _lib.SacdArgs, _lib.ReplicaAvgArgs)
    for which, m in enumerate(mirrors):
        assert lib.frl_struct_size(which) == ctypes.sizeof(m), m.__name__
    assert lib.frl_struct_size(len(mirrors)) == -1


def test_product_path_refuses_cpu_and_missing_library(monkeypatch, tmp_path):
    """No CPU fallback: a CPU device is rejected by the CUDA library, a missing library raises."""
    import torch
    from freerl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.delenv("FREERL_B200_LIB", raising=False)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA devices only"):
            _lib.require_device("cpu")
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("FREERL_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not found"):
        _lib.lib()
    monkeypatch.setattr(_lib, "_lib", None)


def test_entry_points_reject_bad_arguments_without_touching_the_gpu():
    """Error behaviour of the boundary: every function returns an int status, never throws; bad arguments give a negative
    code and a thread-local message from frl_last_error().  Argument validation runs before any CUDA call, so this is checked on the
    REAL CUDA library on a machine without a GPU."""
    so = os.path.join(ROOT, "freerl_b200", "libfreerl_b200.so")
    lib = ctypes.CDLL(so)
    from freerl_b200 import _lib
    _lib._declare(lib)
    N = None
    calls = {
        "frl_replay_add_batch": lambda: lib.frl_replay_add_batch(N, 0, N, N, N, N, N, 4, N),
        "frl_replay_gather": lambda: lib.frl_replay_gather(N, N, 4, N, N, N, N, N, N),
        "frl_gae": lambda: lib.frl_gae(N, N, N, N, N, 8, 8, 0.99, 0.95, N, N, N),
        "frl_adv_norm": lambda: lib.frl_adv_norm(N, 100, ctypes.c_float(1e-8), N, N),
        "frl_dqn_learn": lambda: lib.frl_dqn_learn(N, N),
        "frl_ac_learn": lambda: lib.frl_ac_learn(N, N),
        "frl_ppo_update": lambda: lib.frl_ppo_update(N, N),
        "frl_policy_infer": lambda: lib.frl_policy_infer(N, N),
        "frl_rainbow_learn": lambda: lib.frl_rainbow_learn(N, N),
        "frl_sumtree_update": lambda: lib.frl_sumtree_update(N, 16, N, N, N, 0.0, 0, 0, 4, N, N),
        "frl_sumtree_update_td": lambda: lib.frl_sumtree_update_td(N, 16, N, N, ctypes.c_float(0.01), ctypes.c_float(0.5), 4, N, N),
        "frl_vecnorm": lambda: lib.frl_vecnorm(N, 0, N, 0, 4, 3, 1, N, N, N),
        "frl_reward_scaling": lambda: lib.frl_reward_scaling(N, 0, N, N, 1, 0.99, 4, N, N, N),
        "frl_explore": lambda: lib.frl_explore(N, N),
        "frl_masked_reset": lambda: lib.frl_masked_reset(N, N, 4, 1, 0.0, N),
        "frl_epsilon_greedy": lambda: lib.frl_epsilon_greedy(N, 4, 2, 0.1, N, N, 0, 0, N, N),
        "frl_dis_to_con": lambda: lib.frl_dis_to_con(N, 4, 11, 1, 0, N, N, N, N, N),
        "frl_replica_average": lambda: lib.frl_replica_average(N, N),
    }
    for name, call in calls.items():
        rc = call()
        assert rc < 0, name
        msg = lib.frl_last_error().decode()
        assert name.replace("frl_", "").split("_")[0] in msg or name in msg, (name, msg)
    # a struct with inconsistent fields is refused too (PPO tanh + layer_norm; unknown explore kind)
    g = _lib.PpoArgs()                       # group mode (MAPPO_discrete.py) is Categorical-only; the scalar value loss needs group mode
    g.mb = g.n_updates = g.n_adv = 1
    for k in ("indices", "mb_rows", "gpart", "sumsq", "segcnt", "stats", "out"):
        setattr(g, k, 8)
    g.value_loss = 3
    assert lib.frl_ppo_update(ctypes.byref(g), None) < 0 and "frl_ppo_update" in lib.frl_last_error().decode()
    r = _lib.ReplicaAvgArgs()                # a world of one has nothing to average
    r.n_tensors, r.dp.world = 1, 1
    assert lib.frl_replica_average(ctypes.byref(r), None) < 0 and "frl_replica_average" in lib.frl_last_error().decode()
    a = _lib.ExploreArgs()
    a.kind, a.N, a.A = 7, 4, 2
    assert lib.frl_explore(ctypes.byref(a), None) < 0 and "frl_explore" in lib.frl_last_error().decode()
