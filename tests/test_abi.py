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
    mirrors = (_lib.Layer, _lib.Net, _lib.Replay, _lib.DqnArgs, _lib.AcArgs, _lib.InferArgs, _lib.PpoArgs, _lib.NoisyMap, _lib.RainbowArgs, _lib.ExploreArgs)
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
