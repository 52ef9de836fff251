import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without CUDA skips the gpu-marked tests instead of failing in the CUDA runtime"""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


EMUL_SO = os.path.join(ROOT, "tests", "emul", "libfreerl_emul.so")
CSRC = os.path.join(ROOT, "freerl_b200", "csrc")


def _build_emul():
    """Host emulation of the kernels (g++ -DFRL_EMUL): TEST-ONLY, checks indexing / arithmetic of the CUDA
    sources without a GPU.  The product path refuses it (frl_is_emulation() == 1 unless tests opt in)."""
    import glob
    import subprocess
    srcs = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(ROOT, "include", "freerl_b200.h")]
    if os.path.exists(EMUL_SO) and all(os.path.getmtime(EMUL_SO) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
    subprocess.check_call(["g++", "-x", "c++", "-DFRL_EMUL", "-O2", "-g", "-std=c++17", "-ffp-contract=fast", "-mfma",
                           "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", EMUL_SO, os.path.join(CSRC, "capi.cu")])


@pytest.fixture()
def emul(monkeypatch):
    """Route freerl_b200 to the host-emulation library for this test (CPU tensors allowed)."""
    _build_emul()
    from freerl_b200 import _lib
    monkeypatch.setenv("FREERL_B200_LIB", EMUL_SO)
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_sm_count", None)
    yield _lib
    _lib._lib = None
    _lib._sm_count = None


@pytest.fixture()
def dev(request):
    """Device for parity tests: cuda under -m gpu, cpu (emulation) otherwise."""
    import torch
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


@pytest.fixture(autouse=True)
def _pin_fast_seed_counter():
    """The seed of the on-device generators (fast mode) counts the policies built in the process (``_common.default_seed``); reset the
    counter before every test so that a test's draws do not depend on which tests ran before it."""
    try:
        from freerl_b200 import _common
        _common._fast_seed_counter[0] = 0
    except Exception:
        pass
    yield
