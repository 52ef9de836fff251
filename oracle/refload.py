"""Loader for the UNMODIFIED reference classes (only usable where /root/reference exists).

TEST / FIXTURE-GENERATION INFRASTRUCTURE ONLY.  Nothing under ``freerl_b200/`` may import this.
The reference hot-path files import ``gymnasium`` / ``pettingzoo`` at module top but only use them
inside ``get_env``/``__main__``, so empty stub modules are enough to import the
classes and drive ``add/select_action/learn`` on synthetic data.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("FREERL_REFERENCE", "/root/reference")

_SHARED_NAMES = ("Buffer", "Noisy_net", "normalization", "c_adamw", "util")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "SAC_file"))


def _stub_envs():
    for name in ("gymnasium", "pettingzoo", "pettingzoo.mpe"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:  # tensorboard missing: stub SummaryWriter (only used in __main__)
        m = types.ModuleType("torch.utils.tensorboard")
        m.SummaryWriter = object
        sys.modules["torch.utils.tensorboard"] = m


def load(alg_dir: str, module: str):
    """Import ``/root/reference/<alg_dir>/<module>.py`` under a unique name.

    Same-named helper modules (``Buffer`` exists 11 times with different contents) are purged from
    ``sys.modules`` before every load so each algorithm sees its own directory's copy.
    """
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _stub_envs()
    d = os.path.join(REF_ROOT, alg_dir)
    for n in _SHARED_NAMES:
        sys.modules.pop(n, None)
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("ref_%s_%s" % (alg_dir, module), os.path.join(d, module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
    return mod
