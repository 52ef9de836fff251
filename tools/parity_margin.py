"""How far the CUDA kernels actually sit from the oracle / the reference fixtures, per assertion site: runs the -m gpu parity tests
with ``np.testing.assert_allclose`` and the tests' ``_rel`` helpers wrapped to record the largest error each call site saw, next to
the tolerance it was held to.  Evidence for the tolerances: ``profiles/r4_parity_margin.json``.

    python tools/parity_margin.py [pytest -k expression]      # on a GPU box; writes gpurun_out/parity_margin.json
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REC = {}


def _site(depth=2):
    f = sys._getframe(depth)
    while f and not os.path.basename(f.f_code.co_filename).startswith("test_"):
        f = f.f_back
    return "%s:%d" % (os.path.basename(f.f_code.co_filename), f.f_lineno) if f else "?"


def _note(site, err, tol, kind):
    r = REC.setdefault(site, {"kind": kind, "max_err": 0.0, "tol": tol, "calls": 0})
    r["max_err"] = max(r["max_err"], float(err))
    r["calls"] += 1


_orig_allclose = np.testing.assert_allclose


def _allclose(actual, desired, rtol=1e-7, atol=0, *a, **k):
    x, y = np.asarray(actual, dtype=np.float64), np.asarray(desired, dtype=np.float64)
    if x.shape == y.shape or x.size == 1 or y.size == 1:
        with np.errstate(all="ignore"):
            # error in units of the allowed band atol + rtol |desired| (1.0 = at the limit)
            frac = np.abs(x - y) / (atol + rtol * np.abs(y) + 1e-300)
        if frac.size:
            _note(_site(), np.nanmax(frac), {"rtol": rtol, "atol": atol}, "fraction of the allclose band used")
    return _orig_allclose(actual, desired, rtol, atol, *a, **k)


class Plugin:
    def pytest_collection_finish(self, session):
        for name, mod in list(sys.modules.items()):
            if name.startswith("test_") and hasattr(mod, "_rel"):
                orig = mod._rel

                def rel(a, b, _o=orig):
                    v = _o(a, b)
                    _note(_site(), v, None, "relative error (threshold in the assert)")
                    return v
                mod._rel = rel


if __name__ == "__main__":
    np.testing.assert_allclose = _allclose
    sel = sys.argv[1] if len(sys.argv) > 1 else "parity or select_action or full_size"
    rc = pytest.main([os.path.join(ROOT, "tests"), "-q", "-m", os.environ.get("PARITY_MARGIN_MARK", "gpu"), "-k", sel, "-p", "no:cacheprovider"], plugins=[Plugin()])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_margin.json"), "w") as f:
        json.dump({"pytest_rc": int(rc), "sites": dict(sorted(REC.items()))}, f, indent=1)
    for k, v in sorted(REC.items()):
        print("%-40s %-45s max %.3g  %s" % (k, v["kind"], v["max_err"], v["tol"] or ""))
