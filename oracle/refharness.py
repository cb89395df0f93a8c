"""Import the reference (read-only at /root/reference) behind a matplotlib stub.

Test infrastructure only, and only usable in the build container: the GPU box has no
/root/reference.  Used by oracle/make_golden.py to produce tests/golden/*.npz and by
tests marked ``needs_reference`` (auto-skipped when the tree is absent).

The reference imports matplotlib at module top level on the hot path
(frf_estimator.py:32, fir_fitting.py:26, gp_pipeline.py:16-18) and matplotlib is not
installed here, so a permissive dummy is injected into sys.modules first.  Because the
product reuses the package name ``src``, call :func:`load` only in a process that has NOT
imported the product's ``src`` (make_golden.py and the reference tests run it in a
subprocess / before anything else).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("GP_REFERENCE_ROOT", "/root/reference")


class _Dummy:
    """Object whose every attribute / call / item is another dummy and which unpacks as a pair."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __call__(self, *a, **k):
        return _Dummy()

    def __getitem__(self, key):
        return _Dummy()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter((_Dummy(), _Dummy()))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _DummyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()


def install_matplotlib_stub() -> None:
    try:
        import matplotlib  # noqa: F401
        return
    except ModuleNotFoundError:
        pass
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
                 "matplotlib.patches", "matplotlib.ticker", "matplotlib.gridspec"):
        sys.modules[name] = _DummyModule(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "gpr"))


def load():
    """Put the reference first on sys.path and return its ``src`` package."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "src" in sys.modules and not getattr(sys.modules["src"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("another 'src' package is already imported in this process")
    install_matplotlib_stub()
    import numpy as np
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src  # noqa: F401
    return sys.modules["src"]
