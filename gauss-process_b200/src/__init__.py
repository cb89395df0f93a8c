"""Drop-in ``src`` package: the three hot-path packages of the reference, B200-native underneath.

``src.frequency_transform``, ``src.gpr`` and ``src.fir_model`` keep the reference's module paths,
names, signatures, return types and error behaviour and route the
arithmetic through libb200id.so (include/b200id.h).  Everything else of the reference's ``src``
(``src.pipeline``, ``src.run_baseline``, ``src.utils``, ``src.visualization``,
``src.classical_methods`` -- the callers and the out-of-scope subsystems) is NOT rebuilt here: when
a reference checkout is available (``GP_REFERENCE_ROOT``, default /root/reference) its ``src``
directory is appended to this package's ``__path__``, so ``python main.py ...`` and
``python -m src.run_baseline`` run unchanged with this directory first on ``sys.path``: submodule
lookup finds our three packages first and the reference's for the rest.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_root = _os.path.dirname(_here)
if _pkg_root not in _sys.path:          # makes `b200id` importable next to `src`
    _sys.path.insert(0, _pkg_root)

_ref = _os.path.join(_os.environ.get("GP_REFERENCE_ROOT", "/root/reference"), "src")
if _os.path.isdir(_ref) and _os.path.abspath(_ref) != _here and _ref not in __path__:
    __path__.append(_ref)

B200ID_DROP_IN = True
