"""ctypes binding of libb200id.so (include/b200id.h) -- the only way the Python mirror reaches CUDA.

There is deliberately no CPU fallback: if the shared library is missing or a call fails, the
product raises.  PyTorch is used by callers only to own device buffers and streams; pointers and
the stream handle cross this boundary as plain integers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("B200ID_LIB", PKG_ROOT / "lib" / "libb200id.so"))

P = C.c_void_p
I = C.c_int
L = C.c_int64
D = C.c_double
Z = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/b200id.h declares
SIGNATURES = {
    "b200id_version": (I, []),
    "b200id_last_error": (C.c_char_p, []),
    "b200id_device_info": (I, [C.POINTER(I), C.POINTER(I), C.POINTER(I)]),
    "b200id_frf_workspace_bytes": (Z, [I]),
    "b200id_frf_moments": (I, [P, P, L, L, L, P, P, Z, P]),
    "b200id_frf_project": (I, [P, P, P, L, L, L, P, I, P, D, P, P, Z, P]),
    "b200id_frf_uniformity": (I, [P, L, L, L, D, D, P, P, Z, P]),
    "b200id_frf_project_uniform": (I, [P, P, P, L, L, L, P, I, P, D, D, D, P, P, Z, P]),
    "b200id_frf_timebase_check": (I, [P, L, D, D, P, P, Z, P]),
    "b200id_frf_timebase_fill": (I, [P, L, D, D, P]),
    "b200id_frf_select_recurrence": (I, [I]),
    "b200id_frf_adaptive_force": (I, [I]),
    "b200id_frf_adaptive_diag": (I, [P, Z, I, P, P]),
    "b200id_fir_select_fft": (I, [I]),
    "b200id_frf_project_uniform_mma": (I, [P, P, P, L, L, L, P, I, P, D, D, D, P, P, Z, P]),
    "b200id_cross_power_sums": (I, [P, P, I, I, I, P, P]),
    "b200id_gp_targets": (I, [P, I, D, P, P, P, P, P]),
    "b200id_dft_bins": (I, [P, P, L, L, L, P, I, L, I, P, P, D, P, P, P, Z, P]),
    "b200id_frf_finalize": (I, [P, I, D, P, P, P]),
    "b200id_cross_power": (I, [P, P, I, I, D, P, P]),
    "b200id_gp_gram": (I, [I, P, P, I, P, I, P, P]),
    "b200id_gp_batch_workspace_bytes": (Z, [L, I]),
    "b200id_gp_batch_eval": (I, [I, P, P, L, P, P, I, D, I, P, P, I, D, D, P, P, Z, P]),
    "b200id_gp_batch_eval_pair": (I, [I, P, P, L, P, P, P, I, D, I, P, P, P, I, D, D, D, D, P, P, P, Z, P]),
    "b200id_gp_shared_workspace_bytes": (Z, [I, I, I, I]),
    "b200id_gp_batch_eval_shared": (I, [I, P, L, P, P, I, P, P, P, I, D, I, P, P, P, I, D, D, D, D, P, P, P, Z, P]),
    "b200id_gp_batch_eval_noise": (I, [I, P, P, P, L, P, P, I, I, P, P, I, D, D, P, P, Z, P]),
    "b200id_first_strict_min": (I, [P, L, P, P, P, Z, P]),
    "b200id_first_strict_min_segments": (I, [P, P, I, P, P]),
    "b200id_argmin_pack": (I, [P, I, P, P, P]),
    "b200id_argmin_pick": (I, [P, I, I, P, P, P]),
    "b200id_gp_fit_workspace_bytes": (Z, [I]),
    "b200id_gp_fit": (I, [I, P, P, P, I, D, P, P, P, P, Z, P]),
    "b200id_gp_fit_batch": (I, [I, P, P, P, I, I, D, P, P, P, Z, P]),
    "b200id_gp_fit_pinv": (I, [I, P, P, P, I, D, D, P, P, P, Z, P]),
    "b200id_gp_predict": (I, [I, P, P, I, P, P, P, I, P, P, P]),
    "b200id_chol_blocked_workspace_bytes": (Z, [I]),
    "b200id_chol_blocked": (I, [P, I, I, P, P, Z, P]),
    "b200id_kr_workspace_bytes": (Z, [I, I]),
    "b200id_kr_summaries": (I, [P, P, L, I, P, P, P, P, Z, P]),
    "b200id_kr_nll_batch": (I, [I, P, I, P, P, D, D, I, D, P, P, P, Z, P]),
    "b200id_kr_posterior": (I, [I, P, P, P, I, P, P, P, P, P, Z, P]),
    "b200id_idft_hermitian": (I, [P, I, P, I, P]),
    "b200id_fir_workspace_bytes": (Z, [L]),
    "b200id_fir_eval": (I, [P, P, L, L, P, I, L, D, P, P, P, Z, P]),
    "b200id_fp64_peak_probe": (I, [I, I, P, C.POINTER(D), P]),
}
# host-side scalar math exposed for CPU unit tests of the numerics (not part of b200id.h)
SELFTEST_SIGNATURES = {
    "b200id_selftest_sincos_host": (None, [P, L, P, P]),
    "b200id_selftest_gram_host": (None, [I, P, P, I, P, I, P]),
}

ERROR_NAMES = {-1: "E_ARG", -2: "E_WORKSPACE", -3: "E_CUDA", -4: "E_UNSUPPORTED"}


class B200IDError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200id: {ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None

# kernels launched per successful entry-point call (bench.py reports the sum as gpu_launches)
LAUNCHES_PER_CALL = {
    "b200id_frf_moments": 2, "b200id_frf_project": 2, "b200id_frf_project_uniform": 3, "b200id_frf_uniformity": 2, "b200id_frf_timebase_check": 2, "b200id_frf_timebase_fill": 1, "b200id_frf_finalize": 1, "b200id_cross_power": 1, "b200id_dft_bins": 2,
    "b200id_gp_gram": 1, "b200id_gp_batch_eval": 1, "b200id_gp_batch_eval_noise": 1, "b200id_gp_batch_eval_pair": 1, "b200id_gp_batch_eval_shared": 2, "b200id_first_strict_min": 1, "b200id_gp_fit": 1,
    "b200id_gp_fit_pinv": 1, "b200id_gp_predict": 1, "b200id_idft_hermitian": 1, "b200id_fir_eval": 3,
    "b200id_fp64_peak_probe": 1,
}
launch_count = 0
# optional per-call CUDA-event timing: name -> list of (start_event, end_event); set by bench.py
event_log = None


def load() -> C.CDLL:
    """Load the shared library (once) and type every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python gauss-process_b200/build.py` "
            "(there is no CPU fallback for the b200id hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in {**SIGNATURES, **SELFTEST_SIGNATURES}.items():
        fn = getattr(lib, name)       # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise B200IDError on a non-zero status."""
    global launch_count
    lib = load()
    if event_log is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200IDError(rc, lib.b200id_last_error().decode("utf-8", "replace"))
    if event_log is not None:
        e1.record()
        event_log.setdefault(name, []).append((e0, e1))
    launch_count += LAUNCHES_PER_CALL.get(name, 1)


class timed_region:
    """``with timed_region("host:label"):`` -- CUDA-event timing of a host-side region (collectives, torch glue) into the
    same per-call log as the entry points; free when the log is off."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if event_log is not None:
            import torch
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if event_log is not None and exc[0] is None:
            self.e1.record()
            event_log.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def device_info():
    sm, major, minor = I(0), I(0), I(0)
    call("b200id_device_info", C.byref(sm), C.byref(major), C.byref(minor))
    return sm.value, major.value, minor.value
