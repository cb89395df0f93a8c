"""Host side of K1 (FRF projection): NumPy in / NumPy out, CUDA in the middle.

Mirrors the arithmetic of reference src/frequency_transform/frf_estimator.py:194-252 through the
C ABI (include/b200id.h): b200id_frf_moments -> b200id_frf_project -> b200id_frf_finalize ->
b200id_cross_power.  ``*_device`` variants keep everything resident in HBM (used by bench.py's
device-resident timing and by the multi-GPU shards).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .runtime import F64, Workspace, empty, ptr, require_cuda, stream_ptr, to_device


def _ws(nd: int, dev):
    nbytes = _lib.load().b200id_frf_workspace_bytes(int(nd))
    buf = Workspace.get("frf", nbytes, dev)
    return buf, nbytes


def moments_device(u_d: torch.Tensor, y_d: Optional[torch.Tensor], own: Tuple[int, int]) -> torch.Tensor:
    """Device [2] tensor: sum of u and of y over the owned sample range."""
    dev = u_d.device
    sums = empty(2, dev)
    ws, nbytes = _ws(1, dev)
    _lib.call("b200id_frf_moments", ptr(u_d), ptr(y_d), C.c_int64(u_d.numel()), C.c_int64(own[0]), C.c_int64(own[1]),
              ptr(sums), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return sums


def project_device(t_d: torch.Tensor, u_d: torch.Tensor, y_d: Optional[torch.Tensor], w_d: torch.Tensor,
                   own: Optional[Tuple[int, int]] = None, sums: Optional[torch.Tensor] = None,
                   inv_count: float = 0.0) -> torch.Tensor:
    """Device [4*nd] partial sums {S_uc, S_us, S_yc, S_ys} over the owned range (no scaling)."""
    dev = t_d.device
    n, nd = t_d.numel(), w_d.numel()
    own = (0, n) if own is None else own
    out = empty(4 * nd, dev)
    ws, _ = _ws(nd, dev)
    _lib.call("b200id_frf_project", ptr(t_d), ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int64(own[0]), C.c_int64(own[1]),
              ptr(w_d), C.c_int(nd), ptr(sums), C.c_double(inv_count), ptr(out), ptr(ws), C.c_size_t(ws.numel()),
              stream_ptr())
    return out


# max w * |t_i - (t0 + i dt)| for which the first-order fast path is used: the dropped second-order term is
# eps^2 / 2 <= ~3e-12 in the basis values, three orders below the 1e-9 parity budget
UNIFORM_PHASE_TOL = 2e-6
_FORCE_PATH = os.environ.get("B200ID_FRF_PATH", "auto")     # auto | general | uniform


def uniformity_device(t_d: torch.Tensor, t0: float, dt: float) -> float:
    """max_i |t_i - (t0 + i dt)| (one small D2H read)."""
    dev = t_d.device
    out = empty(1, dev)
    ws, _ = _ws(1, dev)
    _lib.call("b200id_frf_uniformity", ptr(t_d), C.c_int64(t_d.numel()), C.c_double(t0), C.c_double(dt), ptr(out),
              ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return float(out.item())


def project_uniform_device(t_d, u_d, y_d, w_d, t0: float, dt: float, own=None, sums=None, inv_count: float = 0.0):
    """Same partial sums as project_device through the uniform-timebase fast path."""
    dev = t_d.device
    n, nd = t_d.numel(), w_d.numel()
    own = (0, n) if own is None else own
    out = empty(4 * nd, dev)
    ws, _ = _ws(nd, dev)
    _lib.call("b200id_frf_project_uniform", ptr(t_d), ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int64(own[0]),
              C.c_int64(own[1]), ptr(w_d), C.c_int(nd), ptr(sums), C.c_double(inv_count), C.c_double(t0),
              C.c_double(dt), ptr(out), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out


def choose_path(t_d: torch.Tensor, span: float, w_max: float, t0: Optional[float] = None):
    """('uniform', t0, dt) when the record's time stamps sit on an arithmetic progression up to FP64
    rounding (then the fast kernel reproduces the reference's per-sample phase rounding to first order),
    else ('general', None, None)."""
    n = t_d.numel()
    if _FORCE_PATH == "general" or n < 3:
        return "general", None, None
    t0 = float(t_d[0].item()) if t0 is None else float(t0)
    dt = span / (n - 1)
    if not (dt > 0.0):
        return "general", None, None
    dmax = uniformity_device(t_d, t0, dt)
    if _FORCE_PATH == "uniform" or dmax * w_max < UNIFORM_PHASE_TOL:
        return "uniform", t0, dt
    return "general", None, None


def finalize_device(partial: torch.Tensor, nd: int, scale: float, want_y: bool = True):
    """(U, Y) device complex128 tensors from the partial sums; scale = 2 / T."""
    dev = partial.device
    U = empty(2 * nd, dev)
    Y = empty(2 * nd, dev) if want_y else None
    _lib.call("b200id_frf_finalize", ptr(partial), C.c_int(nd), C.c_double(scale), ptr(U), ptr(Y), stream_ptr())
    return torch.view_as_complex(U.view(nd, 2)), (torch.view_as_complex(Y.view(nd, 2)) if want_y else None)


def cross_power_device(U: torch.Tensor, Y: torch.Tensor, eps: float = 1e-15) -> torch.Tensor:
    """G [nd] complex128 from stacked records U, Y of shape [R, nd] (or [nd])."""
    U2 = U.reshape(-1, U.shape[-1]).contiguous()
    Y2 = Y.reshape(-1, Y.shape[-1]).contiguous()
    R, nd = U2.shape
    G = empty(2 * nd, U2.device)
    _lib.call("b200id_cross_power", ptr(torch.view_as_real(U2)), ptr(torch.view_as_real(Y2)), C.c_int(R), C.c_int(nd),
              C.c_double(eps), ptr(G), stream_ptr())
    return torch.view_as_complex(G.view(nd, 2))


def record_coefficients_device(t_d, u_d, y_d, w_d, span: float, subtract_mean: bool = True,
                               t0: Optional[float] = None, w_max: Optional[float] = None, path=None):
    """(U, Y) of one resident record.  ``span`` = t[-1] - t[0] (the caller knows it on the host).

    ``path`` = a (kind, t0, dt) tuple from :func:`choose_path` to skip the uniformity probe (callers that
    process the same record repeatedly cache it); None probes."""
    n, nd = t_d.numel(), w_d.numel()
    sums = moments_device(u_d, y_d, (0, n)) if subtract_mean else None
    if path is None:
        w_max = float(w_d.max().item()) if w_max is None else w_max
        path = choose_path(t_d, span, w_max, t0)
    if path[0] == "uniform":
        part = project_uniform_device(t_d, u_d, y_d, w_d, path[1], path[2], (0, n), sums, 1.0 / n)
    else:
        part = project_device(t_d, u_d, y_d, w_d, (0, n), sums, 1.0 / n)
    return finalize_device(part, nd, 2.0 / span, want_y=y_d is not None)


# ----------------------------------------------------------------------------------------- host API
def _prepare(t, x_list: Sequence[np.ndarray], drop_seconds: float):
    """Transient drop + the reference's two ValueErrors (frf_estimator.py:212-223), on the host."""
    t = np.asarray(t, dtype=np.float64)
    xs = [np.asarray(x, dtype=np.float64) for x in x_list]
    if drop_seconds and drop_seconds > 0.0:
        keep = t >= (t[0] + drop_seconds)
        if keep.all():
            pass
        elif not keep.any():
            t, xs = t[:0], [x[:0] for x in xs]
        elif t.size > 1 and bool(np.all(np.diff(t) > 0)):
            first = int(np.argmax(keep))          # monotone time: the mask is a suffix -> a view
            t, xs = t[first:], [x[first:] for x in xs]
        else:
            t, xs = t[keep], [x[keep] for x in xs]
    if t.size < 2:
        raise ValueError("Not enough samples after transient drop.")
    span = float(t[-1] - t[0])
    if span <= 0:
        raise ValueError("Non-positive record length.")
    return t, xs, span


def demod(t, x, w, drop_seconds: float = 0.0, subtract_mean: bool = True) -> np.ndarray:
    """One-signal synchronous demodulation (the reference's synchronous_coefficients_trapz)."""
    dev = require_cuda()
    t, (x,), span = _prepare(t, [x], drop_seconds)
    w = np.asarray(w, dtype=np.float64)
    U, _ = record_coefficients_device(to_device(t, dev), to_device(x, dev), None, to_device(w.ravel(), dev), span,
                                      subtract_mean, t0=float(t[0]), w_max=float(np.max(np.abs(w))))
    return U.cpu().numpy().reshape(w.shape)


def demod_pair(t, u, y, w, drop_seconds: float = 0.0, subtract_mean: bool = True):
    """(U, Y) with t, u, y read once (fused form of two synchronous_coefficients_trapz calls)."""
    dev = require_cuda()
    t, (u, y), span = _prepare(t, [u, y], drop_seconds)
    w = np.asarray(w, dtype=np.float64)
    U, Y = record_coefficients_device(to_device(t, dev), to_device(u, dev), to_device(y, dev),
                                      to_device(w.ravel(), dev), span, subtract_mean, t0=float(t[0]),
                                      w_max=float(np.max(np.abs(w))))
    UY = torch.stack([U, Y]).cpu().numpy()
    return UY[0].reshape(w.shape), UY[1].reshape(w.shape)


def cross_power(U_list: List[np.ndarray], Y_list: List[np.ndarray], eps: float = 1e-15) -> np.ndarray:
    """frf_cross_power_average (frf_estimator.py:239-252) on the device."""
    if not U_list or not Y_list or len(U_list) != len(Y_list):
        raise ValueError("Empty or mismatched lists for cross-power averaging.")
    dev = require_cuda()
    U = torch.from_numpy(np.ascontiguousarray(np.vstack(U_list), dtype=np.complex128)).to(dev)
    Y = torch.from_numpy(np.ascontiguousarray(np.vstack(Y_list), dtype=np.complex128)).to(dev)
    return cross_power_device(U, Y, eps).cpu().numpy()


def estimate_records(records, w, drop_seconds: float = 0.0, subtract_mean: bool = True, eps: float = 1e-15):
    """G over a list of in-memory (t, u, y) records: per-record demodulation, one cross-power average."""
    dev = require_cuda()
    w_d = to_device(np.asarray(w, dtype=np.float64).ravel(), dev)
    Us, Ys = [], []
    for (t, u, y) in records:
        t, (u, y), span = _prepare(t, [u, y], drop_seconds)
        U, Y = record_coefficients_device(to_device(t, dev), to_device(u, dev), to_device(y, dev), w_d, span,
                                          subtract_mean, t0=float(t[0]), w_max=float(np.max(np.abs(w))))
        Us.append(U)
        Ys.append(Y)
    G = cross_power_device(torch.stack(Us), torch.stack(Ys), eps)
    return G.cpu().numpy(), torch.stack(Us).cpu().numpy(), torch.stack(Ys).cpu().numpy()


# ----------------------------------------------------------------------------------------- time-segment shards
def record_coefficients_sharded(t_d, u_d, y_d, w_d, lo: int, hi: int, n_total: int, span: float, w_max: float,
                                subtract_mean: bool = True):
    """(U, Y) of ONE record whose samples are split over ranks.

    This rank holds the resident slice t_d/u_d/y_d = samples [lo - h0, hi + h1) of the record, with
    h0 = 1 if lo > 0 and h1 = 1 if hi < n_total (one-sample halo for the trapezoid weights), and owns
    [lo, hi).  Two tiny all-reduces: (sum u, sum y) for the mean, then the 4*nd partial sums; every
    rank finalises redundantly.  ``span`` = t[-1] - t[0] of the WHOLE record.
    """
    from . import dist as bdist
    h0 = 1 if lo > 0 else 0
    n_local = t_d.numel()
    own = (h0, h0 + (hi - lo))
    assert own[1] <= n_local
    nd = w_d.numel()
    sums = None
    if subtract_mean:
        sums = moments_device(u_d, y_d, own)
        bdist.allreduce_sum_(sums)
    span_local = float((t_d[-1] - t_d[0]).item())
    kind, t0, dt = choose_path(t_d, span_local, w_max)
    if kind == "uniform":
        part = project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, own, sums, 1.0 / n_total)
    else:
        part = project_device(t_d, u_d, y_d, w_d, own, sums, 1.0 / n_total)
    bdist.allreduce_sum_(part)
    return finalize_device(part, nd, 2.0 / span, want_y=y_d is not None)
