"""Host side of K5/K6 (Hermitian IDFT -> FIR taps, FIR convolution + fused error sums).

Mirrors the arithmetic of reference src/fir_model/fir_helpers.py:75-137, fir_fitting.py:285-306 and
fir_validation.py:93-136,190 through the C ABI (include/b200id.h).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .runtime import F64, Workspace, empty, host_f64, ptr, require_cuda, side_stream, stream_edges, stream_ptr, to_device


def idft_taps_device(G_d: torch.Tensor, n_taps: int) -> torch.Tensor:
    """First n_taps of Re ifft(Hermitian extension of G_d); G_d complex128 [nd] on the device."""
    nd = G_d.numel()
    M = 2 * nd - 1
    if n_taps > M:
        raise ValueError(f"N={n_taps} exceeds available length M={M}. Reduce N.")
    h = empty(n_taps, G_d.device)
    _lib.call("b200id_idft_hermitian", ptr(torch.view_as_real(G_d.contiguous())), C.c_int(nd), ptr(h),
              C.c_int(n_taps), stream_ptr())
    return h


def idft_taps(G_pos, n_taps: int) -> np.ndarray:
    dev = require_cuda()
    G = torch.from_numpy(np.ascontiguousarray(np.asarray(G_pos, dtype=np.complex128).ravel())).to(dev)
    return idft_taps_device(G, n_taps).cpu().numpy()


def fir_eval_device(u_d: torch.Tensor, y_d: Optional[torch.Tensor], g_d: torch.Tensor, skip: int, y0: float,
                    lead: int = 0, want_pred: bool = False, n: Optional[int] = None):
    """(sums[4] = {sum e^2, sum y^2, sum (y-y0)^2, count}, y_pred or None) on the device.

    ``u_d`` points at output index 0; with ``lead`` > 0 the ``lead`` samples before it must belong to
    the same allocation (pass a view into a longer tensor) -- they are the FIR history of a shard.
    """
    dev = u_d.device
    n = u_d.numel() if n is None else n
    sums = empty(4, dev)
    pred = empty(n, dev) if want_pred else None
    nbytes = _lib.load().b200id_fir_workspace_bytes(n)
    ws = Workspace.get("fir", nbytes, dev)
    _lib.call("b200id_fir_eval", ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int64(lead), ptr(g_d), C.c_int(g_d.numel()),
              C.c_int64(skip), C.c_double(y0), ptr(pred), ptr(sums), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return sums, pred


def metrics_from_sums(sums, skip: int, detrend: bool = False) -> Dict[str, float]:
    """RMSE / FIT% / R2 from the fused sums (fir_validation.py:115-128, fir_fitting.py:296-306); ``detrend`` is
    reported as the caller applied it (fir_validation.py:205)."""
    se, sy, st, cnt = (float(v) for v in sums)
    rmse = math.sqrt(se / cnt)
    norm_y = math.sqrt(sy)
    fit = float(100 * (1.0 - math.sqrt(se) / norm_y)) if norm_y > 0 else 0.0
    r2 = float(1.0 - se / st) if st > 0 else 0.0
    return {"rmse": rmse, "fit_percent": fit, "r2": r2, "transient_skip": skip, "detrend": bool(detrend)}


def fir_predict(u, g, n_out: Optional[int] = None) -> np.ndarray:
    """np.convolve(u, g, 'full')[:n_out] on the device."""
    dev = require_cuda()
    u_d = to_device(np.asarray(u, dtype=np.float64).ravel(), dev)
    g_d = to_device(np.asarray(g, dtype=np.float64).ravel(), dev)
    n_out = u_d.numel() if n_out is None else min(n_out, u_d.numel())
    _, pred = fir_eval_device(u_d, None, g_d, 0, 0.0, want_pred=True, n=n_out)
    return pred.cpu().numpy()


STREAM_MIN_SAMPLES = 1 << 20      # host records at least this long are uploaded piecewise, overlapped with the FIR
STREAM_CHUNKS = 6
_STREAM = os.environ.get("B200ID_FIR_STREAM", "1") != "0"


def fir_eval_streamed(u_h, y_h, g_d: torch.Tensor, skip: int, y0: float, n: int, want_pred: bool = False,
                      n_chunks: int = STREAM_CHUNKS):
    """:func:`fir_eval_device` for a HOST record: u, y go up in pieces on the copy stream and each piece's
    outputs are convolved (history = the pieces already resident) and folded into the error sums as soon as
    it has landed, so only the last, short piece's kernel runs after the copy."""
    dev = g_d.device
    uh, yh = host_f64(u_h), host_f64(y_h)
    taps = g_d.numel()
    edges = stream_edges(n, n_chunks)
    main, side = torch.cuda.current_stream(), side_stream(dev)
    u_d, y_d = empty(n, dev), empty(n, dev)
    pred = empty(n, dev) if want_pred else None
    side.wait_stream(main)
    landed = []
    with torch.cuda.stream(side):
        for lo, hi in zip(edges, edges[1:]):
            u_d[lo:hi].copy_(uh[lo:hi], non_blocking=True)
            y_d[lo:hi].copy_(yh[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            landed.append(ev)
    total = None
    for ev, lo, hi in zip(landed, edges, edges[1:]):
        main.wait_event(ev)
        sums, pk = fir_eval_device(u_d[lo:], y_d[lo:hi], g_d, max(0, skip - lo), y0, lead=min(lo, taps - 1),
                                   want_pred=want_pred, n=hi - lo)
        if want_pred:
            pred[lo:hi].copy_(pk)
        total = sums if total is None else total.add_(sums)
    return total, pred


def fir_validate(u, y, g, detrend: bool = False, want_pred: bool = True,
                 skip: Optional[int] = None) -> Tuple[Dict[str, float], Optional[np.ndarray]]:
    """(metrics, y_pred) of FIR taps g on record (u, y): convolution and error sums on the device."""
    dev = require_cuda()
    from .records import STORE
    rec = None if detrend else STORE.lookup(u, y)          # both vectors of one stored file record?
    if rec is not None and (rec.role_of(u), rec.role_of(y)) != ("u", "y"):
        rec = None
    u = np.asarray(u, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    g = np.asarray(g, dtype=np.float64).ravel()
    if detrend:
        u, y = u - np.mean(u), y - np.mean(y)
    skip = min(10, g.size // 10) if skip is None else int(skip)
    n = min(u.size, y.size)
    if not (0 <= skip < n):
        # the reference slices y[skip:] and would fail on the empty slice further down (fir_fitting.py:288-306)
        raise ValueError(f"record of {n} samples is not longer than the transient skip ({skip})")
    g_d = to_device(g, dev)
    if rec is not None:
        _, u_d, y_d = STORE.on_device(rec, dev)            # normally already in HBM, from the FRF of this file
        sums, pred = fir_eval_device(u_d, y_d, g_d, skip, float(y[skip]), want_pred=want_pred, n=n)
    elif _STREAM and n >= STREAM_MIN_SAMPLES:
        sums, pred = fir_eval_streamed(u[:n], y[:n], g_d, skip, float(y[skip]), n, want_pred=want_pred)
    else:
        sums, pred = fir_eval_device(to_device(u, dev), to_device(y, dev), g_d, skip, float(y[skip]), want_pred=want_pred, n=n)
    m = metrics_from_sums(sums.cpu().numpy(), skip, detrend=detrend)
    return m, (pred.cpu().numpy() if want_pred else None)


def fir_validate_sharded(u_d: torch.Tensor, y_d: torch.Tensor, g_d: torch.Tensor, lo: int, hi: int, y0: float,
                         skip_global: Optional[int] = None) -> Dict[str, float]:
    """Validation metrics of FIR taps on ONE record whose outputs [lo, hi) belong to this rank.

    ``u_d`` is the rank's slice of u INCLUDING min(lo, taps - 1) samples of history in front of output
    ``lo``; ``y_d`` is y[lo:hi].  ``y0`` = y[skip] of the whole record (R2 baseline).  One 4-double
    all-reduce (FIR row)."""
    from . import dist as bdist
    taps = g_d.numel()
    skip = min(10, taps // 10) if skip_global is None else skip_global
    lead = min(lo, taps - 1)
    n_out = hi - lo
    sums, _ = fir_eval_device(u_d[lead:], y_d, g_d, max(0, skip - lo), y0, lead=lead, n=n_out)
    bdist.allreduce_sum_(sums)
    return metrics_from_sums(sums.cpu().numpy(), skip)
