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
from .runtime import F64, Workspace, empty, host_f64, ptr, require_cuda, side_stream, stream_edges, stream_ptr, to_device


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


# max w * |t_i - (t0 + i dt)| for which the first-order fast path is used.  Two limits: the dropped second-order term
# eps^2 / 2, and -- the binding one -- the FP32 recurrence that carries the first-order term: its ~1e-4 relative error
# times eps, divided by how weakly a bin is excited (|Y_k| / ||a|| reaches 1e-3 above the plant's corner), has to stay
# below the 1e-9 parity budget.  Measured (tests/test_frf_gpu.py, jittered time stamps): w * delta = 6e-7 gives 4e-9 at
# the weakest output bins, 1.5e-7 (a 1e9-sample record's own rounding) 5e-10.
UNIFORM_PHASE_TOL = 2.5e-7
_FORCE_PATH = os.environ.get("B200ID_FRF_PATH", "auto")     # auto | general | uniform


PHASE_LIMIT = 2.0 ** 36           # |w t| up to which the device sincos (two-constant Cody-Waite) is valid


def check_phase_range(w_max: float, t_first: float, t_last: float) -> None:
    """The FRF kernels reduce the phase fl(w t) with a two-constant Cody-Waite scheme that is exact up to 2^36 rad
    (~7e10: 1e9 samples at 500 Hz and 200 Hz bins stay below 3e9).  Epoch-style time stamps would silently lose the
    phase; there is no CPU fallback on this path, so they are refused with the remedy."""
    tmax = max(abs(float(t_first)), abs(float(t_last)))
    if float(w_max) * tmax >= PHASE_LIMIT:
        raise ValueError(f"FRF phase w*t reaches {float(w_max) * tmax:.3e} rad (limit 2^36 = {PHASE_LIMIT:.3e}): "
                         "shift the time origin (t - t[0]) before estimating the FRF")


def uniformity_device_async(t_d: torch.Tensor, t0: float, dt: float, own=None) -> torch.Tensor:
    """max |t_i - (t0 + i dt)| over i in ``own`` (default: all) as a 1-element device tensor."""
    dev = t_d.device
    own = (0, t_d.numel()) if own is None else own
    out = empty(1, dev)
    ws, _ = _ws(1, dev)
    _lib.call("b200id_frf_uniformity", ptr(t_d), C.c_int64(t_d.numel()), C.c_int64(own[0]), C.c_int64(own[1]),
              C.c_double(t0), C.c_double(dt), ptr(out), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out


def uniformity_device(t_d: torch.Tensor, t0: float, dt: float, own=None) -> float:
    """max_i |t_i - (t0 + i dt)| (one small D2H read)."""
    return float(uniformity_device_async(t_d, t0, dt, own).item())


def timebase_check_async(t_d: torch.Tensor, t0: float, dt: float) -> torch.Tensor:
    """Device [2] tensor whose [0] is the number of samples with t_i != fl(fl(i dt) + t0) (0 = the record's time
    vector is exactly regenerable from (t0, dt), see b200id_frf_timebase_fill)."""
    dev = t_d.device
    out = empty(2, dev)
    ws, _ = _ws(1, dev)
    _lib.call("b200id_frf_timebase_check", ptr(t_d), C.c_int64(t_d.numel()), C.c_double(t0), C.c_double(dt), ptr(out),
              ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out


def timebase_fill_device(n: int, t0: float, dt: float, dev) -> torch.Tensor:
    """t_i = fl(fl(i dt) + t0), i < n, built on the device (bit-identical to ``np.arange(n) * dt + t0``)."""
    t_d = empty(n, dev)
    _lib.call("b200id_frf_timebase_fill", ptr(t_d), C.c_int64(n), C.c_double(t0), C.c_double(dt), stream_ptr())
    return t_d


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


def adaptive_diag(nd: int, dev=None) -> Tuple[int, float]:
    """(bins that took the rounding term, largest skipped-bin bound / |S_k|) of the last bin-adaptive
    :func:`project_uniform_device` call with ``nd`` bins on this device (synchronises; for tests and reports)."""
    dev = require_cuda() if dev is None else dev
    ws, _ = _ws(nd, dev)
    out = empty(2, dev)
    _lib.call("b200id_frf_adaptive_diag", ptr(ws), C.c_size_t(ws.numel()), C.c_int(nd), ptr(out), stream_ptr())
    v = out.cpu().numpy()
    return int(v[0]), float(v[1])


def project_uniform_mma_device(t_d, u_d, y_d, w_d, t0: float, dt: float, own=None, sums=None, inv_count: float = 0.0):
    """Same partial sums through the experimental matrix-product evaluation (b200id_frf_project_uniform_mma)."""
    dev = t_d.device
    n, nd = t_d.numel(), w_d.numel()
    own = (0, n) if own is None else own
    out = empty(4 * nd, dev)
    ws, _ = _ws(nd, dev)
    _lib.call("b200id_frf_project_uniform_mma", ptr(t_d), ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int64(own[0]),
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


def cross_power_sums_device(U: torch.Tensor, Y: torch.Tensor, out: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    """out [3*nd] (+)= {Re sum Y conj(U), Im sum Y conj(U), sum |U|^2} over the records in U, Y ([nd] or [R, nd])."""
    U2 = U.reshape(-1, U.shape[-1]).contiguous()
    Y2 = Y.reshape(-1, Y.shape[-1]).contiguous()
    R, nd = U2.shape
    _lib.call("b200id_cross_power_sums", ptr(torch.view_as_real(U2)), ptr(torch.view_as_real(Y2)), C.c_int(R), C.c_int(nd),
              C.c_int(1 if accumulate else 0), ptr(out), stream_ptr())
    return out


def gp_targets_device(sums: torch.Tensor, nd: int, standardise: bool, eps: float = 1e-15):
    """(G [nd] complex, stats [4] or None, y_re [nd], y_im [nd]) from cross-power sums [3*nd]: G = num / den; with
    ``standardise`` the StandardScaler statistics {mean Re, mean Im, std Re, std Im} and the standardised real /
    imaginary targets, else the raw real / imaginary parts as two contiguous vectors."""
    dev = sums.device
    G = empty(2 * nd, dev)
    stats = empty(4, dev) if standardise else None
    y_re, y_im = empty(nd, dev), empty(nd, dev)
    _lib.call("b200id_gp_targets", ptr(sums), C.c_int(nd), C.c_double(eps), ptr(G), ptr(stats), ptr(y_re), ptr(y_im), stream_ptr())
    return torch.view_as_complex(G.view(nd, 2)), stats, y_re, y_im


def record_coefficients_device(t_d, u_d, y_d, w_d, span: float, subtract_mean: bool = True,
                               t0: Optional[float] = None, w_max: Optional[float] = None, path=None):
    """(U, Y) of one resident record.  ``span`` = t[-1] - t[0] (the caller knows it on the host).

    ``path`` = a (kind, t0, dt) tuple from :func:`choose_path` to skip the uniformity probe (callers that
    process the same record repeatedly cache it); None probes."""
    n, nd = t_d.numel(), w_d.numel()
    sums = moments_device(u_d, y_d, (0, n)) if subtract_mean else None
    if path is None:
        w_max = float(w_d.max().item()) if w_max is None else w_max
        t0 = float(t_d[0].item()) if t0 is None else float(t0)
        check_phase_range(w_max, t0, t0 + span)
        path = choose_path(t_d, span, w_max, t0)
    if path[0] == "uniform":
        part = project_uniform_device(t_d, u_d, y_d, w_d, path[1], path[2], (0, n), sums, 1.0 / n)
    else:
        part = project_device(t_d, u_d, y_d, w_d, (0, n), sums, 1.0 / n)
    return finalize_device(part, nd, 2.0 / span, want_y=y_d is not None)


# ----------------------------------------------------------------------------------------- streamed host records
STREAM_MIN_SAMPLES = 1 << 20      # below this one bulk copy + one launch is faster than chunking
STREAM_CHUNKS = 6
STREAM_HEAD_SAMPLES = 4096       # host-side look at the time base before any of it is on the device
_STREAM = os.environ.get("B200ID_FRF_STREAM", "1") != "0"


class StreamedRecord:
    """One HOST-resident record on its way to the device, the projection overlapped with the copy.

    The signals u, y go up first (the mean needs all of them); the time stamps follow in ``n_chunks`` pieces
    on the copy stream and each piece is projected as soon as it has landed, so only the last piece's kernel
    is exposed after the copy.  The path is guessed from the first few thousand time stamps on the host; a
    'uniform' guess is verified on the device over the whole record once it is resident
    (:meth:`uniformity_async`), and the caller re-projects with the general kernel in the rare case that a
    later part is off the grid (:meth:`reproject_general`).

    Two phases so that several records can share the link back to back: :meth:`enqueue_copies` (copy stream,
    returns at once) for every record first, then :meth:`project` (current stream) for each.
    """

    def __init__(self, t_h, u_h, y_h, w_d, span: float, subtract_mean: bool = True, w_max: Optional[float] = None,
                 n_chunks: int = STREAM_CHUNKS, trace: Optional[list] = None, reuse_timebase: bool = False):
        self.dev = dev = w_d.device
        self.th, self.uh = host_f64(t_h), host_f64(u_h)
        self.yh = None if y_h is None else host_f64(y_h)
        self.w_d, self.span, self.subtract_mean, self.trace = w_d, float(span), subtract_mean, trace
        n = self.th.numel()
        self.n, self.nd = n, w_d.numel()
        self.edges = stream_edges(n, n_chunks)
        self.n_chunks = len(self.edges) - 1
        self.u_d = empty(n, dev)
        self.y_d = None if self.yh is None else empty(n, dev)
        timing = trace is not None
        self.landed = [torch.cuda.Event(enable_timing=timing) for _ in range(self.n_chunks)]
        self.signals_up = torch.cuda.Event(enable_timing=timing)
        self.sums = None
        self.t0, self.dt = float(self.th[0]), self.span / (n - 1)
        self.w_max = float(w_d.max().item()) if w_max is None else float(w_max)
        # Path guess from the head of the record on the host (no device round trip); the device check over the
        # whole record is what decides whether the guess stands.
        self.kind = "general"
        if _FORCE_PATH == "uniform":
            self.kind = "uniform"
        elif _FORCE_PATH != "general" and n >= 3 and self.dt > 0.0:
            head = self.th[:min(n, STREAM_HEAD_SAMPLES)].numpy()
            grid = self.t0 + np.arange(head.size, dtype=np.float64) * self.dt
            if float(np.max(np.abs(head - grid))) * self.w_max < UNIFORM_PHASE_TOL:
                self.kind = "uniform"
        check_phase_range(self.w_max, self.t0, self.t0 + self.span)
        # ---- what earlier calls on this very array established (b200id.records.FACTS): a time vector that is
        # exactly fl(fl(i dt) + t0) and passed the uniformity check over its whole length is rebuilt on the device
        # instead of being uploaded, so u and y are the only per-call inputs
        from .records import FACTS, FACTS_ENABLED
        self.facts = FACTS.get(self.th) if (reuse_timebase and FACTS_ENABLED and n >= 3) else None
        self.exact_dt = None
        if self.facts is not None:
            pend = self.facts.pop("exact_pending", None)
            if pend is not None:
                flags = torch.stack([p[0] for p in pend]).cpu().tolist()     # resolved long ago: the call that queued it has ended
                hit = [p[1] for p, f in zip(pend, flags) if f[0] == 0.0]
                self.facts["exact"] = hit[0] if hit else False
            if self.kind == "uniform" and self.facts.get("exact") and self.facts.get("uniform_ok") == (self.t0, self.dt, self.w_max):
                self.exact_dt = float(self.facts["exact"])
        self.t_d = None if self.exact_dt is not None else empty(n, dev)

    def _mark(self, label, stream=None):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream() if stream is None else stream)
            self.trace.append((label, ev))

    def enqueue_copies(self) -> None:
        main, side = torch.cuda.current_stream(), side_stream(self.dev)
        side.wait_stream(main)        # the fresh buffers may still be in use by earlier work on this stream
        self._mark("start")
        if self.trace is not None:
            self.trace.append(("signals_up", self.signals_up))
            self.trace.extend((f"t_landed[{k}]", self.landed[k]) for k in range(self.n_chunks))
        e = self.edges
        if self.exact_dt is not None:
            # known time base: nothing but the signals crosses the link, piece by piece (u, y interleaved)
            with torch.cuda.stream(side):
                for k in range(self.n_chunks):
                    self.u_d[e[k]:e[k + 1]].copy_(self.uh[e[k]:e[k + 1]], non_blocking=True)
                    if self.y_d is not None:
                        self.y_d[e[k]:e[k + 1]].copy_(self.yh[e[k]:e[k + 1]], non_blocking=True)
                    self.landed[k].record(side)
                self.signals_up.record(side)
            return
        with torch.cuda.stream(side):
            self.u_d.copy_(self.uh, non_blocking=True)
            if self.y_d is not None:
                self.y_d.copy_(self.yh, non_blocking=True)
            self.signals_up.record(side)
            for k in range(self.n_chunks):
                self.t_d[e[k]:e[k + 1]].copy_(self.th[e[k]:e[k + 1]], non_blocking=True)
                self.landed[k].record(side)

    def project(self) -> torch.Tensor:
        """Partial sums [4*nd] of the whole record, enqueued on the current stream piece by piece."""
        main = torch.cuda.current_stream()
        n, e = self.n, self.edges
        if self.exact_dt is not None:
            return self._project_known_timebase()
        main.wait_event(self.signals_up)
        self.sums = moments_device(self.u_d, self.y_d, (0, n)) if self.subtract_mean else None
        self._mark("moments_done")
        part = None
        for k in range(self.n_chunks):
            main.wait_event(self.landed[k])
            # a sample's trapezoid weight needs its right neighbour: piece k owns up to its last-but-one sample
            own = (e[k] - (1 if k > 0 else 0), e[k + 1] - 1 if k + 1 < self.n_chunks else n)
            if self.kind == "uniform":
                pk = project_uniform_device(self.t_d, self.u_d, self.y_d, self.w_d, self.t0, self.dt, own, self.sums, 1.0 / n)
            else:
                pk = project_device(self.t_d, self.u_d, self.y_d, self.w_d, own, self.sums, 1.0 / n)
            part = pk if part is None else part.add_(pk)
            self._mark(f"projected[{k}]")
        return part

    def _project_known_timebase(self) -> torch.Tensor:
        """Partial sums when t is rebuilt on the device: every piece of (u, y) is projected RAW as it lands and the
        mean is taken off at the end through linearity,
            sum_i w_i (x_i - m) c_ki = sum_i w_i x_i c_ki - m E_k,   E_k = sum_i w_i c_ki   (same for the sine sums),
        with the window sums E of this (time vector, grid) computed once by projecting a vector of ones and kept
        with the array's facts.  Nothing but the last piece's projection is exposed after the copy."""
        main = torch.cuda.current_stream()
        n, e, nd, dev = self.n, self.edges, self.nd, self.dev
        self.t_d = timebase_fill_device(n, self.t0, self.exact_dt, dev)
        ekey = ("window_sums", self.w_d.data_ptr(), nd, self.t0, self.dt)
        E = self.facts.get(ekey) if self.subtract_mean else None
        if self.subtract_mean and E is None:
            ones = torch.ones(n, dtype=F64, device=dev)
            E = project_uniform_device(self.t_d, ones, None, self.w_d, self.t0, self.dt, (0, n), None, 0.0)[:2 * nd].clone()
            del ones
            self.facts[ekey] = E
        part, sums = None, None
        for k in range(self.n_chunks):
            main.wait_event(self.landed[k])
            own = (e[k], e[k + 1])
            if self.subtract_mean:
                sk = moments_device(self.u_d, self.y_d, own)
                sums = sk if sums is None else sums.add_(sk)
            pk = project_uniform_device(self.t_d, self.u_d, self.y_d, self.w_d, self.t0, self.dt, own, None, 0.0)
            part = pk if part is None else part.add_(pk)
            self._mark(f"projected[{k}]")
        if self.subtract_mean:
            self.sums = sums
            mean = sums / float(n)                                       # [2] = (mean u, mean y)
            part.view(2, 2, nd).sub_(mean.view(2, 1, 1) * E.view(1, 2, nd))
        return part

    def needs_check(self) -> bool:
        return self.kind == "uniform" and _FORCE_PATH != "uniform" and self.exact_dt is None

    def queue_exactness_probe(self) -> None:
        """After a full upload: ask the device whether t is exactly fl(fl(i dt) + t0) for one of the obvious dt
        candidates; the answer is read at the start of the NEXT call on this array (no wait here)."""
        if self.facts is None or "exact" in self.facts or "exact_pending" in self.facts or self.t_d is None or self.n < 3:
            return
        cands = []
        for dt in (float(self.th[1]) - self.t0, self.dt, float(f"{self.dt:.12g}")):
            if dt > 0.0 and dt not in cands:
                cands.append(dt)
        self.facts["exact_pending"] = [(timebase_check_async(self.t_d, self.t0, dt), dt) for dt in cands]

    def remember_uniform_ok(self) -> None:
        """The whole-record uniformity check of this call came back clean."""
        if self.facts is not None and self.kind == "uniform":
            self.facts["uniform_ok"] = (self.t0, self.dt, self.w_max)

    def uniformity_async(self) -> torch.Tensor:
        """max |t_i - (t0 + i dt)| of the resident record as a 1-element device tensor (after :meth:`project`)."""
        return uniformity_device_async(self.t_d, self.t0, self.dt)

    def guess_holds(self, dmax: float) -> bool:
        return dmax * self.w_max < UNIFORM_PHASE_TOL

    def reproject_general(self) -> torch.Tensor:
        self.kind = "general"
        return project_device(self.t_d, self.u_d, self.y_d, self.w_d, (0, self.n), self.sums, 1.0 / self.n)

    def finalize(self, part: torch.Tensor):
        return finalize_device(part, self.nd, 2.0 / self.span, want_y=self.y_d is not None)


def record_coefficients_streamed(t_h, u_h, y_h, w_d, span: float, subtract_mean: bool = True,
                                 w_max: Optional[float] = None, n_chunks: int = STREAM_CHUNKS, trace: Optional[list] = None,
                                 keep: Optional[dict] = None):
    """(U, Y) of one HOST-resident record, with the projection overlapped with the host->device copy
    (:class:`StreamedRecord`).  Same sums as :func:`record_coefficients_device` up to the order of the final
    additions."""
    sr = StreamedRecord(t_h, u_h, y_h, w_d, span, subtract_mean, w_max, n_chunks, trace)
    sr.enqueue_copies()
    part = sr.project()
    if sr.needs_check() and not sr.guess_holds(float(sr.uniformity_async().item())):
        part = sr.reproject_general()
    if keep is not None:
        keep["device_arrays"] = (sr.t_d, sr.u_d, sr.y_d)      # every copy has been waited for on this stream
    return sr.finalize(part)


def _coefficients(t, u, y, w_d, span, subtract_mean, w_max):
    """Dispatch one prepared host record: streamed when it is large enough to pay, else bulk copy."""
    dev = w_d.device
    if _STREAM and not (isinstance(t, torch.Tensor) and t.is_cuda) and len(t) >= STREAM_MIN_SAMPLES:
        return record_coefficients_streamed(t, u, y, w_d, span, subtract_mean, w_max)
    return record_coefficients_device(to_device(t, dev), to_device(u, dev), None if y is None else to_device(y, dev),
                                      w_d, span, subtract_mean, t0=float(t[0]), w_max=w_max)


# ----------------------------------------------------------------------------------------- stored records
def _stored_pair(rec, w: np.ndarray, drop_seconds: float, subtract_mean: bool):
    """(U, Y) of a record held by :mod:`b200id.records`: both signals in one
    fused pass over the resident copy, memoised per (grid, drop, subtract_mean) -- the reference's
    ``synchronous_coefficients_trapz(t, u, ...)`` followed by ``(t, y, ...)`` (data_loader.py:184-185) costs
    one projection, and asking again for the same file costs none.  None when the record needs the host
    path (transient drop on a non-monotone time base)."""
    from .records import STORE
    dev = require_cuda()
    memo_key = (w.tobytes(), float(drop_seconds or 0.0), bool(subtract_mean))
    hit = rec.coefficients.get(memo_key)
    if hit is not None:
        STORE.stats["coefficient_hits"] += 1
        return hit
    t = rec.t
    first = 0
    if drop_seconds and drop_seconds > 0.0:
        if not rec.monotone():
            return None
        first = int(np.searchsorted(t, t[0] + drop_seconds, side="left"))
    if t.size - first < 2:
        raise ValueError("Not enough samples after transient drop.")
    span = float(t[-1] - t[first])
    if span <= 0:
        raise ValueError("Non-positive record length.")
    w_flat = np.ascontiguousarray(w, dtype=np.float64).ravel()
    w_d, w_max = to_device(w_flat, dev), float(np.max(np.abs(w_flat)))
    if rec.device_arrays is None and first == 0 and _STREAM and t.size >= STREAM_MIN_SAMPLES:
        keep = {}
        U, Y = record_coefficients_streamed(rec.t, rec.u, rec.y, w_d, span, subtract_mean, w_max, keep=keep)
        STORE.adopt(rec, keep["device_arrays"])
    else:
        t_d, u_d, y_d = (a[first:] for a in STORE.on_device(rec, dev))
        path = rec.facts.get(("frf_path", first, w_max))
        if path is None:
            path = rec.facts[("frf_path", first, w_max)] = choose_path(t_d, span, w_max, float(t[first]))
        U, Y = record_coefficients_device(t_d, u_d, y_d, w_d, span, subtract_mean, path=path)
    UY = torch.stack([U, Y]).cpu().numpy()
    out = (UY[0].reshape(w.shape), UY[1].reshape(w.shape))
    for a in out:
        a.setflags(write=False)
    rec.coefficients[memo_key] = out
    return out


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
    from .records import STORE
    rec = STORE.lookup(t, x)
    if rec is not None and rec.role_of(t) == "t" and rec.role_of(x) in ("u", "y"):
        pair = _stored_pair(rec, np.asarray(w, dtype=np.float64), drop_seconds, subtract_mean)
        if pair is not None:
            return pair[0 if rec.role_of(x) == "u" else 1].copy()
    t, (x,), span = _prepare(t, [x], drop_seconds)
    w = np.asarray(w, dtype=np.float64)
    U, _ = _coefficients(t, x, None, to_device(w.ravel(), dev), span, subtract_mean, float(np.max(np.abs(w))))
    return U.cpu().numpy().reshape(w.shape)


def demod_pair(t, u, y, w, drop_seconds: float = 0.0, subtract_mean: bool = True):
    """(U, Y) with t, u, y read once (fused form of two synchronous_coefficients_trapz calls)."""
    dev = require_cuda()
    from .records import STORE
    rec = STORE.lookup(t, u, y)
    if rec is not None and (rec.role_of(t), rec.role_of(u), rec.role_of(y)) == ("t", "u", "y"):
        pair = _stored_pair(rec, np.asarray(w, dtype=np.float64), drop_seconds, subtract_mean)
        if pair is not None:
            return pair[0].copy(), pair[1].copy()
    t, (u, y), span = _prepare(t, [u, y], drop_seconds)
    w = np.asarray(w, dtype=np.float64)
    U, Y = _coefficients(t, u, y, to_device(w.ravel(), dev), span, subtract_mean, float(np.max(np.abs(w))))
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
        U, Y = _coefficients(t, u, y, w_d, span, subtract_mean, float(np.max(np.abs(w))))
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
