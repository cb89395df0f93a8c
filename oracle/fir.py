"""Oracle: Hermitian IDFT -> FIR taps -> convolution validation (rows a-9 .. a-11).

Test infrastructure only.  NumPy restatement of ``src/fir_model/fir_helpers.py``,
``fir_fitting.py`` (paper mode) and ``fir_validation.py`` of the reference.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

PAPER_ND = 1000          # fir_fitting.py:197
PAPER_MAX_TAPS = 1024    # fir_fitting.py:228-229


def hermitian_extend(G_pos) -> np.ndarray:
    """Two-sided odd-length spectrum, M = 2 Nd - 1 (fir_helpers.py:75-108).

    X[0] = Re G[0]; X[1:Nd] = G[1:]; X[M-(Nd-1):] = conj(reversed G[1:]).
    """
    g = np.asarray(G_pos, dtype=np.complex128).ravel()
    nd = g.size
    assert nd >= 2
    X = np.zeros(2 * nd - 1, dtype=np.complex128)
    X[0] = g[0].real
    X[1:nd] = g[1:]
    X[nd:] = np.conj(g[1:])[::-1]
    return X


def idft_taps(X, n_taps: int):
    """(h[:n_taps], h_full) with h_full = Re ifft(X) (fir_helpers.py:115-137)."""
    X = np.asarray(X, dtype=np.complex128).ravel()
    M = X.size
    h_full = np.fft.ifft(X, n=M).real
    if n_taps > M:
        raise ValueError(f"N={n_taps} exceeds available length M={M}. Reduce N.")
    return h_full[:n_taps].copy(), h_full


def time_metrics(y_true, y_pred, n_taps: int) -> Dict[str, float]:
    """RMSE / FIT% / R2 with skip = min(10, n_taps // 10); R2 baseline is the first kept sample.

    fir_validation.py:93-136 and fir_fitting.py:288-306.
    """
    skip = min(10, n_taps // 10)
    yt = np.asarray(y_true, dtype=np.float64)[skip:]
    yp = np.asarray(y_pred, dtype=np.float64)[skip:]
    e = yt - yp
    rmse = float(np.sqrt(np.mean(e ** 2)))
    ny = np.linalg.norm(yt)
    fit = float(100 * (1.0 - np.linalg.norm(e) / ny)) if ny > 0 else 0.0
    ss_res = float(np.sum(e ** 2))
    ss_tot = float(np.sum((yt - yt[0]) ** 2))
    r2 = float(1.0 - ss_res / ss_tot) if ss_tot > 0 else 0.0
    return {"rmse": rmse, "fit_percent": fit, "r2": r2, "transient_skip": skip, "detrend": False}


def fir_predict(u, g, n_out: Optional[int] = None) -> np.ndarray:
    """Causal FIR output: np.convolve(u, g, 'full')[:len(y)] (fir_fitting.py:285, fir_validation.py:190)."""
    u = np.asarray(u, dtype=np.float64)
    n_out = u.size if n_out is None else n_out
    return np.convolve(u, np.asarray(g, dtype=np.float64), mode="full")[:n_out]


def fir_validate(u, y, g, detrend: bool = False):
    """(metrics, y_pred) for FIR taps g on record (u, y); fir_validation.py:143-210."""
    u = np.asarray(u, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if detrend:
        u, y = u - np.mean(u), y - np.mean(y)
    yp = fir_predict(u, g, y.size)
    return time_metrics(y, yp, len(g)), yp


def paper_fir(omega, G, predict: Optional[Callable[[np.ndarray], np.ndarray]] = None,
              fir_order: Optional[int] = None):
    """Paper-mode GP -> FIR taps (fir_fitting.py:190-236): (taps, omega_uni, G_uni, X).

    1000-point uniform omega grid over [min, max]; G there from ``predict`` or, without one,
    np.interp of Re/Im with held edges (:208-212); Hermitian extension; IDFT; first
    N = min(fir_order or min(M, 1024), M) taps.
    """
    omega = np.asarray(omega, dtype=np.float64).ravel()
    G = np.asarray(G, dtype=np.complex128).ravel()
    w_uni = np.linspace(float(np.min(omega)), float(np.max(omega)), PAPER_ND)
    if predict is not None:
        G_uni = np.asarray(predict(w_uni), dtype=np.complex128)
    else:
        re = np.interp(w_uni, omega, G.real, left=G.real[0], right=G.real[-1])
        im = np.interp(w_uni, omega, G.imag, left=G.imag[0], right=G.imag[-1])
        G_uni = re + 1j * im
    X = hermitian_extend(G_uni)
    M = X.size
    want = int(fir_order if fir_order is not None else min(M, PAPER_MAX_TAPS))
    taps, _ = idft_taps(X, min(want, M))
    return taps, w_uni, G_uni, X
