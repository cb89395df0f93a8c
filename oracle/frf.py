"""Oracle: FRF estimation (SURVEY.md section 8 rows a-1 .. a-4).  Test infrastructure only.

Restates, in NumPy, what the reference computes in
``src/frequency_transform/frf_estimator.py`` and ``transform.py``.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence, Tuple

import numpy as np


def freq_grid(nd: int, f_low_log10: float = -1.0, f_up_log10: float = 2.3) -> Tuple[np.ndarray, np.ndarray]:
    """MATLAB-style log grid without the upper endpoint.

    Follows frf_estimator.py:176-183 (``matlab_freq_grid``).
    """
    k = np.arange(nd, dtype=np.float64)
    exponents = f_low_log10 + ((f_up_log10 - f_low_log10) / nd) * k
    hz = np.power(10.0, exponents)
    return hz, 2.0 * np.pi * hz


def _trapz(f: np.ndarray, t: np.ndarray) -> float:
    # numpy.trapz: sum(diff(t) * (f[1:] + f[:-1]) / 2), frf_estimator.py:229-230
    return float((np.diff(t) * (f[1:] + f[:-1]) / 2.0).sum())


def demod_trapz(t, x, w, drop_seconds: float = 0.0, subtract_mean: bool = True,
                threads: int = 1) -> np.ndarray:
    """Synchronous demodulation with trapezoidal weights.

    C(w_k) = (2/T) * [ trapz(x cos(w_k t), t) - j trapz(x sin(w_k t), t) ]
    Follows frf_estimator.py:194-232 (``synchronous_coefficients_trapz``), including the
    transient drop (:212-216), the two ValueErrors (:217-223) and mean removal (:219-220).
    ``threads`` > 1 spreads the per-bin loop (:226-231) over a thread pool; the per-bin
    arithmetic is unchanged, so the result is bit-identical for any thread count.
    """
    t = np.asarray(t, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    if drop_seconds and drop_seconds > 0.0:
        sel = t >= (t[0] + drop_seconds)
        t, x = t[sel], x[sel]
    if t.size < 2:
        raise ValueError("Not enough samples after transient drop.")
    if subtract_mean:
        x = x - np.mean(x)
    span = float(t[-1] - t[0])
    if span <= 0:
        raise ValueError("Non-positive record length.")
    gain = 2.0 / span

    def one_bin(wk: float) -> complex:
        ph = wk * t
        re = gain * _trapz(x * np.cos(ph), t)
        im = -gain * _trapz(x * np.sin(ph), t)
        return re + 1j * im

    out = np.empty(w.shape, dtype=np.complex128)
    if threads <= 1 or w.size <= 1:
        for k, wk in enumerate(w):
            out[k] = one_bin(wk)
    else:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            for k, c in enumerate(pool.map(one_bin, w)):
                out[k] = c
    return out


def cross_power(U_list: Sequence[np.ndarray], Y_list: Sequence[np.ndarray], eps: float = 1e-15) -> np.ndarray:
    """G = sum_r Y_r conj(U_r) / sum_r |U_r|^2, NaN+NaNj where the denominator is <= eps.

    Follows frf_estimator.py:239-252 (``frf_cross_power_average``).
    """
    if not U_list or not Y_list or len(U_list) != len(Y_list):
        raise ValueError("Empty or mismatched lists for cross-power averaging.")
    U = np.vstack(U_list)
    Y = np.vstack(Y_list)
    num = np.sum(Y * np.conj(U), axis=0)
    den = np.sum(np.abs(U) ** 2, axis=0)
    G = np.full(den.shape, np.nan + 1j * np.nan, dtype=np.complex128)
    good = den > eps
    G[good] = num[good] / den[good]
    return G


def estimate_frf(records, nd: int, f_low: float = -1.0, f_up: float = 2.3, drop_seconds: float = 0.0,
                 time_duration: Optional[float] = None, subtract_mean: bool = True, threads: int = 1):
    """FRF of a list of in-memory ``(t, u, y)`` records -> (omega, G).

    Follows transform.py:120-157 (``_estimate_frf``) minus the .mat loading: optional
    ``time_duration`` window (:130-137), argsort when non-monotone (:140-143), per-record
    demodulation of u and y (:145-146), cross-power average (:150), sort by omega (:153-155).
    """
    _, omega = freq_grid(nd, f_low, f_up)
    Us, Ys = [], []
    for (t, u, y) in records:
        t = np.asarray(t, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        drop = drop_seconds
        if time_duration is not None:
            lo = t[0] + drop_seconds
            hi = lo + time_duration
            keep = (t >= lo) & (t <= hi)
            if not np.any(keep):
                raise ValueError(f"No data in [{lo:.2f}, {hi:.2f}]")
            t, u, y = t[keep], u[keep], y[keep]
            drop = 0.0
        if t.size > 1 and not bool(np.all(np.diff(t) > 0)):
            order = np.argsort(t)
            t, u, y = t[order], u[order], y[order]
        Us.append(demod_trapz(t, u, omega, drop, subtract_mean, threads))
        Ys.append(demod_trapz(t, y, omega, drop, subtract_mean, threads))
    G = cross_power(Us, Ys)
    order = np.argsort(omega)
    return omega[order], G[order]
