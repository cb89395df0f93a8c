"""Oracle: FFT-based transfer-function estimator.  Test infrastructure only.

Restates, in NumPy / SciPy, what the reference computes in
``src/frequency_transform/fourier_estimator.py`` and the ``method='fourier'`` branch of ``transform.py``.
Pinned against the reference itself by tests/golden/fourier.npz (oracle/make_golden.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.fft import fft, fftfreq
from scipy.signal import get_window


def apply_time_window(t, u, y, drop_seconds: float = 0.0, time_duration: Optional[float] = None):
    """fourier_estimator.py:80-114: drop ``t < t0 + drop`` first, then keep ``t <= t[0] + duration`` of the rest."""
    if drop_seconds > 0:
        keep = t >= (t[0] + drop_seconds)
        t, u, y = t[keep], u[keep], y[keep]
    if time_duration is not None:
        keep = t <= t[0] + time_duration
        t, u, y = t[keep], u[keep], y[keep]
    return t, u, y


def spectrum(t, x, window: Optional[str] = "hann", detrend: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """fourier_estimator.py:120-171 with n_freq = n: mean removal, window / (sum(win)/n), FFT, bins with
    fftfreq >= 0, scaled by dt = mean(diff(t))."""
    dt = np.mean(np.diff(t))
    n = len(x)
    if detrend:
        x = x - np.mean(x)
    if window is not None:
        win = get_window(window, n)
        x = x * win
        x = x / (np.sum(win) / n)
    spec = fft(x, n=n)
    freq = fftfreq(n, dt)
    pos = freq >= 0
    return freq[pos], spec[pos] * dt


def transfer_function(t, u, y, window: Optional[str] = "hann") -> Tuple[np.ndarray, np.ndarray]:
    """fourier_estimator.py:174-211: G = Y / (U + 1e-10), zero where |U| < 1e-9."""
    f, U = spectrum(t, u, window)
    _, Y = spectrum(t, y, window)
    eps = 1e-10
    G = Y / (U + eps)
    G[np.abs(U) < eps * 10] = 0.0 + 0.0j
    return f, G


def interpolate(f, G, grid) -> np.ndarray:
    """fourier_estimator.py:236-254: np.interp on real and imaginary parts separately."""
    return np.interp(grid, f, np.real(G)) + 1j * np.interp(grid, f, np.imag(G))


def estimate(records: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], nd: int, f_min: float, f_max: Optional[float],
             drop_seconds: float = 0.0, time_duration: Optional[float] = None, window: str = "hann"):
    """transform.py:160-181: per-record transfer function, linear grid up to min(f_max, largest available
    frequency), interpolation, complex mean over records.  Returns (f_grid, G)."""
    win = None if window == "none" else window
    per_file, top = [], 0.0
    for t, u, y in records:
        t, u, y = apply_time_window(t, u, y, drop_seconds, time_duration)
        f, G = transfer_function(t, u, y, win)
        top = max(top, f[-1])
        per_file.append((f, G))
    grid = np.linspace(f_min, min(f_max, top) if f_max is not None else top, nd)
    Gs = [interpolate(f, G, grid) for f, G in per_file]
    return grid, (np.mean(np.vstack(Gs), axis=0) if len(Gs) > 1 else Gs[0])
