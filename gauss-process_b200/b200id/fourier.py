"""Host side of the Fourier (FFT-ratio) estimator: only the FFT bins the interpolation reads, on the GPU.

Mirrors the arithmetic of reference src/frequency_transform/fourier_estimator.py:120-254 and the
``method='fourier'`` branch of transform.py:160-181 through ``b200id_dft_bins`` (include/b200id.h).  The
reference takes two n-point FFTs per record and then linearly interpolates Y/U onto N_d grid frequencies;
the interpolation looks at the two bins either side of each grid point, so at most 2 N_d + 2 of the n/2
bins matter.  Those are evaluated directly (O(n N_d), one pass over u and y, no FFT workspace), which for
N_d ~ 100 costs about as much as one FRF projection and needs no power-of-two length.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .frf import _ws, moments_device
from .runtime import empty, ptr, require_cuda, stream_ptr, to_device

WINDOW_NONE, WINDOW_HANN, WINDOW_GIVEN = 0, 1, 2
U_EPS = 1e-10                     # fourier_estimator.py:203


def positive_bin_count(n: int) -> int:
    """Number of entries of ``fftfreq(n, d)`` that are >= 0 (fourier_estimator.py:163-165)."""
    return (n - 1) // 2 + 1


def bracket_bins(grid: np.ndarray, n: int, dt: float) -> np.ndarray:
    """Sorted unique FFT bin indices that ``np.interp(grid, fftfreq(n, dt)[>= 0], .)`` reads."""
    n_pos = positive_bin_count(n)
    val = 1.0 / (n * dt)                                   # scipy.fft.fftfreq: arange * (1 / (n d))
    j = np.floor(np.asarray(grid, dtype=np.float64) / val).astype(np.int64)
    j = np.clip(j, 0, n_pos - 1)
    for _ in range(3):                                     # settle the floor against the rounded bin frequencies
        j = np.where((j + 1 < n_pos) & ((j + 1) * val <= grid), j + 1, j)
        j = np.where((j > 0) & (j * val > grid), j - 1, j)
    return np.unique(np.concatenate([j, np.minimum(j + 1, n_pos - 1), [0, n_pos - 1]])).astype(np.int64)


def spectra_at_bins_device(u_d: torch.Tensor, y_d: Optional[torch.Tensor], bins: np.ndarray, dt: float,
                           window: Optional[str] = "hann", detrend: bool = True):
    """(U[k], Y[k]) complex128 NumPy at FFT bins ``bins`` of the resident record: mean removal, window with
    the sum(win)/n compensation, dt scaling (fourier_estimator.py:143-169)."""
    dev = u_d.device
    n, nb = u_d.numel(), int(len(bins))
    bins_d = torch.as_tensor(np.ascontiguousarray(bins, dtype=np.int64), device=dev)
    sums = moments_device(u_d, y_d, (0, n)) if detrend else None
    mode, win_d, wsum_d, wsum = WINDOW_NONE, None, None, float(n)
    if window is not None:
        if str(window).lower() in ("hann", "hanning"):
            mode, wsum_d = WINDOW_HANN, empty(2, dev)
        else:
            from scipy.signal import get_window
            win = get_window(window, n)
            mode, win_d, wsum = WINDOW_GIVEN, to_device(win, dev), float(np.sum(win))
    part = empty(4 * nb, dev)
    ws, _ = _ws(nb, dev)
    _lib.call("b200id_dft_bins", ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int64(0), C.c_int64(n), ptr(bins_d), C.c_int(nb),
              C.c_int64(n), C.c_int(mode), ptr(win_d), ptr(sums), C.c_double(1.0 / n), ptr(part), ptr(wsum_d),
              ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    host = (torch.cat([part, wsum_d]) if wsum_d is not None else part).cpu().numpy()
    if wsum_d is not None:
        wsum = float(host[4 * nb])
    S = host[:4 * nb].reshape(4, nb)
    scale = dt if window is None else dt / (wsum / n)
    U = (S[0] - 1j * S[1]) * scale
    Y = (S[2] - 1j * S[3]) * scale if y_d is not None else None
    return U, Y


def transfer_function_on_grid(t, u, y, grid: np.ndarray, window: Optional[str] = "hann",
                              device_arrays=None, dt: Optional[float] = None) -> np.ndarray:
    """``interpolate_to_grid(*compute_transfer_function(t, u, y, window=window), grid)`` of the reference
    (fourier_estimator.py:174-254) without forming the full spectra."""
    dev = require_cuda()
    n = len(u)
    dt = float(np.mean(np.diff(t))) if dt is None else dt
    bins = bracket_bins(grid, n, dt)
    if device_arrays is None:
        u_d, y_d = to_device(np.asarray(u, dtype=np.float64), dev), to_device(np.asarray(y, dtype=np.float64), dev)
    else:
        u_d, y_d = device_arrays
    U, Y = spectra_at_bins_device(u_d, y_d, bins, dt, window)
    G = Y / (U + U_EPS)
    G[np.abs(U) < U_EPS * 10] = 0.0 + 0.0j
    f = bins * (1.0 / (n * dt))
    return np.interp(grid, f, np.real(G)) + 1j * np.interp(grid, f, np.imag(G))


def top_frequency(n: int, dt: float) -> float:
    """Largest non-negative FFT frequency of an n-sample record (``freq[-1]`` in transform.py:170)."""
    return (positive_bin_count(n) - 1) * (1.0 / (n * dt))
