"""FIR extraction helpers with the reference's surface (mirror of src/fir_model/fir_helpers.py).

``idft_to_fir_coeffs_two_sided`` runs on the GPU (b200id_idft_hermitian) when its input is the
Hermitian spectrum built by ``build_two_sided_hermitian_from_positive``; the small array shuffles
stay on the host.  The legacy irfft spectrum builder (``_build_H_for_irfft``, reference :196-267) is
not part of the paper-mode path the pipeline uses and is not rebuilt.
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Optional, Tuple

import numpy as np
from scipy.io import loadmat  # noqa: F401  (kept: part of the reference module's namespace)

from b200id import records as _records

from b200id import fir as _fir


def build_uniform_omega_linear(omega_meas: np.ndarray, G_meas: np.ndarray, Nd: Optional[int] = None):
    """(omega_uni, G_uni): Nd-point uniform omega grid over [min, max], Re/Im interpolated linearly with
    held edges (reference :33-68)."""
    w = np.asarray(omega_meas, dtype=float).ravel()
    G = np.asarray(G_meas, dtype=complex).ravel()
    assert w.size == G.size and w.size >= 2
    Nd = w.size if Nd is None else Nd
    grid = np.linspace(float(np.min(w)), float(np.max(w)), Nd)
    re = np.interp(grid, w, G.real, left=G.real[0], right=G.real[-1])
    im = np.interp(grid, w, G.imag, left=G.imag[0], right=G.imag[-1])
    return grid, re + 1j * im


def build_two_sided_hermitian_from_positive(G_uni: np.ndarray) -> np.ndarray:
    """Odd-length Hermitian spectrum X (M = 2 Nd - 1): X[0] = Re G[0], X[1:Nd] = G[1:],
    X[Nd:] = conj(G[1:]) reversed (reference :75-108)."""
    G = np.asarray(G_uni, dtype=complex).ravel()
    nd = G.size
    assert nd >= 2
    X = np.zeros(2 * nd - 1, dtype=complex)
    X[0] = G[0].real
    X[1:nd] = G[1:]
    X[nd:] = np.conj(G[1:])[::-1]
    return X


def _is_hermitian_extension(X: np.ndarray) -> bool:
    M = X.size
    if M < 3 or M % 2 == 0:
        return False
    nd = (M + 1) // 2
    return X[0].imag == 0.0 and np.array_equal(X[nd:], np.conj(X[1:nd])[::-1])


def idft_to_fir_coeffs_two_sided(X: np.ndarray, N: int) -> Tuple[np.ndarray, np.ndarray]:
    """(h[:N], h_full) with h_full = Re ifft(X) (reference :115-137); ValueError when N > M."""
    X = np.asarray(X, dtype=complex).ravel()
    M = X.size
    if N > M:
        raise ValueError(f"N={N} exceeds available length M={M}. Reduce N.")
    if not _is_hermitian_extension(X):
        raise NotImplementedError("the device IDFT covers Hermitian spectra from "
                                  "build_two_sided_hermitian_from_positive (the only producer in the pipeline)")
    nd = (M + 1) // 2
    G_pos = np.concatenate(([X[0]], X[1:nd]))
    h_full = _fir.idft_taps(G_pos, M)
    return h_full[:N].copy(), h_full


def _load_timebase_from_mat(mat_file: Path) -> Tuple[np.ndarray, float]:
    """(t, median Ts) from 't' / 'time' or the first row / column of a 3xN / Nx3 array (reference :144-189).
    Parsed and reduced once per file (b200id/records.py); t is read-only."""
    return _records.STORE.file_fact(mat_file, "timebase", _timebase_of)


def _timebase_of(data: dict) -> Tuple[np.ndarray, float]:
    t = None
    for key in ("t", "time"):
        if key in data:
            t = np.ravel(data[key]).astype(float)
            break
    if t is None:
        for key, v in data.items():
            if key.startswith("__") or not isinstance(v, np.ndarray):
                continue
            if v.ndim == 2 and (v.shape[0] == 3 or v.shape[1] == 3):
                t = np.ravel(v[0, :] if v.shape[0] == 3 else v[:, 0]).astype(float)
                break
    if t is None or t.size < 2:
        raise ValueError("Could not infer time vector from MAT file; need 't'/'time' or 3-column array [t,y,u].")
    dt = np.median(np.diff(t))
    if dt <= 0 or not np.isfinite(dt):
        raise ValueError("Invalid time vector; cannot compute Ts.")
    t.setflags(write=False)
    return t, float(dt)


def _build_H_for_irfft(*args, **kwargs):
    raise NotImplementedError("legacy irfft mode (paper_mode=False) is outside the b200id hot path; "
                              "the pipeline only uses paper mode (gp_pipeline.py:257-260)")
