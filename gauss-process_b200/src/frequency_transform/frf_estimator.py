"""FRF estimation with the reference's call surface, computed on the B200.

Mirror of reference src/frequency_transform/frf_estimator.py: same function names, arguments,
return types and exceptions.  The two numerical routines run on the GPU through libb200id.so:

* ``synchronous_coefficients_trapz``  -> b200id_frf_moments / _project / _finalize   (reference :194-232)
* ``frf_cross_power_average``         -> b200id_cross_power                          (reference :239-252)

File discovery, .mat parsing, the frequency grid and the writers are host-side NumPy/SciPy as in the
reference (the grid stays on the host so that the omega bits are identical, reference :176-183).
"""
from __future__ import annotations

import csv
import glob
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.io import loadmat, savemat

from b200id import frf as _frf
from b200id import records as _records


# ------------------------------------------------------------------ files
def resolve_mat_files(targets: Sequence[str], recursive: bool = False) -> List[Path]:
    """Files, directories (with ``recursive``) and glob patterns -> sorted unique .mat paths (reference :40-57)."""
    found = {}
    for target in targets:
        path = Path(target)
        if path.exists():
            if path.is_file():
                candidates = [path]
            elif recursive and path.is_dir():
                candidates = list(path.rglob("*.mat"))
            else:
                candidates = []
        else:
            candidates = [Path(m) for m in glob.glob(target, recursive=recursive)]
        for c in candidates:
            if c.is_file():
                found.setdefault(c.resolve(), None)
    return sorted(found)


def _as_vector(x) -> Optional[np.ndarray]:
    try:
        v = np.asarray(x).squeeze().astype(float).ravel()
    except Exception:
        return None
    return v if v.size > 0 and np.isfinite(v).all() else None


def _split_matrix(arr):
    """3xN or Nx3 numeric array ordered [t, y, u] -> (t, u, y) or None."""
    arr = np.asarray(arr)
    if arr.ndim != 2 or not np.issubdtype(arr.dtype, np.number):
        return None
    if arr.shape[0] == 3:
        rows = (arr[0], arr[1], arr[2])
    elif arr.shape[1] == 3:
        rows = (arr[:, 0], arr[:, 1], arr[:, 2])
    else:
        return None
    t, y, u = (_as_vector(r) for r in rows)
    return (t, u, y) if (t is not None and u is not None and y is not None) else None


def load_time_u_y(mat_path: Path, y_col: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(t, u, y) from a .mat file: named t/u/y (y column ``y_col``), else a [t, y, u] matrix, ``output``
    preferred (reference :70-134).

    The file is parsed once per (path, mtime, size) and the arrays handed out are the record store's
    read-only FP64 vectors (b200id/records.py): passing them on to ``synchronous_coefficients_trapz`` /
    the FIR validation lets those run on the copy already resident in HBM."""
    rec = _records.STORE.record(mat_path, ("time_u_y", int(y_col)), lambda data: _extract_time_u_y(data, mat_path, y_col))
    return rec.t, rec.u, rec.y


def _extract_time_u_y(data: dict, mat_path, y_col: int):
    t, u, y_raw = _as_vector(data.get("t")), _as_vector(data.get("u")), data.get("y")
    if t is not None and u is not None and y_raw is not None:
        ya = np.asarray(y_raw)
        y = None
        if ya.ndim == 1:
            y = _as_vector(ya)
        elif ya.ndim == 2:
            if ya.shape[1] == 1:
                y = _as_vector(ya[:, 0])
            else:
                if not (0 <= y_col < ya.shape[1]):
                    raise ValueError(f"y_col {y_col} out of range for y with shape {ya.shape}")
                y = _as_vector(ya[:, y_col])
        if y is not None and len(t) > 1 and len(u) == len(t) and len(y) == len(t):
            return t, u, y
    if "output" in data:
        got = _split_matrix(data["output"])
        if got is not None:
            return got
    for key, value in data.items():
        if key.startswith("__"):
            continue
        got = _split_matrix(value)
        if got is not None:
            return got
    raise RuntimeError(
        f"Could not locate compatible [t,u,y] in {mat_path}.\n"
        f"Expected either t/u/y variables (with y 1D or 2D) or a 3xN (or Nx3) array with rows/cols [t,y,u].")


def try_load_frequency_from_mat(mat_path: Path) -> Optional[np.ndarray]:
    """omega [rad/s] stored by an earlier run (omega / w, frequency / freq in Hz, or row 0 of P), else None
    (reference :137-169)."""
    try:
        data = loadmat(mat_path)
    except Exception:
        return None
    for key, scale in (("omega", 1.0), ("w", 1.0), ("frequency", 2.0 * np.pi), ("freq", 2.0 * np.pi)):
        if key in data:
            v = np.asarray(data[key]).squeeze()
            if v.size > 0 and np.isfinite(v).all():
                return scale * v.astype(float).ravel() if scale != 1.0 else v.astype(float).ravel()
    if "P" in data:
        P = np.asarray(data["P"])
        if P.ndim == 2 and P.shape[0] >= 1:
            w = np.ravel(P[0, :]).astype(float)
            if np.isfinite(w).all():
                return w
    return None


# ------------------------------------------------------------------ math
def matlab_freq_grid(n_points: int, f_low_log10: float, f_up_log10: float) -> Tuple[np.ndarray, np.ndarray]:
    """(freqs_hz, omega): 10**(f_low + k*(f_up-f_low)/N), k < N, upper end excluded (reference :176-183)."""
    step = (f_up_log10 - f_low_log10) / n_points
    logs = f_low_log10 + step * np.arange(n_points, dtype=np.float64)
    freqs_hz = np.power(10.0, logs)
    return freqs_hz, 2.0 * np.pi * freqs_hz


def describe_timebase(t: np.ndarray) -> Tuple[float, float, bool]:
    """(median dt, span, strictly increasing?) (reference :186-191); remembered for stored records."""
    if t.size <= 1:
        return float("nan"), 0.0, True
    rec = _records.STORE.lookup(t)
    if rec is not None and "timebase" in rec.facts:
        return rec.facts["timebase"]
    d = np.diff(t)
    out = (float(np.median(d)), float(t[-1] - t[0]), bool(np.all(d > 0)))
    if rec is not None and rec.role_of(t) == "t":
        rec.facts["timebase"] = out
        rec.facts["monotone"] = out[2]
    return out


def synchronous_coefficients_trapz(t, x, w, drop_seconds: float = 0.0, subtract_mean: bool = True) -> np.ndarray:
    """C_x(w_k) = (2/T) [trapz(x cos w_k t) - j trapz(x sin w_k t)] on the GPU (reference :194-232).

    Raises ValueError("Not enough samples after transient drop.") / ("Non-positive record length.")
    exactly where the reference does."""
    return _frf.demod(t, x, w, drop_seconds, subtract_mean)


def synchronous_coefficients_trapz_pair(t, u, y, w, drop_seconds: float = 0.0, subtract_mean: bool = True):
    """(C_u, C_y) in one pass over (t, u, y): the fused form of two ``synchronous_coefficients_trapz`` calls."""
    return _frf.demod_pair(t, u, y, w, drop_seconds, subtract_mean)


def frf_cross_power_average(U_list: List[np.ndarray], Y_list: List[np.ndarray], eps: float = 1e-15) -> np.ndarray:
    """G = sum Y conj(U) / sum |U|^2, NaN+NaNj where the denominator <= eps (reference :239-252)."""
    return _frf.cross_power(U_list, Y_list, eps)


# ------------------------------------------------------------------ writers (host I/O, same formats)
def save_csv(prefix: Path, w: np.ndarray, G: np.ndarray) -> None:
    with open(str(prefix) + "_frf.csv", "w", newline="") as fp:
        out = csv.writer(fp)
        out.writerow(["omega_rad_s", "freq_Hz", "ReG", "ImG", "absG", "phase_rad"])
        for wk, gk in zip(w, G):
            if np.isfinite(gk):
                out.writerow([wk, wk / (2.0 * np.pi), np.real(gk), np.imag(gk), np.abs(gk), np.angle(gk)])


def _matlab_P(w, G):
    return np.vstack((w, np.abs(G), np.arctan2(-np.imag(G), -np.real(G))))   # MATLAB phase convention


def save_mat(prefix: Path, w: np.ndarray, G: np.ndarray) -> None:
    savemat(str(prefix) + "_frf.mat", {"w": w, "f": w / (2.0 * np.pi), "G": G, "P": _matlab_P(w, G)})


def save_matlab_dat(prefix: Path, w: np.ndarray, G: np.ndarray) -> None:
    np.savetxt(str(prefix) + "_frf.dat", _matlab_P(w, G), delimiter=",")


def save_plots(prefix: Path, w: np.ndarray, G: np.ndarray) -> None:
    """Bode gain / Nyquist / Bode phase PNGs; needs matplotlib (plotting is out of the hot path)."""
    import matplotlib.pyplot as plt
    for suffix, xs, ys, logx in (("_bode_mag.png", w, 20.0 * np.log10(np.abs(G)), True),
                                 ("_nyquist.png", np.real(G), np.imag(G), False),
                                 ("_bode_phase.png", w, np.unwrap(np.angle(G)), True)):
        fig, ax = plt.subplots()
        (ax.semilogx if logx else ax.plot)(xs, ys, marker="*", linestyle="None")
        ax.grid(True, which="both")
        fig.tight_layout()
        fig.savefig(str(prefix) + suffix, dpi=150)
        plt.close(fig)
