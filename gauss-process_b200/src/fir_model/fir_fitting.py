"""GP -> FIR conversion, paper mode, with the reference's surface (mirror of
src/fir_model/fir_fitting.py:143-333).

Steps (reference :190-306): 1000-point uniform omega grid over [min, max]; G there from the GP
predictor (or linear interpolation); two-sided Hermitian extension (M = 1999); IDFT on the device,
first N = min(fir_order or min(M, 1024), M) taps; validation on the MAT record by causal convolution
+ fused error sums on the device (skip = min(10, N // 10), no detrend).  Result-dict keys are the
reference's.  Legacy irfft mode (paper_mode=False) is not on the hot path.
"""
from __future__ import annotations

import warnings
from pathlib import Path
from typing import Callable, Dict, Optional

import numpy as np
from scipy.io import loadmat

from b200id import fir as _fir
from src.fir_model.fir_helpers import (  # noqa: F401
    build_uniform_omega_linear,
    build_two_sided_hermitian_from_positive,
    idft_to_fir_coeffs_two_sided,
    _load_timebase_from_mat,
)

PAPER_ND = 1000
PAPER_MAX_TAPS = 1024


def _extract_tyu_from_mat_data(data: dict):
    """(T, y, u) from a loadmat dict: first 3xN / Nx3 array, else named t/time, y, u; (None,)*3 on failure."""
    for key, val in data.items():
        if key.startswith("__") or not isinstance(val, np.ndarray):
            continue
        if val.ndim == 2 and (val.shape[0] == 3 or val.shape[1] == 3):
            rows = val if val.shape[0] == 3 else val.T
            return tuple(np.ravel(rows[k]).astype(float) for k in range(3))
    try:
        return (np.ravel(data.get("t", data.get("time"))).astype(float), np.ravel(data.get("y")), np.ravel(data.get("u")))
    except Exception:
        return None, None, None


def _validation_record(mat_file):
    """(T, y, u) of the validation file through the record store: the same vectors (and HBM copies) that
    ``fir_validation.load_mat_time_series`` hands out, so FRF, paper-mode FIR and Wave.mat validation of one
    file share one parse and one upload (the reference parses it three times: data_loader.py:170,
    fir_fitting.py:267, fir_validation.py:163)."""
    from src.fir_model.fir_validation import load_mat_time_series
    try:
        ts = load_mat_time_series(mat_file)
    except ValueError:
        return None, None, None
    return ts["t"], ts["y"], ts["u"]


def plot_gp_fir_results_fixed(t, y, y_pred, u, rmse, fit_percent, r2, output_dir: Path,
                              prefix: str = "gp_fir_fixed", save_eps: bool = True):
    """Output-vs-predicted and error figures; plotting is outside the hot path (needs matplotlib)."""
    from src.fir_model.fir_validation import _plot_validation
    _plot_validation(np.asarray(t), np.asarray(y), np.asarray(y_pred), rmse, Path(output_dir), prefix)


def gp_to_fir_direct_pipeline(
    omega: np.ndarray,
    G: np.ndarray,
    gp_predict_func: Optional[Callable[[np.ndarray], np.ndarray]] = None,
    mat_file: Optional[Path] = None,
    output_dir: Optional[Path] = None,
    fir_length: int = 1024,
    N_fft: Optional[int] = None,
    paper_mode: bool = True,
    fir_order: Optional[int] = None,
) -> Dict[str, object]:
    omega = np.asarray(omega, dtype=float).ravel()
    G = np.asarray(G, dtype=complex).ravel()
    output_dir = Path("fir_gp_fixed") if output_dir is None else Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    if not paper_mode:
        raise NotImplementedError("legacy irfft mode (paper_mode=False) is outside the b200id hot path; "
                                  "the pipeline only calls paper mode (gp_pipeline.py:257-260)")

    print("[Paper Mode] Using paper-based FIR procedure:")
    print("  Step 1: Uniform linear omega grid + GP prediction")
    print("  Step 2: Two-sided Hermitian spectrum (odd-length)")
    print("  Step 3: IDFT -> first N taps as FIR coefficients")
    omega_uni = np.linspace(float(np.min(omega)), float(np.max(omega)), PAPER_ND)
    if gp_predict_func is not None:
        print("  Using GP predictor for uniform grid (higher accuracy)")
        G_uni = np.asarray(gp_predict_func(omega_uni), dtype=complex).ravel()
    else:
        print("  Using linear interpolation (GP predictor not available)")
        G_uni = (np.interp(omega_uni, omega, G.real, left=G.real[0], right=G.real[-1])
                 + 1j * np.interp(omega_uni, omega, G.imag, left=G.imag[0], right=G.imag[-1]))
    print(f"  Nd = {PAPER_ND} frequency points")
    M = 2 * G_uni.size - 1
    print(f"  M = {M} (IDFT length, odd)")
    wanted = int(fir_order if fir_order is not None else min(M, PAPER_MAX_TAPS))
    N = min(wanted, M)
    if N < wanted:
        warnings.warn(f"Requested fir_order={wanted} exceeds M={M}. Using N={N} instead.")
    # Hermitian extension + IDFT in one device kernel (X[0] = Re G[0] is applied there)
    g = _fir.idft_taps(G_uni, N)
    print(f"  N = {N} FIR taps extracted (requested: {wanted})")

    results: Dict[str, object] = {
        "paper_mode": True, "Nd": int(PAPER_ND), "M_idft": int(M), "fir_order": int(N), "fir_coefficients": g,
        "method": "gp_paper_mode", "omega_range": [float(omega_uni[0]), float(omega_uni[-1])],
        "rmse": None, "fit_percent": None, "r2": None,
    }
    if mat_file is not None and Path(mat_file).exists():
        _, Ts = _load_timebase_from_mat(Path(mat_file))
        print(f"  Ts = {Ts:.6f} s (from MAT file)")
        results["Ts"] = float(Ts)
        T, y, u = _validation_record(mat_file)
        if T is not None and y is not None and u is not None:
            print("  [STRICT MODE] No detrend - evaluating full frequency range including DC")
            metrics, y_pred = _fir.fir_validate(u, y, g, detrend=False, want_pred=True)
            print(f"  Transient skip: {metrics['transient_skip']} samples (was: {N // 2} samples)")
            results.update({"rmse": metrics["rmse"], "fit_percent": metrics["fit_percent"], "r2": metrics["r2"],
                            "transient_skip": metrics["transient_skip"], "detrend": False})
            print(f"  Validation (STRICT): RMSE={metrics['rmse']:.3e}, FIT={metrics['fit_percent']:.1f}%, "
                  f"R2={metrics['r2']:.3f}")
            plot_gp_fir_results_fixed(t=T, y=y, y_pred=y_pred, u=u, rmse=metrics["rmse"],
                                      fit_percent=metrics["fit_percent"], r2=metrics["r2"],
                                      output_dir=output_dir, prefix="gp_fir_paper")
        else:
            print("  Warning: Could not load validation data from MAT file")
    else:
        print("  No validation MAT file provided")
        results["Ts"] = None
    print("[Paper Mode] FIR extraction complete")
    return results
