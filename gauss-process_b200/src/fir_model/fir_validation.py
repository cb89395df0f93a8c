"""Time-domain validation of FIR models (mirror of reference src/fir_model/fir_validation.py).

``validate_fir_with_mat`` runs the causal convolution and the error sums on the GPU (b200id_fir_eval);
``compute_time_domain_metrics`` takes an already computed prediction and reduces it on the device too.
Metric conventions (reference :93-136): skip = min(10, fir_length // 10); RMSE = sqrt(mean(e^2));
FIT = 100 (1 - ||e|| / ||y||); R2 = 1 - sum e^2 / sum (y - y[skip])^2  (baseline = first kept sample).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional

import numpy as np
from scipy.io import loadmat

from b200id import fir as _fir
from b200id import records as _records


def load_mat_time_series(mat_path) -> Dict[str, np.ndarray]:
    """{'t', 'y', 'u'} from a 3xN / Nx3 [t, y, u] array or named variables (reference :30-86).  Parsed once
    per file; the arrays are the record store's read-only vectors (b200id/records.py)."""
    mat_path = Path(mat_path)
    rec = _records.STORE.record(mat_path, "time_series", lambda data: _extract_time_series(data, mat_path))
    return {"t": rec.t, "y": rec.y, "u": rec.u}


def _extract_time_series(data: dict, mat_path):
    """(t, u, y) -- store order -- by the reference's rule: first 3xN / Nx3 array, else named variables."""
    T = y = u = None
    for key, val in data.items():
        if key.startswith("__") or not isinstance(val, np.ndarray):
            continue
        if val.ndim == 2 and (val.shape[0] == 3 or val.shape[1] == 3):
            rows = val if val.shape[0] == 3 else val.T
            T, y, u = (np.ravel(rows[k]).astype(float) for k in range(3))
            break
    if T is None:
        try:
            T = np.ravel(data.get("t", data.get("time"))).astype(float)
            y = np.ravel(data.get("y")).astype(float)
            u = np.ravel(data.get("u")).astype(float)
        except Exception:
            pass
    if T is None or y is None or u is None:
        raise ValueError(f"Could not extract [t, y, u] from {mat_path}. Expected named variables or a 3xN / Nx3 matrix.")
    return T, u, y


def compute_time_domain_metrics(y_true: np.ndarray, y_pred: np.ndarray, fir_length: int) -> Dict[str, float]:
    """RMSE / FIT% / R2 with the minimal transient skip (reference :93-136), reduced on the device:
    an identity tap on y_pred turns the fused FIR kernel into the plain error-sum reduction."""
    y_true = np.asarray(y_true, dtype=float).ravel()
    y_pred = np.asarray(y_pred, dtype=float).ravel()
    metrics, _ = _fir.fir_validate(y_pred, y_true, np.array([1.0]), want_pred=False, skip=min(10, fir_length // 10))
    return metrics


def validate_fir_with_mat(fir_coefficients: np.ndarray, mat_path, output_dir=None, prefix: str = "fir",
                          detrend: bool = False) -> Dict[str, object]:
    """Evaluate FIR taps on the [t, y, u] record of a MAT file (reference :143-210)."""
    g = np.asarray(fir_coefficients, dtype=float).ravel()
    ts = load_mat_time_series(mat_path)
    metrics, y_pred = _fir.fir_validate(ts["u"], ts["y"], g, detrend=detrend, want_pred=output_dir is not None)
    metrics["validation_file"] = str(mat_path)
    print(f"  Validation ({Path(mat_path).name}): RMSE={metrics['rmse']:.3e}, "
          f"FIT={metrics['fit_percent']:.1f}%, R2={metrics['r2']:.3f}")
    if output_dir is not None:
        _plot_validation(ts["t"], ts["y"], y_pred, rmse=metrics["rmse"], output_dir=Path(output_dir), prefix=prefix)
    return metrics


def _plot_validation(t, y, y_pred, rmse, output_dir: Path, prefix: str) -> None:
    """Output-vs-predicted and error PNGs; plotting is outside the hot path (needs matplotlib)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return
    output_dir.mkdir(parents=True, exist_ok=True)
    sel = t <= min(t[-1], t[0] + 10)
    for name, series in (("output_vs_predicted", ((y, "k-", "Measured Output"), (y_pred, "r--", "FIR Predicted"))),
                         ("error", ((y - y_pred, "b-", "Error (Measured - Predicted)"),))):
        fig, ax = plt.subplots(figsize=(12, 8))
        for data, style, label in series:
            ax.plot(t[sel], data[sel], style, label=label)
        ax.set_xlabel("Time [s]")
        ax.set_title(f"FIR Model Validation\nRMSE={rmse:.3e}")
        ax.legend(); ax.grid(True, alpha=0.3)
        plt.tight_layout()
        plt.savefig(output_dir / f"{prefix}_{name}.png", dpi=300, bbox_inches="tight")
        plt.close(fig)
