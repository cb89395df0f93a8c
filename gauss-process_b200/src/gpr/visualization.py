"""GP result metrics and the Nyquist plot (mirror of reference src/gpr/visualization.py).

``plot_gp_results`` only computes RMSE / R^2 (as in the reference, :19-47).  ``plot_complex_gp`` and
``configure_plot_style`` are plotting, outside the hot path: they use matplotlib when it is installed
and are no-ops otherwise.
"""
from pathlib import Path
from typing import Dict, Optional

import numpy as np


def configure_plot_style() -> None:
    try:
        from src.visualization.plot_styles import configure_plot_style as _ref
        _ref()
    except Exception:
        pass


def plot_gp_results(omega, y_true, y_pred, y_std, title, output_path) -> Dict[str, float]:
    """{'rmse', 'r2'} between truth and prediction; the other arguments are kept for API compatibility."""
    res = np.asarray(y_true) - np.asarray(y_pred)
    rmse = np.sqrt(np.mean(res ** 2))
    r2 = 1 - np.sum(res ** 2) / np.sum((y_true - np.mean(y_true)) ** 2)
    return {"rmse": rmse, "r2": r2}


def plot_complex_gp(omega, G_true, G_pred, G_std_real: Optional[np.ndarray], G_std_imag: Optional[np.ndarray],
                    output_prefix: Path, save_eps: bool = True) -> None:
    """Nyquist diagram of measured vs GP mean with 2-sigma ellipses (reference :50-111)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return
    configure_plot_style()
    fig, ax = plt.subplots(figsize=(10, 10))
    ax.plot(np.real(G_true), np.imag(G_true), "k.", markersize=14, label="Measured", alpha=0.6)
    ax.plot(np.real(G_pred), np.imag(G_pred), "r-", linewidth=4.0, label="GP mean")
    if G_std_real is not None and G_std_imag is not None:
        ang = np.linspace(0, 2 * np.pi, 100)
        for i in np.linspace(0, len(omega) - 1, min(20, len(omega)), dtype=int):
            ax.plot(np.real(G_pred[i]) + 2 * G_std_real[i] * np.cos(ang),
                    np.imag(G_pred[i]) + 2 * G_std_imag[i] * np.sin(ang), "r-", alpha=0.2, linewidth=1.0)
    ax.set_xlabel("Real{G(jw)}"); ax.set_ylabel("Imag{G(jw)}")
    ax.set_title("Nyquist Plot with GP Regression")
    ax.legend(loc="best"); ax.grid(True, alpha=0.3); ax.axis("equal")
    plt.tight_layout()
    plt.savefig(str(output_prefix) + "_nyquist_gp.png", dpi=300, bbox_inches="tight")
    if save_eps:
        plt.savefig(str(output_prefix) + "_nyquist_gp.eps", format="eps", bbox_inches="tight")
    plt.close(fig)
