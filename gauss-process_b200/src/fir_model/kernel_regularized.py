"""Kernel-regularised FIR identification in the time domain with DC, SS(2) and SI kernels -- drop-in for the
reference's src/fir_model/kernel_regularized.py (same names, arguments, dictionary keys, files written and printed
warnings), with the arithmetic on the device through b200id.kr:

* regressor summaries G = U^T U, b = U^T y, yy (reference :337-340) -- ``b200id_kr_summaries``;
* ``_nll`` (:224-261) -- ``b200id_kr_nll_batch``; inside ``run_fir_experiment`` SciPy's L-BFGS-B stays the host driver
  and the forward-difference points of each gradient go to the device as ONE batch (SciPy >= 1.16 ``workers`` hook);
* ``_posterior_mean_g`` (:264-291) -- ``b200id_kr_posterior``;
* y_hat = U g_hat (:386) -- the FIR convolution kernel (``b200id_fir_eval``).

There is no CPU fallback: a missing CUDA library raises.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.io import loadmat
from scipy.optimize import minimize

from b200id import fir as _fir
from b200id import kr as _kr

_KDIM = {"dc": 3, "ss": 2, "si": 4}


def _load_io_mat(io_mat: Path, nmax: Optional[int] = 100_000) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Load (t, u, y) from a MAT file holding a 3 x N or N x 3 matrix [time; y; u], truncated to nmax samples."""
    io_data = loadmat(io_mat)
    mat = None
    for name, arr in io_data.items():
        if not name.startswith("__"):
            mat = arr
            break
    if mat is None:
        raise ValueError("MAT file does not contain a numeric array")
    if mat.ndim != 2 or (mat.shape[0] != 3 and mat.shape[1] != 3):
        raise ValueError("Expected a 3xN or Nx3 matrix with rows [time; y; u]")
    t_all, y_all, u_all = (mat[k, :].ravel() for k in range(3)) if mat.shape[0] == 3 else (mat[:, k].ravel() for k in range(3))
    if nmax is not None and len(u_all) > nmax:
        t_all, y_all, u_all = t_all[:nmax], y_all[:nmax], u_all[:nmax]
    return t_all.astype(float), u_all.astype(float), y_all.astype(float)


def _build_regressor(u: np.ndarray, L: int) -> np.ndarray:
    """Regressor U with rows [u[n], u[n-1], ..., u[n-L+1]] (zeros before the record), shape (N, L).  Host view for
    callers that want the matrix itself; the identification below never forms it."""
    from numpy.lib.stride_tricks import sliding_window_view
    N = len(u)
    u_pad = np.concatenate([np.zeros(L - 1, dtype=float), np.asarray(u, dtype=float)])
    return sliding_window_view(u_pad, window_shape=L)[:N, ::-1]


def _sigmoid(x: float) -> float:
    return 1.0 / (1.0 + math.exp(-x))


def _logit(p: float) -> float:
    return math.log(p / (1.0 - p))


def _kernel_dc(L: int, c: float, lam: float, rho: float) -> np.ndarray:
    """k(i, j) = c lam^((i+j)/2) rho^|i-j| on the device."""
    return _kr.build_kernel(L, "dc", [math.log(c), _logit(lam), math.atanh(rho)])


def _kernel_ss2(L: int, c: float, lam: float) -> np.ndarray:
    """Second-order stable spline K = c (C H)(C H)^T on the device."""
    return _kr.build_kernel(L, "ss", [math.log(c), _logit(lam)])


def _kernel_si(L: int, r: float, omega: float, q: float, barlam: float) -> np.ndarray:
    """Simulation-induced kernel on the device."""
    return _kr.build_kernel(L, "si", [_logit(r), _logit(omega / math.pi), math.log(q), _logit(barlam)])


def _build_kernel(L: int, kernel: str, theta: np.ndarray) -> np.ndarray:
    """K for the selected kernel and unconstrained theta."""
    if kernel.lower() not in _KDIM:
        raise ValueError(f"Unknown kernel: {kernel}")
    return _kr.build_kernel(L, kernel, np.asarray(theta, dtype=float)[:_KDIM[kernel.lower()]])


def _nll(theta_all: np.ndarray, G: np.ndarray, b: np.ndarray, yy: float, N: int, L: int, kernel: str,
         jitter: float = 1e-10) -> float:
    """Negative log marginal likelihood (up to a constant) at one point, host summaries in."""
    if kernel.lower() not in _KDIM:
        raise ValueError("kernel")
    val, info = _kr.nll_batch(kernel, [theta_all], _kr.summaries_from_host(G, b, yy, N), jitter)
    if info[0] in (1, 3):
        raise np.linalg.LinAlgError("matrix not positive definite")
    return float(val[0])


def _posterior_mean_g(theta_all: np.ndarray, G: np.ndarray, b: np.ndarray, L: int, kernel: str) -> Tuple[np.ndarray, float, np.ndarray]:
    """(g_hat, sigma2, K) for given parameters, host summaries in."""
    if kernel.lower() not in _KDIM:
        raise ValueError("kernel")
    return _kr.posterior(kernel, theta_all, _kr.summaries_from_host(G, b, 0.0, 1))


@dataclass
class FIRConfig:
    io_mat: Path = Path("./fir/data/data_hour.mat")
    out_dir: Path = Path("./test_output")
    kernel: str = "dc"     # 'dc' | 'ss' | 'si'
    L: int = 128           # FIR length (effective memory)
    multi_starts: int = 3
    optimize: bool = True
    maxiter: int = 200
    demean: bool = True


def _lbfgsb_takes_workers() -> bool:
    try:
        import inspect
        from scipy.optimize import _lbfgsb_py
        return "workers" in inspect.signature(_lbfgsb_py._minimize_lbfgsb).parameters
    except Exception:
        return False


_LBFGSB_TAKES_WORKERS = _lbfgsb_takes_workers()


def _save_figures(cfg, t, y_raw, yhat, e, g_hat, rmse, fit, r2, fig_path, fig2_path) -> None:
    """fir_results.png (output vs prediction, error) and fir_impulse.png (reference :404-432); plotting is outside the hot
    path and is skipped where matplotlib is not installed."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    fig, (ax1, ax2) = plt.subplots(2, 1, figsize=(10, 7))
    ax1.plot(t, y_raw, label="Measured y", color="k")
    ax1.plot(t, yhat, label="Predicted yhat", color="r", linestyle="--")
    ax1.set_title(f"Output vs Prediction ({cfg.kernel.upper()} kernel)")
    ax1.set_xlabel("time [s]"); ax1.set_ylabel("y"); ax1.legend(); ax1.grid(True)
    ax2.plot(t, e, label="error", color="b")
    ax2.set_title(f"Error  RMSE={rmse:.3e}, FIT={fit:.2f}%, R2={r2:.3f}")
    ax2.set_xlabel("time [s]"); ax2.set_ylabel("e"); ax2.legend(); ax2.grid(True)
    fig.tight_layout()
    fig.savefig(fig_path, dpi=300)
    plt.close(fig)
    fig2, ax = plt.subplots(1, 1, figsize=(10, 4))
    ax.stem(np.arange(cfg.L), g_hat, linefmt="C0-", markerfmt="C0o", basefmt="k-")
    ax.set_title(f"Estimated impulse response g (L={cfg.L})")
    ax.set_xlabel("lag k"); ax.set_ylabel("g[k]"); ax.grid(True)
    fig2.tight_layout()
    fig2.savefig(fig2_path, dpi=300)
    plt.close(fig2)


def run_fir_experiment(cfg: FIRConfig) -> Dict[str, object]:
    """Run the kernel-regularised FIR identification and export figures + metrics (reference :310-460)."""
    cfg.out_dir.mkdir(parents=True, exist_ok=True)
    if Path(cfg.io_mat).exists():
        t, u_raw, y_raw = _load_io_mat(cfg.io_mat)
    else:
        print("[Warning] I/O MAT file not found, using synthetic data")
        N = 5000
        t = np.arange(N, dtype=float) * 1.0
        rng = np.random.default_rng(0)
        u_raw = rng.standard_normal(N).astype(float)
        L0 = min(cfg.L, 64)
        g0 = np.exp(-0.05 * np.arange(L0)) * (np.sin(0.2 * np.arange(L0)))
        y_clean = np.convolve(u_raw, g0, mode="full")[:N]
        y_raw = (y_clean + 0.02 * np.std(y_clean) * rng.standard_normal(N)).astype(float)
    N = len(u_raw)
    if cfg.demean:
        u_mean, y_mean = float(np.mean(u_raw)), float(np.mean(y_raw))
        u, y = u_raw - u_mean, y_raw - y_mean
    else:
        u, y = u_raw.copy(), y_raw.copy()

    summ = _kr.summaries(u, y, cfg.L)                  # G, b on the device; yy, N on the host

    kname = cfg.kernel.lower()
    if kname == "dc":
        theta0 = np.array([math.log(1.0), math.log(0.98 / (1 - 0.98)), np.arctanh(0.8)])
    elif kname == "ss":
        theta0 = np.array([math.log(1.0), math.log(0.96 / (1 - 0.96))])
    elif kname == "si":
        theta0 = np.array([math.log(0.98 / (1 - 0.98)), math.log(0.25), math.log(1e-2), math.log(0.96 / (1 - 0.96))])
    else:
        raise ValueError("Unknown kernel")
    theta0 = np.concatenate([theta0, np.array([math.log(1e-2)])])

    def nll_many(points):
        vals, info = _kr.nll_batch(cfg.kernel, np.asarray(points, dtype=float), summ)
        if np.any((info == 1) | (info == 3)):
            raise np.linalg.LinAlgError("matrix not positive definite")
        return vals

    def batched_map(_fun, points):
        points = list(points)
        return [float(v) for v in nll_many(points)] if points else []

    options = {"maxiter": cfg.maxiter, "ftol": 1e-9, "gtol": 1e-6, "maxcor": 20}
    if _LBFGSB_TAKES_WORKERS:
        options["workers"] = batched_map

    best = (float("inf"), theta0.copy())
    rng = np.random.default_rng(0)
    for s in range(max(1, int(cfg.multi_starts))):
        th_init = theta0.copy() if s == 0 else theta0 + 0.2 * rng.standard_normal(theta0.shape)
        res = minimize(lambda th: float(nll_many([th])[0]), th_init, method="L-BFGS-B", options=options)
        if res.fun < best[0]:
            best = (res.fun, res.x.copy())
    theta_hat = best[1]

    g_hat, sig2_hat, K_hat = _kr.posterior(cfg.kernel, theta_hat, summ)

    yhat = _fir.fir_predict(u, g_hat)                  # U @ g_hat: causal FIR with zeros before the record
    if cfg.demean:
        yhat = yhat + y_mean

    e = y_raw - yhat
    e_tail, y_tail = e[cfg.L - 1:], y_raw[cfg.L - 1:]
    rmse = float(np.sqrt(np.mean(e_tail ** 2)))
    nrmse = float(1.0 - np.linalg.norm(e_tail) / (np.linalg.norm(y_tail - np.mean(y_tail)) + 1e-12))
    r2 = float(1.0 - np.sum(e_tail ** 2) / (np.sum((y_tail - np.mean(y_tail)) ** 2) + 1e-12))
    fit = float(100.0 * (1.0 - np.linalg.norm(e_tail) / (np.linalg.norm(y_tail - np.mean(y_tail)) + 1e-12)))

    fig_path, fig2_path = cfg.out_dir / "fir_results.png", cfg.out_dir / "fir_impulse.png"
    _save_figures(cfg, t, y_raw, yhat, e, g_hat, rmse, fit, r2, fig_path, fig2_path)

    np.savez(cfg.out_dir / "fir_coefficients.npz", g_hat=g_hat, K=K_hat, kernel=cfg.kernel, L=cfg.L)
    metrics = {"rmse": rmse, "nrmse": nrmse, "r2": r2, "fit_percent": fit, "L": int(cfg.L), "sigma2": float(sig2_hat),
               "kernel": cfg.kernel}
    with open(cfg.out_dir / "fir_metrics.json", "w", encoding="utf-8") as f:
        json.dump(metrics, f, ensure_ascii=False, indent=2)
    return {"results_png": str(fig_path), "impulse_png": str(fig2_path), "coeff_npz": str(cfg.out_dir / "fir_coefficients.npz"),
            **metrics}


def _parse_argv() -> Optional[FIRConfig]:
    import argparse
    p = argparse.ArgumentParser(description="Kernel-regularized FIR identification")
    p.add_argument("--io", type=str, default="./fir/data/data_hour.mat", help="MAT file with [time;y;u]")
    p.add_argument("--out", type=str, default="./test_output_test", help="Output directory")
    p.add_argument("--kernel", type=str, default="dc", choices=["dc", "ss", "si"], help="Kernel type")
    p.add_argument("--L", type=int, default=128, help="FIR length (memory)")
    p.add_argument("--starts", type=int, default=3, help="Number of multi-starts for ML")
    p.add_argument("--noopt", action="store_true", help="Disable hyperparameter optimization")
    args = p.parse_args()
    return FIRConfig(io_mat=Path(args.io), out_dir=Path(args.out), kernel=args.kernel, L=int(args.L),
                     multi_starts=int(args.starts), optimize=(not args.noopt))


if __name__ == "__main__":
    cfg = _parse_argv()
    res = run_fir_experiment(cfg)
    print(json.dumps(res, indent=2))
