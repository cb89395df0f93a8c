"""Grid search over kernel hyperparameters, every candidate evaluated in one batched GPU launch.

Mirror of reference src/gpr/grid_search.py: ``get_default_param_grids`` (:21-45) and
``grid_search_hyperparameters`` (:48-232).  The candidate list is generated on the host exactly as
the reference does (itertools.product in parameter order, ``random.seed(42); random.sample`` above
``max_combinations``); b200id_gp_batch_eval computes the metric of every candidate (validation RMSE
when validation data is given, else NLL; pivot <= 0 -> 1e10) and b200id_first_strict_min reproduces
the sequential ``metric < best`` scan (:186-196).  Noise variance is held fixed (:139-141).
"""
from itertools import product
from typing import Dict, Optional

import numpy as np

from b200id import gp as _gp
from src.gpr.kernels import Kernel

GRID_POINTS = 30


def _axis(low: float, high: float) -> np.ndarray:
    if low > 0 and high / low > 100:                       # wide positive range -> log spacing
        return np.logspace(np.log10(low), np.log10(high), GRID_POINTS)
    return np.linspace(low, high, GRID_POINTS)


def get_default_param_grids(kernel: Kernel) -> Dict:
    """{'param_i': 30-point grid} from the kernel's bounds (reference :21-45)."""
    return {f"param_{i}": _axis(lo, hi) for i, (lo, hi) in enumerate(kernel.bounds)}


def _scaler_stats(y_scaler):
    """(mean, scale) of a fitted sklearn StandardScaler (inverse_transform = x * scale_ + mean_)."""
    if y_scaler is None:
        return None, None
    mean = float(np.ravel(y_scaler.mean_)[0]) if getattr(y_scaler, "with_mean", True) else 0.0
    scale = float(np.ravel(y_scaler.scale_)[0]) if getattr(y_scaler, "with_std", True) else 1.0
    return mean, scale


def grid_search_hyperparameters(kernel: Kernel,
                                X_train: np.ndarray,
                                y_train: np.ndarray,
                                noise_variance: float,
                                param_grids: Optional[Dict] = None,
                                max_combinations: int = 5000,
                                validation_X: Optional[np.ndarray] = None,
                                validation_y: Optional[np.ndarray] = None,
                                y_scaler: object = None):
    """Set ``kernel``'s parameters to the grid point with the smallest validation RMSE / NLL."""
    import random

    use_validation = validation_X is not None and validation_y is not None
    metric_name = "RMSE" if use_validation else "NLL"
    if use_validation:
        scale_txt = "original scale" if y_scaler is not None else "normalized scale"
        print(f"  Grid search: Using validation data (N={len(validation_y)}) for error-based evaluation ({scale_txt})")
    else:
        print("  Grid search: Using log marginal likelihood for evaluation")
    if param_grids is None:
        param_grids = get_default_param_grids(kernel)
    print(f"  Noise variance: {noise_variance:.3e} (fixed, not optimized)")

    y_mean, y_scale = _scaler_stats(y_scaler)
    val = dict(Xv=validation_X, yv=validation_y, y_mean=y_mean, y_scale=y_scale) if use_validation else {}

    start = kernel.get_params()
    axes = [param_grids.get(f"param_{i}", _axis(lo, hi)) for i, (lo, hi) in enumerate(kernel.bounds)]
    combos = list(product(*axes))
    n_total = len(combos)
    if n_total > max_combinations:
        random.seed(42)
        combos = random.sample(combos, max_combinations)
    candidates = np.array(combos, dtype=np.float64).reshape(len(combos), len(axes))

    # ONE device round trip: metric of the start point and of every candidate, the first-strict-minimum scan,
    # the fit of the winner (noise + 1e-10 jitter, as GaussianProcessRegressor.fit will ask for next) and its
    # fitted mean at the training inputs.  The prints below keep the reference's order.
    res = _gp.search_and_fit(kernel.kernel_id, kernel.with_fixed(start), kernel.with_fixed(candidates), X_train, y_train,
                             noise_variance, noise_variance + 1e-10, **val)
    if use_validation:
        print(f"  Initial (pre-grid-search) {metric_name}: {res['start_metric']:.6e}")
        print(f"  Initial params: {start}, noise: {noise_variance:.3e} (fixed)")
    print(f"  Grid search: {n_total} combinations to evaluate...")
    if n_total > max_combinations:
        print(f"  Too many combinations ({n_total}), randomly sampling {max_combinations}...")

    best_idx, best_metric = res["index"], res["best"]
    best_params = candidates[best_idx].copy() if best_idx >= 0 else None
    # as in the reference: if no candidate beat +inf (all NaN), set_params(None) raises here
    kernel.set_params(best_params)
    print(f"  Grid search complete. Best {metric_name}: {best_metric:.6e}")
    print(f"  Optimal params: {best_params}, noise: {noise_variance:.3e} (fixed)")

    if use_validation:
        # diagnostic only (reference :209-232): training RMSE of the selected model with the fit jitter
        try:
            if res["used_pinv"]:                                   # singular winner: pseudo-inverse fit
                alpha, _, _ = _gp.fit(kernel.kernel_id, kernel.theta(), X_train, y_train, noise_variance + 1e-10, want_kinv=False)
                pred = _gp.gram(kernel.kernel_id, kernel.theta(), X_train) @ alpha
            else:
                pred = res["train_pred"]
            truth = np.asarray(y_train, dtype=float).ravel()
            if y_scaler is not None:
                pred, truth = pred * y_scale + y_mean, truth * y_scale + y_mean
            train_rmse = float(np.sqrt(np.mean((truth - pred) ** 2)))
            print(f"  Training RMSE: {train_rmse:.6e}, Validation RMSE: {best_metric:.6e}")
            if best_metric > train_rmse * 1.5:
                print(f"  WARNING: Validation RMSE is {best_metric / train_rmse:.2f}x higher than training RMSE")
                print("           This may indicate poor generalization or data mismatch")
        except Exception:
            pass
