"""Oracle: GP kernels, candidate metric, grid search, fit/predict (rows a-6 .. a-8).

Test infrastructure only.  NumPy restatement of ``src/gpr/kernels.py``,
``src/gpr/gpr_fitting.py`` and ``src/gpr/grid_search.py`` of the reference.
"""
from __future__ import annotations

import itertools
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# name -> (kernel id shared with the CUDA side, parameter names in get_params() order, bounds)
WIDE = (1e-3, 1e3)
KERNELS: Dict[str, Tuple[int, Tuple[str, ...], Tuple[Tuple[float, float], ...]]] = {
    "rbf":           (0, ("length_scale", "variance"), (WIDE, WIDE)),                 # kernels.py:90-110
    "matern12":      (1, ("length_scale", "variance"), (WIDE, WIDE)),                 # kernels.py:113-147, nu=0.5
    "matern32":      (2, ("length_scale", "variance"), (WIDE, WIDE)),                 # nu=1.5
    "matern52":      (3, ("length_scale", "variance"), (WIDE, WIDE)),                 # nu=2.5
    "exp":           (4, ("omega", "variance"), (WIDE, WIDE)),                        # kernels.py:175-198
    "dc":            (5, ("alpha", "beta", "rho"), ((1e-3, 0.999), WIDE, (-0.999, 0.999))),  # kernels.py:227-250
    "di":            (6, ("beta", "alpha"), (WIDE, (1e-3, 0.999))),                   # kernels.py:253-278
    "ss1":           (7, ("beta", "variance"), (WIDE, WIDE)),                         # kernels.py:282-301
    "ss2":           (8, ("beta", "variance"), (WIDE, WIDE)),                         # kernels.py:304-329
    "sshf":          (9, ("beta", "variance"), (WIDE, WIDE)),                         # kernels.py:332-355
    "stable_spline": (10, ("beta", "variance"), (WIDE, WIDE)),                        # kernels.py:358-382
    "rq":            (11, ("length_scale", "variance", "alpha"), (WIDE, WIDE, WIDE)),  # kernels.py:150-171
    "tc":            (12, ("omega", "variance"), (WIDE, WIDE)),                       # kernels.py:201-224
    # general-nu Matern: theta = [length_scale, variance, nu]; nu is a fixed shape parameter (kernels.py:131-135)
    "matern_nu":     (13, ("length_scale", "variance"), (WIDE, WIDE)),
}
BASELINE_KERNELS = ["dc", "di", "exp", "matern12", "matern32", "matern52",
                    "rbf", "ss1", "ss2", "sshf", "stable_spline"]        # run_baseline.py:37-41
FAIL_METRIC = 1e10                                                      # grid_search.py:132-133


def _flat(X) -> np.ndarray:
    return np.asarray(X, dtype=np.float64).reshape(-1)


def gram(name: str, theta: Sequence[float], X1, X2=None) -> np.ndarray:
    """K(X1, X2) for 1-D inputs; formulas and operation order follow kernels.py:90-382."""
    a = _flat(X1)
    b = a if X2 is None else _flat(X2)
    A, B = a[:, None], b[None, :]
    th = [float(v) for v in theta]
    if name == "rbf":                                   # :98-100 (cdist sqeuclidean == (x-x')**2 in 1-D)
        d2 = (A - B) ** 2
        return th[1] * np.exp(-0.5 * d2 / (th[0] ** 2))
    if name in ("matern12", "matern32", "matern52"):    # :121-136 (cdist euclidean == |x-x'| in 1-D)
        d = np.sqrt((A - B) ** 2) / th[0]
        if name == "matern12":
            K = np.exp(-d)
        elif name == "matern32":
            K = (1.0 + np.sqrt(3.0) * d) * np.exp(-np.sqrt(3.0) * d)
        else:
            K = (1.0 + np.sqrt(5.0) * d + 5.0 / 3.0 * d ** 2) * np.exp(-np.sqrt(5.0) * d)
        K *= th[1]
        return K
    if name == "matern_nu":                             # :121,131-136 (scipy.special.kv branch)
        from scipy.special import gamma, kv
        d = np.sqrt((A - B) ** 2) / th[0]
        nu = th[2]
        K = d ** nu
        K *= kv(nu, np.sqrt(2.0 * nu) * d)
        K *= 2.0 ** (1.0 - nu) / gamma(nu)
        K[d == 0] = 1.0
        K *= th[1]
        return K
    if name == "rq":                                    # :158-161
        d2 = (A - B) ** 2
        return th[1] * (1.0 + d2 / (2.0 * th[2] * th[0] ** 2)) ** (-th[2])
    if name in ("exp", "tc"):                           # :183-188, :209-214
        Ha, Hb = (a >= 0.0).astype(float), (b >= 0.0).astype(float)
        arg = (A + B) if name == "exp" else np.maximum.outer(a, b)
        K = np.exp(-th[0] * arg)
        K *= Ha[:, None] * Hb[None, :]
        return th[1] * K
    if name == "dc":                                    # :235-240
        i = np.rint(a).astype(int)
        j = np.rint(b).astype(int)
        half_sum = (i[:, None] + j[None, :]) / 2.0
        gap = np.abs(i[:, None] - j[None, :])
        return th[1] * ((th[0] ** half_sum) * (th[2] ** gap))
    if name == "di":                                    # :261-268
        i = np.rint(a).astype(int)
        j = np.rint(b).astype(int)
        K = np.zeros((i.size, j.size))
        r, c = np.where(i[:, None] == j[None, :])
        K[r, c] = th[0] * (th[1] ** i[r])
        return K
    if name == "ss1":                                   # :290-291
        return th[1] * np.exp(-th[0] * np.minimum.outer(a, b))
    if name == "ss2":                                   # :312-319
        mx = np.maximum.outer(a, b)
        return th[1] * (0.5 * np.exp(-th[0] * ((A + B) + mx)) - (1.0 / 6.0) * np.exp(-3.0 * th[0] * mx))
    if name == "sshf":                                  # :340-345
        ia, ib = np.rint(a).astype(int), np.rint(b).astype(int)
        sign = np.power(-1.0, ia[:, None] + ib[None, :])
        return th[1] * sign * np.maximum(np.exp(-th[0] * A), np.exp(-th[0] * B))
    if name == "stable_spline":                         # :366-372
        ea, eb = np.exp(-th[0] * a), np.exp(-th[0] * b)
        lo, hi = np.minimum.outer(ea, eb), np.maximum.outer(ea, eb)
        return th[1] * (0.5 * lo ** 2 * (hi - lo / 3.0))
    raise ValueError(f"Unknown kernel type: {name}. Available: {list(KERNELS)}")


def _alpha_from_chol(K: np.ndarray, y: np.ndarray):
    # gpr_fitting.py:80-82 / grid_search.py:103-105: cholesky + two general solves
    L = np.linalg.cholesky(K)
    z = np.linalg.solve(L, y)
    return L, np.linalg.solve(L.T, z)


def nll_value(L: np.ndarray, alpha: np.ndarray, y: np.ndarray) -> float:
    """0.5 y.alpha + sum log L_ii + 0.5 n log 2pi  (gpr_fitting.py:117-119, grid_search.py:128-130)."""
    v = 0.5 * (y @ alpha)
    v += np.sum(np.log(np.diag(L)))
    v += 0.5 * len(y) * np.log(2 * np.pi)
    return float(v)


def candidate_metric(name: str, theta, X, y, noise: float, Xv=None, yv=None,
                     y_mean: Optional[float] = None, y_scale: Optional[float] = None) -> float:
    """Metric of one grid candidate; follows grid_search.py:91-133.

    K + noise*I (no jitter), Cholesky, alpha; validation RMSE on the de-standardised
    prediction when (Xv, yv) are given (y_mean/y_scale = StandardScaler mean_/scale_, None = no
    scaler), else NLL; LinAlgError -> 1e10.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    K = gram(name, theta, X)
    K += noise * np.eye(K.shape[0])
    try:
        L, alpha = _alpha_from_chol(K, y)
    except np.linalg.LinAlgError:
        return FAIL_METRIC
    if Xv is not None and yv is not None:
        pred = gram(name, theta, Xv, X) @ alpha
        if y_scale is not None:
            pred = pred * y_scale + y_mean          # StandardScaler.inverse_transform
        return float(np.sqrt(np.mean((np.asarray(yv, dtype=np.float64) - pred) ** 2)))
    return nll_value(L, alpha, y)


def default_grids(name: str, points: int = 30) -> List[np.ndarray]:
    """30-point grid per parameter, log-spaced when lo > 0 and hi/lo > 100 (grid_search.py:21-45)."""
    grids = []
    for lo, hi in KERNELS[name][2]:
        if lo > 0 and hi / lo > 100:
            grids.append(np.logspace(np.log10(lo), np.log10(hi), points))
        else:
            grids.append(np.linspace(lo, hi, points))
    return grids


def grid_candidates(name: str, max_combinations: int = 5000, grids=None) -> np.ndarray:
    """Candidate list [C, p] in the reference's scan order (grid_search.py:173-183)."""
    grids = default_grids(name) if grids is None else grids
    combos = list(itertools.product(*grids))
    if len(combos) > max_combinations:
        random.seed(42)
        combos = random.sample(combos, max_combinations)
    return np.array(combos, dtype=np.float64).reshape(len(combos), len(grids))


def first_strict_min(metrics: np.ndarray) -> Tuple[int, float]:
    """Sequential ``metric < best`` scan starting from +inf (grid_search.py:186-196).

    NaNs never win; returns (-1, inf) when nothing beats +inf.
    """
    best_i, best = -1, np.inf
    for i, m in enumerate(metrics):
        if m < best:
            best, best_i = float(m), i
    return best_i, best


def grid_search(name: str, X, y, noise: float, max_combinations: int = 5000, Xv=None, yv=None,
                y_mean=None, y_scale=None, grids=None):
    """(best_theta, best_metric, metrics[C], candidates[C,p]); grid_search.py:48-203."""
    cands = grid_candidates(name, max_combinations, grids)
    metrics = np.array([candidate_metric(name, th, X, y, noise, Xv, yv, y_mean, y_scale) for th in cands])
    i, best = first_strict_min(metrics)
    return (cands[i].copy() if i >= 0 else None), best, metrics, cands


def fit(name: str, theta, X, y, noise: float):
    """Final fit: K + (noise + 1e-10) I -> (alpha, K_inv, used_pinv); gpr_fitting.py:73-87."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    K = gram(name, theta, X)
    K += (noise + 1e-10) * np.eye(K.shape[0])
    try:
        _, alpha = _alpha_from_chol(K, y)
        return alpha, np.linalg.inv(K), False
    except np.linalg.LinAlgError:
        K_inv = np.linalg.pinv(K)
        return K_inv @ y, K_inv, True


def predict(name: str, theta, X_train, alpha, K_inv, Xs, return_std: bool = False):
    """mean = K* alpha; std = sqrt(max(diag K** - sum((K* K_inv) o K*), 0)); gpr_fitting.py:89-100."""
    Ks = gram(name, theta, Xs, X_train)
    mean = Ks @ alpha
    if not return_std:
        return mean
    var = np.diag(gram(name, theta, Xs)) - np.sum((Ks @ K_inv) * Ks, axis=1)
    return mean, np.sqrt(np.maximum(var, 0))


def restart_population(name: str, n: int, noise: float, seed: int = 42) -> np.ndarray:
    """n initial points drawn with the L-BFGS-B restart rule (gpr_fitting.py:132-141).

    Per restart and per bound (kernel bounds + the log-noise bound): log-uniform when lo > 0 and
    hi/lo > 100, else uniform, drawn in that order from np.random (seeded); the noise slot is
    then discarded (it is overwritten with log(noise) in the reference).  Used for BASELINE
    config 3 (11 kernels x 4096 restarts).  Returns [n, p].
    """
    rs = np.random.RandomState(seed)
    bounds = list(KERNELS[name][2]) + [(np.log(1e-10), np.log(1e-1))]
    out = np.empty((n, len(bounds) - 1))
    for r in range(n):
        row = []
        for lo, hi in bounds:
            if lo > 0 and hi / lo > 100:
                row.append(np.exp(rs.uniform(np.log(lo), np.log(hi))))
            else:
                row.append(rs.uniform(lo, hi))
        out[r] = row[:-1]
    return out


def standardize(v: np.ndarray):
    """sklearn StandardScaler on one column: (z, mean_, scale_) with population std; gp_pipeline.py:130-132."""
    v = np.asarray(v, dtype=np.float64)
    mean = float(np.mean(v))
    scale = float(np.sqrt(np.mean((v - mean) ** 2)))
    if scale == 0.0:
        scale = 1.0
    return (v - mean) / scale, mean, scale
