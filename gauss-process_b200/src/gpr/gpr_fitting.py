"""GaussianProcessRegressor with the reference's surface; fit / predict / NLL on the GPU.

Mirror of reference src/gpr/gpr_fitting.py:20-155.  ``fit`` = optional hyperparameter optimisation
(batched grid search, or SciPy L-BFGS-B driving device NLL evaluations), then the final factorisation
of K + (noise + 1e-10) I on the device (b200id_gp_fit; pseudo-inverse branch b200id_gp_fit_pinv when a
pivot is non-positive, reference :84-87).  ``predict`` = b200id_gp_predict.  Attributes kept:
``kernel, noise_variance, X_train, y_train, alpha, K_inv, X_scaler, y_scaler``.
"""
from typing import Dict, Tuple, Union

import numpy as np
from scipy.optimize import minimize

from b200id import gp as _gp
from src.gpr.kernels import Kernel

FIT_JITTER = 1e-10                      # reference :76
LOG_NOISE_BOUNDS = (np.log(1e-10), np.log(1e-1))   # reference :130


class GaussianProcessRegressor:
    """Gaussian Process Regressor with extensible kernel support (device-backed)."""

    def __init__(self, kernel: Kernel, noise_variance: float = 0.0):
        self.kernel = kernel
        self.noise_variance = noise_variance
        self.X_train = None
        self.y_train = None
        self._K_inv = None
        self._fit_args = None
        self.alpha = None
        self.X_scaler = None
        self.y_scaler = None

    def fit(self, X: np.ndarray, y: np.ndarray, optimize: bool = True, n_restarts: int = 3,
            use_grid_search: bool = False, param_grids: Dict = None, max_grid_combinations: int = 5000,
            validation_X: np.ndarray = None, validation_y: np.ndarray = None, y_scaler: object = None):
        """Fit the GP; arguments as in the reference (:33-87)."""
        self.X_train = X.copy()
        self.y_train = y.copy()
        if optimize:
            if use_grid_search:
                from src.gpr.grid_search import grid_search_hyperparameters, get_default_param_grids
                if param_grids is None:
                    param_grids = get_default_param_grids(self.kernel)
                grid_search_hyperparameters(
                    kernel=self.kernel, X_train=self.X_train, y_train=self.y_train,
                    noise_variance=self.noise_variance, param_grids=param_grids,
                    max_combinations=max_grid_combinations, validation_X=validation_X,
                    validation_y=validation_y, y_scaler=y_scaler)
            else:
                self._optimize_hyperparameters(n_restarts)
        self._fit_args = (self.kernel.kernel_id, self.kernel.theta().copy(), self.noise_variance + FIT_JITTER)
        self.alpha, self._K_inv, _ = _gp.fit(*self._fit_args[:2], self.X_train, self.y_train, self._fit_args[2],
                                             want_kinv=False)

    @property
    def K_inv(self):
        """(K + (noise + 1e-10) I)^-1 of the last fit (reference :78-87).  The posterior mean needs only
        ``alpha``; the n x n inverse is formed on first use (``predict(return_std=True)`` or a direct read)
        from the same factorisation inputs, so its values are those an eager fit would have stored."""
        if self._K_inv is None and self._fit_args is not None:
            kid, theta, diag_add = self._fit_args
            _, self._K_inv, _ = _gp.fit(kid, theta, self.X_train, self.y_train, diag_add, want_kinv=True)
        return self._K_inv

    @K_inv.setter
    def K_inv(self, value):
        self._K_inv = value

    def predict(self, X: np.ndarray, return_std: bool = False) -> Union[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
        """Posterior mean (and std = sqrt(max(k** - sum((K* K_inv) o K*), 0))) at X (reference :89-100)."""
        return _gp.predict(self.kernel.kernel_id, self.kernel.theta(), self.X_train, self.alpha,
                           self.K_inv if return_std else None, X, return_std)

    def _optimize_hyperparameters(self, n_restarts: int):
        """L-BFGS-B on the negative log marginal likelihood with random restarts (reference :102-155).

        SciPy stays the host driver (its own forward-difference gradients, same bounds, same restart rule and
        np.random stream); every NLL it asks for is a device evaluation.  The p + 1 forward-difference points
        of one gradient are independent, so they go to the device as ONE batched launch with a per-candidate
        noise variance (``b200id_gp_batch_eval_noise``) through SciPy's ``workers`` map hook; a SciPy without
        that hook evaluates them one by one, as the reference does."""
        kid = self.kernel.kernel_id

        def nll_many(points):
            pts = np.asarray(points, dtype=np.float64).reshape(len(points), -1)
            return _gp.batch_eval(kid, self.kernel.with_fixed(pts[:, :-1]), self.X_train, self.y_train, np.exp(pts[:, -1]))

        def nll(params):
            self.kernel.set_params(params[:-1])
            self.noise_variance = np.exp(params[-1])
            return float(nll_many([params])[0])

        def batched_map(_fun, points):
            points = list(points)
            if not points:
                return []
            vals = [float(v) for v in nll_many(points)]
            # the reference's objective mutates the model on EVERY evaluation (reference :105-106), finite-difference
            # points included, and the next restart starts from log(self.noise_variance) (:143): leave the model in
            # the state the sequential evaluation would have left it in, i.e. that of the last point
            last = np.asarray(points[-1], dtype=np.float64)
            self.kernel.set_params(last[:-1])
            self.noise_variance = np.exp(last[-1])
            return vals

        options = {"workers": batched_map} if _LBFGSB_TAKES_WORKERS else {}
        bounds = self.kernel.bounds + [LOG_NOISE_BOUNDS]
        best_x, best_f = None, np.inf
        for restart in range(n_restarts):
            start = []
            for low, high in bounds:
                if low > 0 and high / low > 100:
                    start.append(np.exp(np.random.uniform(np.log(low), np.log(high))))
                else:
                    start.append(np.random.uniform(low, high))
            start = np.array(start)
            start[-1] = np.log(self.noise_variance)
            if restart == 0:
                start[:-1] = self.kernel.get_params()
            res = minimize(nll, start, bounds=bounds, method="L-BFGS-B", options=options)
            if res.fun < best_f:
                best_f, best_x = res.fun, res.x
        self.kernel.set_params(best_x[:-1])
        self.noise_variance = np.exp(best_x[-1])


def _lbfgsb_takes_workers() -> bool:
    """Does this SciPy's L-BFGS-B forward a ``workers`` map to its finite differences?  (SciPy >= 1.16.)"""
    try:
        import inspect
        from scipy.optimize import _lbfgsb_py
        return "workers" in inspect.signature(_lbfgsb_py._minimize_lbfgsb).parameters
    except Exception:
        return False


_LBFGSB_TAKES_WORKERS = _lbfgsb_takes_workers()
