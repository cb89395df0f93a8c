"""CPU restatement of the reference's kernel-regularised FIR identification (test infrastructure only).

Follows /root/reference/src/fir_model/kernel_regularized.py line by line in NumPy / SciPy: the Toeplitz regressor and
its summaries (:84-92, :337-340), the DC / second-order stable-spline / simulation-induced covariances (:99-181), the
unconstrained parametrisation (:192-217), the marginal likelihood through the L x L identities (:224-261) and the
posterior mean (:264-291).  Pinned by tests/golden/kr_fir.npz (oracle/make_golden.py::golden_kr runs the reference
itself).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from scipy.linalg import cholesky, cho_solve

KDIM = {"dc": 3, "ss": 2, "si": 4}


def build_regressor(u: np.ndarray, L: int) -> np.ndarray:
    """Rows phi[n] = [u[n], u[n-1], ..., u[n-L+1]], zeros before the record (:84-92)."""
    N = len(u)
    u_pad = np.concatenate([np.zeros(L - 1), np.asarray(u, dtype=float)])
    return sliding_window_view(u_pad, window_shape=L)[:N, ::-1]


def summaries(u: np.ndarray, y: np.ndarray, L: int):
    """G = U^T U, b = U^T y, yy = y^T y (:337-340)."""
    U = build_regressor(u, L)
    return U.T @ U, U.T @ y, float(y @ y)


def kernel_dc(L: int, c: float, lam: float, rho: float) -> np.ndarray:
    """k(i, j) = c lam^((i+j)/2) rho^|i-j| (:99-112)."""
    i = np.arange(L)[:, None]
    j = np.arange(L)[None, :]
    K = (lam ** ((i + j) / 2.0)) * np.power(np.abs(rho), np.abs(i - j))
    if rho < 0:
        K = K * ((-1.0) ** np.abs(i - j))
    return c * K


def kernel_ss2(L: int, c: float, lam: float) -> np.ndarray:
    """K = c (C H)(C H)^T, H[i, j] = lam^(i-j) lower triangular, C the cumulative-sum matrix (:115-127)."""
    idx = np.arange(L)
    with np.errstate(all="ignore"):
        H = np.tril(lam ** (idx[:, None] - idx[None, :]).astype(float))
    H[~np.isfinite(H)] = 0.0
    CH = np.tril(np.ones((L, L))) @ H
    return c * (CH @ CH.T)


def kernel_si(L: int, r: float, omega: float, q: float, barlam: float) -> np.ndarray:
    """Simulation-induced kernel: z0 term q C A^t (A^s)^T C^T plus the process-noise sum (:130-181)."""
    A = r * np.array([[math.cos(omega), -math.sin(omega)], [math.sin(omega), math.cos(omega)]])
    P = np.eye(2)
    v = np.zeros((L, 2))
    for n in range(L):
        v[n] = P[0]
        P = P @ A
    g = v[:, 0].copy()
    K = q * (v @ v.T)
    b2 = barlam ** np.arange(L)
    for t in range(L):
        for s in range(L):
            m = min(t, s)
            if m:
                k = np.arange(m)
                K[t, s] += float(np.sum(b2[k] * g[t - 1 - k] * g[s - 1 - k]))
    return K


def _sigmoid(x: float) -> float:
    return 1.0 / (1.0 + math.exp(-x))


def build_kernel(L: int, kernel: str, theta) -> np.ndarray:
    """Unconstrained theta -> covariance (:192-217)."""
    k = kernel.lower()
    if k == "dc":
        return kernel_dc(L, math.exp(theta[0]), _sigmoid(theta[1]), math.tanh(theta[2]))
    if k == "ss":
        return kernel_ss2(L, math.exp(theta[0]), min(max(_sigmoid(theta[1]), 1e-6), 1 - 1e-6))
    if k == "si":
        r = min(max(_sigmoid(theta[0]), 1e-4), 1 - 1e-6)
        omega = min(max(math.pi * _sigmoid(theta[1]), 1e-3), math.pi - 1e-3)
        barlam = min(max(_sigmoid(theta[3]), 1e-6), 1 - 1e-6)
        return kernel_si(L, r, omega, math.exp(theta[2]), barlam)
    raise ValueError(f"Unknown kernel: {kernel}")


def nll(theta_all, G, b, yy: float, N: int, L: int, kernel: str, jitter: float = 1e-10) -> float:
    """Negative log marginal likelihood up to a constant (:224-261)."""
    kdim = KDIM[kernel.lower()]
    sig2 = max(math.exp(theta_all[kdim]), 1e-30)
    K = build_kernel(L, kernel, theta_all[:kdim])
    Lk = cholesky(K + jitter * np.eye(L), lower=True, check_finite=False)
    S = np.eye(L) + (Lk.T @ (G @ Lk)) / sig2
    try:
        Ls = cholesky(S + jitter * np.eye(L), lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        Ls = cholesky(S + 1e-6 * np.eye(L), lower=True, check_finite=False)
    w = Lk.T @ b
    x = cho_solve((Ls, True), w, check_finite=False)
    quad = (yy - (w @ x) / sig2) / sig2
    return float(0.5 * (quad + N * math.log(sig2) + 2.0 * np.sum(np.log(np.diag(Ls)))))


def posterior_mean_g(theta_all, G, b, L: int, kernel: str):
    """(g_hat, sigma^2, K) (:264-291)."""
    kdim = KDIM[kernel.lower()]
    sig2 = math.exp(theta_all[kdim])
    K = build_kernel(L, kernel, theta_all[:kdim])
    Lk = cholesky(K + 1e-10 * np.eye(L), lower=True, check_finite=False)
    S = np.eye(L) + (Lk.T @ (G @ Lk)) / max(sig2, 1e-30)
    Ls = cholesky(S + 1e-10 * np.eye(L), lower=True, check_finite=False)
    w = Lk.T @ b
    x = cho_solve((Ls, True), w, check_finite=False)
    return ((Lk @ x) / max(sig2, 1e-30)).reshape(-1), sig2, K
