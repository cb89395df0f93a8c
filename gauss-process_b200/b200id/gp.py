"""Host side of K2/K3/K4 (GP Gram, batched candidate metric, fit, predict, blocked Cholesky).

Mirrors the arithmetic of reference src/gpr/kernels.py, src/gpr/grid_search.py:91-133,186-196 and
src/gpr/gpr_fitting.py:73-100 through the C ABI (include/b200id.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .runtime import F64, Workspace, empty, ptr, require_cuda, stream_ptr, to_device

# create_kernel names (kernels.py:392-407) -> (C ABI id, number of parameters)
KERNEL_IDS = {
    "rbf": (0, 2), "matern12": (1, 2), "matern32": (2, 2), "matern52": (3, 2), "exp": (4, 2),
    "dc": (5, 3), "di": (6, 2), "ss1": (7, 2), "ss2": (8, 2), "sshf": (9, 2), "stable_spline": (10, 2),
    "rq": (11, 3), "tc": (12, 2),
}
MAX_PARAMS = 3
METRIC_NLL, METRIC_VAL_RMSE = 0, 1
FAIL_METRIC = 1e10
PINV_RCOND = 1e-15          # np.linalg.pinv default (gpr_fitting.py:86)
SMEM_NMAX = 160             # largest n the in-shared-memory kernels accept


def _theta_block(thetas) -> np.ndarray:
    th = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    out = np.zeros((th.shape[0], MAX_PARAMS), dtype=np.float64)
    out[:, :th.shape[1]] = th
    return out


def _flat(X) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1))


def gram(kernel_id: int, theta, X1, X2=None) -> np.ndarray:
    dev = require_cuda()
    x1 = to_device(_flat(X1), dev)
    x2 = x1 if X2 is None else to_device(_flat(X2), dev)
    th = to_device(_theta_block(theta)[0], dev)
    K = empty(x1.numel() * x2.numel(), dev)
    _lib.call("b200id_gp_gram", C.c_int(kernel_id), ptr(th), ptr(x1), C.c_int(x1.numel()), ptr(x2),
              C.c_int(x2.numel()), ptr(K), stream_ptr())
    return K.view(x1.numel(), x2.numel()).cpu().numpy()


def batch_eval_device(kernel_id: int, kernel_ids_d: Optional[torch.Tensor], theta_d: torch.Tensor,
                      X_d: torch.Tensor, y_d: torch.Tensor, noise: float, metric: int = METRIC_NLL,
                      Xv_d: Optional[torch.Tensor] = None, yv_d: Optional[torch.Tensor] = None,
                      y_mean: float = 0.0, y_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """metric[C] (device) for theta_d [C, 3]; one CTA per candidate."""
    dev = theta_d.device
    c, n = theta_d.shape[0], X_d.numel()
    out = empty(c, dev) if out is None else out
    nv = 0 if Xv_d is None else Xv_d.numel()
    ws = Workspace.get("gp", _lib.load().b200id_gp_batch_workspace_bytes(c, n), dev)
    _lib.call("b200id_gp_batch_eval", C.c_int(kernel_id), ptr(kernel_ids_d), ptr(theta_d), C.c_int64(c), ptr(X_d),
              ptr(y_d), C.c_int(n), C.c_double(noise), C.c_int(metric), ptr(Xv_d), ptr(yv_d), C.c_int(nv),
              C.c_double(y_mean), C.c_double(y_scale), ptr(out), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out


def first_strict_min_device(metric_d: torch.Tensor):
    """(best[1], idx[1]) device tensors: first strict minimum in index order, NaNs skipped."""
    dev = metric_d.device
    best = empty(1, dev)
    idx = torch.empty(1, dtype=torch.int64, device=dev)
    ws = Workspace.get("gp", 256, dev)
    _lib.call("b200id_first_strict_min", ptr(metric_d), C.c_int64(metric_d.numel()), ptr(best), ptr(idx), ptr(ws),
              C.c_size_t(ws.numel()), stream_ptr())
    return best, idx


def batch_eval(kernel, thetas, X, y, noise: float, Xv=None, yv=None, y_mean: Optional[float] = None,
               y_scale: Optional[float] = None, kernel_ids: Optional[Sequence[int]] = None) -> np.ndarray:
    """Candidate metrics on the host: NLL, or validation RMSE when (Xv, yv) are given."""
    dev = require_cuda()
    kid = KERNEL_IDS[kernel][0] if isinstance(kernel, str) else int(kernel)
    th = to_device(_theta_block(thetas), dev)
    ids = None if kernel_ids is None else torch.as_tensor(np.asarray(kernel_ids, dtype=np.int32), device=dev)
    use_val = Xv is not None and yv is not None
    m = batch_eval_device(
        kid, ids, th, to_device(_flat(X), dev), to_device(_flat(y), dev), float(noise),
        METRIC_VAL_RMSE if use_val else METRIC_NLL,
        to_device(_flat(Xv), dev) if use_val else None, to_device(_flat(yv), dev) if use_val else None,
        0.0 if y_mean is None else float(y_mean), 1.0 if y_scale is None else float(y_scale))
    return m.cpu().numpy()


def select(kernel, thetas, X, y, noise: float, **kw) -> Tuple[int, float, np.ndarray]:
    """(index, metric, all metrics) of the first strict minimum over the candidate list."""
    dev = require_cuda()
    kid = KERNEL_IDS[kernel][0] if isinstance(kernel, str) else int(kernel)
    th = to_device(_theta_block(thetas), dev)
    Xv, yv = kw.get("Xv"), kw.get("yv")
    use_val = Xv is not None and yv is not None
    ym, ys = kw.get("y_mean"), kw.get("y_scale")
    m = batch_eval_device(
        kid, None, th, to_device(_flat(X), dev), to_device(_flat(y), dev), float(noise),
        METRIC_VAL_RMSE if use_val else METRIC_NLL,
        to_device(_flat(Xv), dev) if use_val else None, to_device(_flat(yv), dev) if use_val else None,
        0.0 if ym is None else float(ym), 1.0 if ys is None else float(ys))
    best, idx = first_strict_min_device(m)
    return int(idx.item()), float(best.item()), m.cpu().numpy()


def fit(kernel_id: int, theta, X, y, diag_add: float, want_kinv: bool = True):
    """(alpha[n], K_inv[n,n] or None, used_pinv) -- Cholesky path, pseudo-inverse path when a pivot is <= 0."""
    dev = require_cuda()
    x, yy = to_device(_flat(X), dev), to_device(_flat(y), dev)
    n = x.numel()
    th = to_device(_theta_block(theta)[0], dev)
    want_kinv = want_kinv or (127 < n <= SMEM_NMAX)
    alpha, Kinv = empty(n, dev), (empty(n * n, dev) if want_kinv else None)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = _lib.load().b200id_gp_fit_workspace_bytes(n)
    ws = Workspace.get("gpfit", nbytes, dev)
    _lib.call("b200id_gp_fit", C.c_int(kernel_id), ptr(th), ptr(x), ptr(yy), C.c_int(n), C.c_double(diag_add),
              ptr(alpha), ptr(Kinv), ptr(info), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    used_pinv = int(info.item()) != 0
    if used_pinv:
        if Kinv is None:
            Kinv = empty(n * n, dev)
        _lib.call("b200id_gp_fit_pinv", C.c_int(kernel_id), ptr(th), ptr(x), ptr(yy), C.c_int(n), C.c_double(diag_add),
                  C.c_double(PINV_RCOND), ptr(alpha), ptr(Kinv), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return alpha.cpu().numpy(), (Kinv.view(n, n).cpu().numpy() if Kinv is not None else None), used_pinv


def predict(kernel_id: int, theta, X_train, alpha, K_inv, Xs, return_std: bool = False):
    dev = require_cuda()
    x, xs = to_device(_flat(X_train), dev), to_device(_flat(Xs), dev)
    n, m = x.numel(), xs.numel()
    th = to_device(_theta_block(theta)[0], dev)
    a = to_device(_flat(alpha), dev)
    kinv = to_device(_flat(K_inv), dev) if return_std else None
    mean = empty(m, dev)
    std = empty(m, dev) if return_std else None
    _lib.call("b200id_gp_predict", C.c_int(kernel_id), ptr(th), ptr(x), C.c_int(n), ptr(a), ptr(kinv), ptr(xs),
              C.c_int(m), ptr(mean), ptr(std), stream_ptr())
    if return_std:
        return mean.cpu().numpy(), std.cpu().numpy()
    return mean.cpu().numpy()


def cholesky_blocked_device(A_d: torch.Tensor) -> int:
    """In-place lower Cholesky of a square row-major device matrix; returns info (0 = ok)."""
    n = A_d.shape[0]
    dev = A_d.device
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = _lib.load().b200id_chol_blocked_workspace_bytes(n)
    ws = Workspace.get("cholb", nbytes, dev)
    _lib.call("b200id_chol_blocked", ptr(A_d), C.c_int(n), C.c_int(A_d.stride(0)), ptr(info), ptr(ws),
              C.c_size_t(ws.numel()), stream_ptr())
    return int(info.item())
