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
    "rq": (11, 3), "tc": (12, 2), "matern_nu": (13, 2),      # matern_nu: theta[2] = nu, fixed (not optimised)
}
MAX_PARAMS = 3
METRIC_NLL, METRIC_VAL_RMSE = 0, 1
FAIL_METRIC = 1e10
PINV_RCOND = 1e-15          # np.linalg.pinv default (gpr_fitting.py:86)
SMEM_NMAX = 160             # largest n the in-shared-memory kernels accept
TILED_NMAX = 127            # largest n of the register-tiled candidate kernel (126 with two targets)
# shared covariance shapes (b200id_gp_batch_eval_shared): used when a list has at least SHARED_MIN_GAIN candidates per
# distinct shape and the shape tables stay below SHARED_MAX_TABLE_BYTES; B200ID_GP_SHARED=0 switches it off
SHARED_MIN_GAIN = 2
SHARED_MAX_TABLE_BYTES = 256 << 20
AMPLITUDE_SLOT = 1


def _theta_block(thetas) -> np.ndarray:
    th = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    out = np.zeros((th.shape[0], MAX_PARAMS), dtype=np.float64)
    out[:, :th.shape[1]] = th
    return out


def shape_groups(kernel_id: int, thetas) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """Candidates that differ only in the amplitude theta[1] share the shape of their covariance (every kernel but DI
    multiplies by theta[1] last, kernels.py:98-372), and a grid-search list is a product of 1-D grids
    (grid_search.py:173).  Returns (shape_of [C] int32, shape_theta [S, 3]) when sharing pays, else None."""
    import os
    if os.environ.get("B200ID_GP_SHARED", "1") == "0" or kernel_id == KERNEL_IDS["di"][0]:
        return None
    th = _theta_block(thetas)
    if th.shape[0] < 2 * SHARED_MIN_GAIN:
        return None
    key = th.copy()
    key[:, AMPLITUDE_SLOT] = 0.0
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    if first.size * SHARED_MIN_GAIN > th.shape[0]:
        return None
    return np.ascontiguousarray(inverse.reshape(-1), dtype=np.int32), np.ascontiguousarray(th[first])


def _flat(X) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1))


def _points(X, what: str = "X") -> np.ndarray:
    """1-D GP inputs as a flat vector.  The device Gram templates are written for scalar inputs -- what the
    pipeline feeds them (log10 omega, gp_pipeline.py:281-282); the reference kernels built on ``cdist`` would also
    accept (N, D > 1), which is refused here instead of being silently read as N*D points."""
    a = np.asarray(X, dtype=np.float64)
    if a.ndim > 2 or (a.ndim == 2 and a.shape[1] != 1):
        raise ValueError(f"{what} must have shape (n,) or (n, 1): the b200id GP kernels take 1-D inputs, got {a.shape}")
    return np.ascontiguousarray(a.reshape(-1))


def _targets(y, n: int, what: str = "y") -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
    if a.size != n:
        raise ValueError(f"{what} has {a.size} values for {n} input points")
    return a


def gram(kernel_id: int, theta, X1, X2=None) -> np.ndarray:
    dev = require_cuda()
    x1 = to_device(_points(X1, "X1"), dev)
    x2 = x1 if X2 is None else to_device(_points(X2, "X2"), dev)
    th = to_device(_theta_block(theta)[0], dev)
    K = empty(x1.numel() * x2.numel(), dev)
    _lib.call("b200id_gp_gram", C.c_int(kernel_id), ptr(th), ptr(x1), C.c_int(x1.numel()), ptr(x2),
              C.c_int(x2.numel()), ptr(K), stream_ptr())
    return K.view(x1.numel(), x2.numel()).cpu().numpy()


def batch_eval_device(kernel_id: int, kernel_ids_d: Optional[torch.Tensor], theta_d: torch.Tensor,
                      X_d: torch.Tensor, y_d: torch.Tensor, noise, metric: int = METRIC_NLL,
                      Xv_d: Optional[torch.Tensor] = None, yv_d: Optional[torch.Tensor] = None,
                      y_mean: float = 0.0, y_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """metric[C] (device) for theta_d [C, 3]; one CTA per candidate.  ``noise`` = one float for all
    candidates, or a device tensor [C] with one noise variance per candidate."""
    dev = theta_d.device
    c, n = theta_d.shape[0], X_d.numel()
    out = empty(c, dev) if out is None else out
    nv = 0 if Xv_d is None else Xv_d.numel()
    ws = Workspace.get("gp", _lib.load().b200id_gp_batch_workspace_bytes(c, n), dev)
    if isinstance(noise, torch.Tensor):
        assert noise.numel() == c and noise.dtype == F64
        _lib.call("b200id_gp_batch_eval_noise", C.c_int(kernel_id), ptr(kernel_ids_d), ptr(theta_d), ptr(noise), C.c_int64(c),
                  ptr(X_d), ptr(y_d), C.c_int(n), C.c_int(metric), ptr(Xv_d), ptr(yv_d), C.c_int(nv),
                  C.c_double(y_mean), C.c_double(y_scale), ptr(out), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
        return out
    _lib.call("b200id_gp_batch_eval", C.c_int(kernel_id), ptr(kernel_ids_d), ptr(theta_d), C.c_int64(c), ptr(X_d),
              ptr(y_d), C.c_int(n), C.c_double(noise), C.c_int(metric), ptr(Xv_d), ptr(yv_d), C.c_int(nv),
              C.c_double(y_mean), C.c_double(y_scale), ptr(out), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out


def batch_eval_pair_device(kernel_id: int, kernel_ids_d: Optional[torch.Tensor], theta_d: torch.Tensor,
                           X_d: torch.Tensor, y_a: torch.Tensor, y_b: torch.Tensor, noise: float, metric: int = METRIC_NLL,
                           Xv_d: Optional[torch.Tensor] = None, yv_a: Optional[torch.Tensor] = None,
                           yv_b: Optional[torch.Tensor] = None, scale_a=(0.0, 1.0), scale_b=(0.0, 1.0)):
    """(metric_a[C], metric_b[C]) for two targets on the same inputs / candidates / noise: ONE factorisation per
    candidate serves both right-hand sides (b200id_gp_batch_eval_pair).  scale_* = (y_mean, y_scale)."""
    dev = theta_d.device
    c, n = theta_d.shape[0], X_d.numel()
    out = empty(2 * c, dev)
    nv = 0 if Xv_d is None else Xv_d.numel()
    ws = Workspace.get("gp", _lib.load().b200id_gp_batch_workspace_bytes(c, n), dev)
    _lib.call("b200id_gp_batch_eval_pair", C.c_int(kernel_id), ptr(kernel_ids_d), ptr(theta_d), C.c_int64(c), ptr(X_d),
              ptr(y_a), ptr(y_b), C.c_int(n), C.c_double(noise), C.c_int(metric), ptr(Xv_d), ptr(yv_a), ptr(yv_b), C.c_int(nv),
              C.c_double(scale_a[0]), C.c_double(scale_a[1]), C.c_double(scale_b[0]), C.c_double(scale_b[1]),
              ptr(out[:c]), ptr(out[c:]), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return out[:c], out[c:]


def shared_fits(n_shapes: int, n: int, nv: int, two_targets: bool) -> bool:
    """Can b200id_gp_batch_eval_shared take this problem (kernel size limit, table size cap)?"""
    if n > (TILED_NMAX - 1 if two_targets else TILED_NMAX):
        return False
    return 0 < _lib.load().b200id_gp_shared_workspace_bytes(n_shapes, n, nv, int(two_targets)) <= SHARED_MAX_TABLE_BYTES


def batch_eval_shared_device(kernel_id: int, theta_d: torch.Tensor, shape_of_d: torch.Tensor, shape_theta_d: torch.Tensor,
                             X_d: torch.Tensor, y_a: torch.Tensor, y_b: Optional[torch.Tensor], noise: float,
                             metric: int = METRIC_NLL, Xv_d: Optional[torch.Tensor] = None,
                             yv_a: Optional[torch.Tensor] = None, yv_b: Optional[torch.Tensor] = None,
                             scale_a=(0.0, 1.0), scale_b=(0.0, 1.0), out: Optional[torch.Tensor] = None):
    """Metrics of a candidate list with shared covariance shapes (see :func:`shape_groups`): the shape tables are
    evaluated once per distinct shape, every candidate scales its table (b200id_gp_batch_eval_shared); values
    bit-identical to :func:`batch_eval_device` / :func:`batch_eval_pair_device`.  Returns metric_a, or
    (metric_a, metric_b) when a second target ``y_b`` is given."""
    dev = theta_d.device
    c, n, s = theta_d.shape[0], X_d.numel(), shape_theta_d.shape[0]
    two = y_b is not None
    nv = 0 if (Xv_d is None or metric != METRIC_VAL_RMSE) else Xv_d.numel()
    if out is None:
        out = empty(2 * c if two else c, dev)
    ws = Workspace.get("gpshape", _lib.load().b200id_gp_shared_workspace_bytes(s, n, nv, int(two)), dev)
    _lib.call("b200id_gp_batch_eval_shared", C.c_int(kernel_id), ptr(theta_d), C.c_int64(c), ptr(shape_of_d), ptr(shape_theta_d),
              C.c_int(s), ptr(X_d), ptr(y_a), ptr(y_b), C.c_int(n), C.c_double(noise), C.c_int(metric), ptr(Xv_d), ptr(yv_a),
              ptr(yv_b), C.c_int(0 if Xv_d is None else Xv_d.numel()),
              C.c_double(scale_a[0]), C.c_double(scale_a[1]), C.c_double(scale_b[0]), C.c_double(scale_b[1]),
              ptr(out[:c]), ptr(out[c:]) if two else ptr(None), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return (out[:c], out[c:]) if two else out


def batch_eval_pair(kernel, thetas, X, y_a, y_b, noise: float, Xv=None, yv_a=None, yv_b=None,
                    scale_a=(0.0, 1.0), scale_b=(0.0, 1.0), kernel_ids: Optional[Sequence[int]] = None):
    """Host form of :func:`batch_eval_pair_device`: NumPy in, (metric_a, metric_b) NumPy out."""
    dev = require_cuda()
    kid = KERNEL_IDS[kernel][0] if isinstance(kernel, str) else int(kernel)
    th = to_device(_theta_block(thetas), dev)
    ids = None if kernel_ids is None else torch.as_tensor(np.asarray(kernel_ids, dtype=np.int32), device=dev)
    use_val = Xv is not None and yv_a is not None and yv_b is not None
    d = lambda a: to_device(a, dev)
    x_h = _points(X)
    xv_h = _points(Xv, "Xv") if use_val else None
    ma, mb = batch_eval_pair_device(kid, ids, th, d(x_h), d(_targets(y_a, x_h.size, "y_a")), d(_targets(y_b, x_h.size, "y_b")),
                                    float(noise), METRIC_VAL_RMSE if use_val else METRIC_NLL,
                                    d(xv_h) if use_val else None, d(_targets(yv_a, xv_h.size, "yv_a")) if use_val else None,
                                    d(_targets(yv_b, xv_h.size, "yv_b")) if use_val else None, scale_a, scale_b)
    both = torch.cat([ma, mb]).cpu().numpy()
    return both[:len(both) // 2], both[len(both) // 2:]


def first_strict_min_device(metric_d: torch.Tensor):
    """(best[1], idx[1]) device tensors: first strict minimum in index order, NaNs skipped."""
    dev = metric_d.device
    best = empty(1, dev)
    idx = torch.empty(1, dtype=torch.int64, device=dev)
    ws = Workspace.get("gp", 256, dev)
    _lib.call("b200id_first_strict_min", ptr(metric_d), C.c_int64(metric_d.numel()), ptr(best), ptr(idx), ptr(ws),
              C.c_size_t(ws.numel()), stream_ptr())
    return best, idx


def batch_eval(kernel, thetas, X, y, noise, Xv=None, yv=None, y_mean: Optional[float] = None,
               y_scale: Optional[float] = None, kernel_ids: Optional[Sequence[int]] = None) -> np.ndarray:
    """Candidate metrics on the host: NLL, or validation RMSE when (Xv, yv) are given.  ``noise`` is a
    scalar, or one value per candidate."""
    dev = require_cuda()
    kid = KERNEL_IDS[kernel][0] if isinstance(kernel, str) else int(kernel)
    th = to_device(_theta_block(thetas), dev)
    ids = None if kernel_ids is None else torch.as_tensor(np.asarray(kernel_ids, dtype=np.int32), device=dev)
    use_val = Xv is not None and yv is not None
    if np.ndim(noise) == 0:
        noise_arg = float(noise)
    else:
        noise_arg = to_device(np.ascontiguousarray(noise, dtype=np.float64).ravel(), dev)
        if noise_arg.numel() != th.shape[0]:
            raise ValueError(f"batch_eval: {noise_arg.numel()} noise values for {th.shape[0]} candidates")
    x_h = _points(X)
    xv_h = _points(Xv, "Xv") if use_val else None
    m = batch_eval_device(
        kid, ids, th, to_device(x_h, dev), to_device(_targets(y, x_h.size), dev), noise_arg,
        METRIC_VAL_RMSE if use_val else METRIC_NLL,
        to_device(xv_h, dev) if use_val else None, to_device(_targets(yv, xv_h.size, "yv"), dev) if use_val else None,
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
    x_h = _points(X)
    xv_h = _points(Xv, "Xv") if use_val else None
    m = batch_eval_device(
        kid, None, th, to_device(x_h, dev), to_device(_targets(y, x_h.size), dev), float(noise),
        METRIC_VAL_RMSE if use_val else METRIC_NLL,
        to_device(xv_h, dev) if use_val else None, to_device(_targets(yv, xv_h.size, "yv"), dev) if use_val else None,
        0.0 if ym is None else float(ym), 1.0 if ys is None else float(ys))
    best, idx = first_strict_min_device(m)
    return int(idx.item()), float(best.item()), m.cpu().numpy()


# Last factorisation, keyed by its exact inputs.  grid_search_hyperparameters fits the winner (training-RMSE
# diagnostic, reference grid_search.py:209-232) and GaussianProcessRegressor.fit then factorises the very same
# system (gpr_fitting.py:73-87): same inputs, same deterministic kernel, so the second request is served from here.
_fit_memo = {}


def _fit_key(kernel_id: int, theta, X, y, diag_add: float):
    return (int(kernel_id), _theta_block(theta)[0].tobytes(), _flat(X).tobytes(), _flat(y).tobytes(), float(diag_add))


def _memo_get(key, want_kinv: bool):
    hit = _fit_memo.get(key)
    if hit is None or (want_kinv and hit[1] is None):
        return None
    alpha, kinv, used_pinv = hit
    return alpha.copy(), (kinv.copy() if (want_kinv and kinv is not None) else None), used_pinv


def _memo_put(key, alpha, kinv, used_pinv):
    _fit_memo.clear()
    _fit_memo[key] = (alpha.copy(), None if kinv is None else kinv.copy(), used_pinv)


def fit(kernel_id: int, theta, X, y, diag_add: float, want_kinv: bool = True):
    """(alpha[n], K_inv[n,n] or None, used_pinv) -- Cholesky path, pseudo-inverse path when a pivot is <= 0."""
    dev = require_cuda()
    x_h = _points(X)
    y_h = _targets(y, x_h.size)
    key = _fit_key(kernel_id, theta, x_h, y_h, diag_add)
    hit = _memo_get(key, want_kinv)
    if hit is not None:
        return hit
    x, yy = to_device(x_h, dev), to_device(y_h, dev)
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
    out = (alpha.cpu().numpy(), (Kinv.view(n, n).cpu().numpy() if Kinv is not None else None), used_pinv)
    _memo_put(key, *out)
    return out


def fit_batch(kernel_id: int, thetas, X, ys, diag_add: float):
    """(alpha [B, n], info [B]) of B fits sharing X, the kernel family and the diagonal term, in one launch
    (b200id_gp_fit_batch; n <= 127 or n > 160).  info[b] > 0 = failing pivot of problem b (use :func:`fit`)."""
    dev = require_cuda()
    th = to_device(_theta_block(thetas), dev)
    yy = to_device(np.ascontiguousarray(np.asarray(ys, dtype=np.float64)).reshape(th.shape[0], -1), dev)
    x = to_device(_points(X), dev)
    b, n = yy.shape
    if n != x.numel():
        raise ValueError(f"fit_batch: ys has {n} values per problem for {x.numel()} input points")
    alpha = empty(b * n, dev)
    info = torch.zeros(b, dtype=torch.int32, device=dev)
    ws = Workspace.get("gpfit", b * _lib.load().b200id_gp_fit_workspace_bytes(n), dev)
    _lib.call("b200id_gp_fit_batch", C.c_int(kernel_id), ptr(th), ptr(x), ptr(yy), C.c_int(n), C.c_int(b), C.c_double(diag_add),
              ptr(alpha), ptr(info), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return alpha.view(b, n).cpu().numpy(), info.cpu().numpy()


def search_and_fit(kernel_id: int, start_theta, candidates, X, y, noise: float, fit_diag_add: float,
                   Xv=None, yv=None, y_mean: Optional[float] = None, y_scale: Optional[float] = None):
    """Whole grid search of one target in ONE host<->device round trip.

    Evaluates the metric of ``start_theta`` (the reference's "initial" print, grid_search.py:143-157) and of
    every candidate in one batched launch, picks the first strict minimum over the candidates on the device
    (:186-196), factorises K(theta*) + fit_diag_add I for the winner and evaluates the fitted mean at the
    training inputs (:209-232) -- four launches queued back to back, then a single device->host read.
    Returns dict(start_metric, metrics[C], index, best, alpha[n], train_pred[n], used_pinv)."""
    dev = require_cuda()
    cand = _theta_block(candidates)
    c = cand.shape[0]
    x_h = _points(X)
    n = x_h.size
    y_h = _targets(y, n)
    use_val = Xv is not None and yv is not None
    xv_h = _points(Xv, "Xv") if use_val else None
    nv = xv_h.size if use_val else 0
    # one staging buffer up: [thetas (c+1)*3 | X n | y n | Xv nv | yv nv]
    parts = [np.vstack([_theta_block(start_theta), cand]).ravel(), x_h, y_h] + ([xv_h, _targets(yv, nv, "yv")] if use_val else [])
    up = to_device(np.concatenate(parts), dev)
    o = 0
    th = up[o:o + 3 * (c + 1)].view(c + 1, MAX_PARAMS); o += 3 * (c + 1)
    x_d = up[o:o + n]; o += n
    y_d = up[o:o + n]; o += n
    xv_d = yv_d = None
    if use_val:
        xv_d = up[o:o + nv]; o += nv
        yv_d = up[o:o + nv]
    # one result buffer down: [metrics c+1 | best | idx | info | alpha n | pred n]
    down = empty(c + 4 + 2 * n, dev)
    m = down[:c + 1]
    metric = METRIC_VAL_RMSE if use_val else METRIC_NLL
    scale = (0.0 if y_mean is None else float(y_mean), 1.0 if y_scale is None else float(y_scale))
    groups = shape_groups(kernel_id, cand)
    if groups is not None and shared_fits(groups[1].shape[0], n, nv, False):
        # product grid: the start point alone, then the list through its shared covariance shapes
        batch_eval_device(kernel_id, None, th[:1], x_d, y_d, float(noise), metric, xv_d, yv_d, scale[0], scale[1], out=m[:1])
        batch_eval_shared_device(kernel_id, th[1:], torch.as_tensor(groups[0], device=dev), to_device(groups[1], dev),
                                 x_d, y_d, None, float(noise), metric, xv_d, yv_d, None, scale, out=m[1:])
    else:
        batch_eval_device(kernel_id, None, th, x_d, y_d, float(noise), metric, xv_d, yv_d, scale[0], scale[1], out=m)
    best, idx = first_strict_min_device(m[1:])
    th_best = th[(idx + 1).clamp_(min=0)].reshape(-1).contiguous()        # idx = -1 (no finite metric) -> start theta
    alpha = down[c + 4:c + 4 + n]
    pred = down[c + 4 + n:]
    want_kinv = 127 < n
    kinv = empty(n * n, dev) if want_kinv else None
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = Workspace.get("gpfit", _lib.load().b200id_gp_fit_workspace_bytes(n), dev)
    _lib.call("b200id_gp_fit", C.c_int(kernel_id), ptr(th_best), ptr(x_d), ptr(y_d), C.c_int(n), C.c_double(fit_diag_add),
              ptr(alpha), ptr(kinv), ptr(info), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    _lib.call("b200id_gp_predict", C.c_int(kernel_id), ptr(th_best), ptr(x_d), C.c_int(n), ptr(alpha), ptr(None),
              ptr(x_d), C.c_int(n), ptr(pred), ptr(None), stream_ptr())
    down[c + 1:c + 2].copy_(best)
    down[c + 2:c + 3].copy_(idx)
    down[c + 3:c + 4].copy_(info)
    host = down.cpu().numpy()
    index, used_pinv = int(host[c + 2]), int(host[c + 3]) != 0
    res = dict(start_metric=float(host[0]), metrics=host[1:c + 1].copy(), index=index, best=float(host[c + 1]),
               alpha=host[c + 4:c + 4 + n].copy(), train_pred=host[c + 4 + n:].copy(), used_pinv=used_pinv)
    if index >= 0 and not used_pinv:
        _memo_put(_fit_key(kernel_id, cand[index], x_h, y_h, fit_diag_add), res["alpha"],
                  kinv.view(n, n).cpu().numpy() if kinv is not None else None, False)
    return res


def predict(kernel_id: int, theta, X_train, alpha, K_inv, Xs, return_std: bool = False):
    dev = require_cuda()
    x, xs = to_device(_points(X_train, "X_train"), dev), to_device(_points(Xs, "Xs"), dev)
    n, m = x.numel(), xs.numel()
    th = to_device(_theta_block(theta)[0], dev)
    a = to_device(_targets(alpha, n, "alpha"), dev)
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
