"""Host side of the kernel-regularised FIR identification (reference src/fir_model/kernel_regularized.py) through the
C ABI (include/b200id.h: b200id_kr_summaries / b200id_kr_nll_batch / b200id_kr_posterior).  NumPy in, NumPy out;
the regressor summaries stay on the device between calls (:class:`Summaries`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from . import _lib
from .runtime import F64, Workspace, empty, ptr, require_cuda, stream_ptr, to_device

KERNEL_IDS = {"dc": (0, 3), "ss": (1, 2), "si": (2, 4)}          # name -> (id, kernel parameters before log sigma^2)


def kernel_id(kernel: str) -> Tuple[int, int]:
    try:
        return KERNEL_IDS[kernel.lower()]
    except KeyError:
        raise ValueError(f"Unknown kernel: {kernel}") from None


def _ws(L: int, batch: int, dev):
    nbytes = _lib.load().b200id_kr_workspace_bytes(int(L), int(batch))
    return Workspace.get("kr", nbytes, dev)


@dataclass
class Summaries:
    """G = U^T U [L, L], b = U^T y [L] on the device, yy = y^T y and the sample count on the host (:337-340)."""
    G: torch.Tensor
    b: torch.Tensor
    yy: float
    n: int
    L: int


def summaries(u, y, L: int) -> Summaries:
    """Regressor summaries of the zero-padded Toeplitz regressor, computed on the device from u, y."""
    dev = require_cuda()
    u_d, y_d = to_device(np.asarray(u, dtype=np.float64).ravel(), dev), to_device(np.asarray(y, dtype=np.float64).ravel(), dev)
    if u_d.numel() != y_d.numel():
        raise ValueError("u and y must have the same length")
    n = u_d.numel()
    G, b, yy = empty(L * L, dev), empty(L, dev), empty(1, dev)
    ws = _ws(L, 1, dev)
    _lib.call("b200id_kr_summaries", ptr(u_d), ptr(y_d), C.c_int64(n), C.c_int(L), ptr(G), ptr(b), ptr(yy), ptr(ws),
              C.c_size_t(ws.numel()), stream_ptr())
    return Summaries(G.view(L, L), b, float(yy.item()), n, L)


def summaries_from_host(G, b, yy: float, n: int) -> Summaries:
    """Wrap summaries computed elsewhere (the reference's ``_nll(theta, G, b, yy, N, L, kernel)`` signature)."""
    dev = require_cuda()
    G = np.asarray(G, dtype=np.float64)
    L = G.shape[0]
    return Summaries(to_device(G, dev).view(L, L), to_device(np.asarray(b, dtype=np.float64).ravel(), dev), float(yy), int(n), L)


def nll_batch(kernel: str, thetas, s: Summaries, jitter: float = 1e-10) -> Tuple[np.ndarray, np.ndarray]:
    """(nll [B], info [B]) for B unconstrained points [B, kdim + 1]; NaN where the reference would raise LinAlgError."""
    kid, kdim = kernel_id(kernel)
    th = np.ascontiguousarray(np.asarray(thetas, dtype=np.float64).reshape(-1, kdim + 1))
    dev = s.G.device
    B = th.shape[0]
    th_d = to_device(th, dev)
    out = empty(B, dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    ws = _ws(s.L, B, dev)
    _lib.call("b200id_kr_nll_batch", C.c_int(kid), ptr(th_d), C.c_int(B), ptr(s.G), ptr(s.b), C.c_double(s.yy),
              C.c_double(float(s.n)), C.c_int(s.L), C.c_double(jitter), ptr(out), ptr(info), ptr(ws), C.c_size_t(ws.numel()),
              stream_ptr())
    return out.cpu().numpy(), info.cpu().numpy()


def posterior(kernel: str, theta, s: Summaries, want_K: bool = True):
    """(g_hat [L], sigma^2, K [L, L] or None) at one point (:264-291); raises LinAlgError where the reference does."""
    kid, kdim = kernel_id(kernel)
    dev = s.G.device
    th_d = to_device(np.asarray(theta, dtype=np.float64).reshape(kdim + 1), dev)
    g, sig2 = empty(s.L, dev), empty(1, dev)
    K = empty(s.L * s.L, dev) if want_K else None
    info = torch.empty(1, dtype=torch.int32, device=dev)
    ws = _ws(s.L, 1, dev)
    _lib.call("b200id_kr_posterior", C.c_int(kid), ptr(th_d), ptr(s.G), ptr(s.b), C.c_int(s.L), ptr(g), ptr(K), ptr(sig2),
              ptr(info), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("kernel-regularised FIR: matrix not positive definite")
    return g.cpu().numpy(), float(sig2.item()), (K.view(s.L, s.L).cpu().numpy() if want_K else None)


def build_kernel(L: int, kernel: str, theta) -> np.ndarray:
    """K for unconstrained kernel parameters theta [kdim] (:192-217), evaluated on the device."""
    kid, kdim = kernel_id(kernel)
    dev = require_cuda()
    th_d = to_device(np.concatenate([np.asarray(theta, dtype=np.float64).reshape(kdim), [0.0]]), dev)
    K = empty(L * L, dev)
    info = torch.empty(1, dtype=torch.int32, device=dev)
    ws = _ws(L, 1, dev)
    # G and b are not read on this path: any readable device pointers do
    _lib.call("b200id_kr_posterior", C.c_int(kid), ptr(th_d), ptr(K), ptr(K), C.c_int(L), ptr(None), ptr(K), ptr(None),
              ptr(info), ptr(ws), C.c_size_t(ws.numel()), stream_ptr())
    return K.view(L, L).cpu().numpy()
