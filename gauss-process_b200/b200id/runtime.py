"""Device plumbing shared by the host wrappers: torch owns buffers and streams, nothing else."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

F64 = torch.float64


def require_cuda() -> torch.device:
    """Current CUDA device, or raise: the b200id path has no CPU fallback."""
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("b200id needs a CUDA device (B200, sm_100a); there is no CPU fallback for this path")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device pointer handoff needs a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def to_device(a, dev: torch.device, dtype=F64) -> torch.Tensor:
    """Host array (NumPy or CPU torch, pinned or not) or device tensor -> contiguous device tensor."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64 if dtype == F64 else None))
    if t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


def empty(n, dev, dtype=F64) -> torch.Tensor:
    return torch.empty(n, dtype=dtype, device=dev)


class Workspace:
    """Grow-only device scratch buffer, one per (device, purpose)."""

    _pool = {}

    @classmethod
    def get(cls, key: str, nbytes: int, dev: torch.device) -> torch.Tensor:
        k = (key, dev.index)
        buf = cls._pool.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            cls._pool[k] = buf
        return buf
