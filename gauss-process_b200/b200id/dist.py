"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Only the three exchanges the path really has:
  * FRF: one all-reduce(sum) of the per-bin cross-power numerator / denominator (3*nd doubles) when
    records are sharded, or of the raw partial sums (4*nd doubles) when one record is time-sharded;
  * GP candidate search: an all-gather of (best metric, global index) per rank, then the host picks
    the first strict minimum in global scan order (grid_search.py:186-196);
  * FIR validation: one all-reduce(sum) of the 4 fused error sums.
Everything else runs rank-local with no data-path collective.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; no-op for a single process."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            # B200ID_DIST_BACKEND=gloo: several ranks sharing ONE GPU (NCCL refuses that) -- the kernels still run on
            # the device, only the small exchanges go through gloo; used by the shard tests on a one-GPU box
            backend = os.environ.get("B200ID_DIST_BACKEND") or ("nccl" if n_dev > 0 else "gloo")
        if n_dev > 0:
            torch.cuda.set_device(local % n_dev)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    elif n_dev > 0:
        torch.cuda.set_device(local % n_dev)
    return rank, ws, local


def bind_to_gpu_numa_node(local_rank: int = None) -> dict:
    """Pin this process (CPU affinity, hence first-touch memory placement of the pinned host buffers it allocates
    afterwards) to the NUMA node its GPU hangs off, from sysfs.  With one process per GPU on a two-socket host the
    host->device copies of the ranks otherwise cross the socket interconnect and share its bandwidth.  No-op (and says
    so) when the topology cannot be read or the host has a single node.  Call before allocating host buffers."""
    info = {"bound": False}
    try:
        if not torch.cuda.is_available():
            return info
        dev = torch.cuda.current_device() if local_rank is None else local_rank
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        info.update(pci=bdf, numa_node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:                              # topology not readable: leave the affinity alone
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


def _host_staged() -> bool:
    """gloo has no device all-gather: device tensors are staged through the host for it (shard tests that run two
    ranks on one GPU); NCCL takes them as they are."""
    return dist.get_backend() == "gloo"


def all_gather_into(out: torch.Tensor, mine: torch.Tensor) -> torch.Tensor:
    """dist.all_gather_into_tensor for small FP64 tables, on NCCL (device) or gloo (host-staged)."""
    if _host_staged() and mine.is_cuda:
        tmp = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(tmp, mine.cpu())
        out.copy_(tmp)
        return out
    dist.all_gather_into_tensor(out, mine)
    return out


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (identity for a single process)."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t: torch.Tensor) -> torch.Tensor:
    """In-place max over ranks (identity for a single process)."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def shard_round_robin(n_items: int, rank: int, world_size: int) -> np.ndarray:
    """Global indices i with i % world_size == rank (keeps per-kernel cost balanced across ranks)."""
    return np.arange(rank, n_items, world_size, dtype=np.int64)


def shard_contiguous(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the rank's contiguous slice (time-segment / FIR output shards)."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def argmin_gather(local_best: float, local_global_idx: int, device=None) -> Tuple[int, float]:
    """Global first strict minimum from each rank's (best metric, global index of it).

    A rank with no finite candidate passes (inf, -1).  Ties across ranks go to the smallest global
    index, which reproduces the sequential ``metric < best`` scan of grid_search.py:194.
    """
    rank, ws = world()
    if ws == 1:
        return int(local_global_idx), float(local_best)
    if device is None or _host_staged():
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor([float(local_best), float(local_global_idx)], dtype=torch.float64, device=device)
    allv = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(allv, mine)
    return pick_first_strict_min([(float(v[0].item()), int(v[1].item())) for v in allv])


def argmin_gather_multi(local_pairs: Sequence[Tuple[float, int]], device=None) -> List[Tuple[int, float]]:
    """Several independent argmin gathers (e.g. the Re and Im searches) in ONE collective and one
    device->host read: local_pairs[k] = (best metric, global index) of search k on this rank."""
    rank, ws = world()
    if ws == 1:
        return [(int(i), float(m)) for m, i in local_pairs]
    if device is None or _host_staged():
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    k = len(local_pairs)
    mine = torch.tensor([[float(m), float(i)] for m, i in local_pairs], dtype=torch.float64, device=device).reshape(-1)
    out = torch.empty(ws * 2 * k, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, mine)
    tab = out.cpu().reshape(ws, k, 2)
    return [pick_first_strict_min([(float(tab[r, j, 0]), int(tab[r, j, 1])) for r in range(ws)]) for j in range(k)]


def pick_first_strict_min(pairs: Sequence[Tuple[float, int]]) -> Tuple[int, float]:
    best, best_idx = float("inf"), -1
    for m, i in pairs:
        if i < 0 or not (m == m):
            continue
        if m < best or (m == best and i < best_idx):
            best, best_idx = m, i
    return best_idx, best
