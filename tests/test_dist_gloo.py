"""Multi-rank host logic of the sharded path on CPU: world_size-2 gloo processes.

Covers what the N>1 bench does between kernels: candidate round-robin sharding + the argmin gather
(first strict minimum in GLOBAL scan order, NaNs skipped, ties to the smallest index), contiguous
segment sharding with halos, and the all-reduce of per-bin sums (emulated shard sums from the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200id import dist as bdist
from oracle import frf as ofrf, gp as ogp, fir as ofir


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = bdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and bdist.world() == (rank, world)
    out = {}
    # ---- candidate sharding + argmin gather against the sequential scan
    rng = np.random.default_rng(0)
    for case in range(6):
        metrics = rng.integers(0, 12, 257).astype(np.float64)
        metrics[rng.integers(0, 257, 40)] = np.nan
        if case == 4:
            metrics[:] = np.nan
        if case == 5:
            metrics[:] = 1e10                                # every candidate failed: index 0 must win
        mine = bdist.shard_round_robin(len(metrics), rank, world)
        li, lb = ogp.first_strict_min(metrics[mine])
        gi, gb = bdist.argmin_gather(lb, int(mine[li]) if li >= 0 else -1, device=torch.device("cpu"))
        ei, eb = ogp.first_strict_min(metrics)
        assert (gi, gb) == (ei, eb), (case, gi, gb, ei, eb)
    # ---- time-segment sharding of one record: per-shard trapezoid sums add up to the whole-record sums
    n, nd = 4001, 12
    t = np.arange(n) * 0.002
    x = np.random.default_rng(1).standard_normal(n)
    _, wgrid = ofrf.freq_grid(nd)
    lo, hi = bdist.shard_contiguous(n, rank, world)
    wt = 0.5 * np.r_[t[1] - t[0], t[2:] - t[:-2], t[-1] - t[-2]]
    part = np.array([[np.sum(wt[lo:hi] * x[lo:hi] * np.cos(wk * t[lo:hi])), np.sum(wt[lo:hi] * x[lo:hi] * np.sin(wk * t[lo:hi]))] for wk in wgrid])
    tot = torch.from_numpy(part.copy())
    bdist.allreduce_sum_(tot)
    whole = np.array([[np.sum(wt * x * np.cos(wk * t)), np.sum(wt * x * np.sin(wk * t))] for wk in wgrid])
    assert np.allclose(tot.numpy(), whole, rtol=1e-12, atol=1e-15)
    ref = ofrf.demod_trapz(t, x, wgrid, subtract_mean=False)
    got = (2.0 / (t[-1] - t[0])) * (tot.numpy()[:, 0] - 1j * tot.numpy()[:, 1])
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-14)
    # ---- FIR output shards with history: additive error sums
    u = np.random.default_rng(2).standard_normal(n); g = np.random.default_rng(3).standard_normal(64) / 64
    y = ofir.fir_predict(u, g) + 0.01
    yp = ofir.fir_predict(u, g)
    skip = 6
    sl = slice(max(lo, skip), hi)
    sums = torch.tensor([np.sum((y[sl] - yp[sl]) ** 2), np.sum(y[sl] ** 2), np.sum((y[sl] - y[skip]) ** 2), float(hi - max(lo, skip))])
    bdist.allreduce_sum_(sums)
    m, _ = ofir.fir_validate(u, y, g)
    assert abs(np.sqrt(sums[0].item() / sums[3].item()) - m["rmse"]) < 1e-14
    dist.barrier()
    q.put((rank, "ok"))
    dist.destroy_process_group()


def test_two_rank_gloo_exchanges():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5)[0] for _ in range(world)) == [0, 1]


def test_shard_helpers():
    assert bdist.shard_round_robin(10, 1, 4).tolist() == [1, 5, 9]
    cover = []
    for r in range(3):
        lo, hi = bdist.shard_contiguous(10, r, 3)
        cover += list(range(lo, hi))
    assert cover == list(range(10))
    assert bdist.pick_first_strict_min([(3.0, 7), (3.0, 2), (float("nan"), 0), (float("inf"), -1)]) == (2, 3.0)
    assert bdist.pick_first_strict_min([(float("inf"), -1)]) == (-1, float("inf"))


def _worker_multi(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    bdist.init_from_env(backend="gloo")
    rng = np.random.default_rng(7)
    a = rng.integers(0, 9, 101).astype(np.float64); b = rng.integers(0, 9, 101).astype(np.float64)
    b[::3] = np.nan
    pairs = []
    for metrics in (a, b):
        mine = bdist.shard_round_robin(len(metrics), rank, world)
        li, lb = ogp.first_strict_min(metrics[mine])
        pairs.append((lb, int(mine[li]) if li >= 0 else -1))
    got = bdist.argmin_gather_multi(pairs, device=torch.device("cpu"))
    assert got == [ogp.first_strict_min(a), ogp.first_strict_min(b)], got
    q.put(rank)
    dist.destroy_process_group()


def test_two_rank_multi_argmin_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_multi, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def _worker_device_pick(rank, world, port, q):
    """pipeline._device_pick (the pick bench.py's N>1 step uses: both searches in ONE all-gather, winner's
    hyperparameters taken from the full list without a host round trip) against the sequential scan."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    bdist.init_from_env(backend="gloo")
    from b200id import pipeline as P
    rng = np.random.default_rng(3)
    C = 203
    full = np.c_[rng.uniform(0.1, 5, C), rng.uniform(0.1, 5, C), np.zeros(C)]
    mine = bdist.shard_round_robin(C, rank, world)
    cs = P.CandidateSet(kernel="matern52", theta_host=full[mine], theta=torch.from_numpy(full[mine]), kernel_ids=None,
                        global_index=mine, full_theta=full, full_theta_dev=torch.from_numpy(full),
                        global_index_dev=torch.from_numpy(mine))
    for case in range(5):
        a = rng.integers(0, 7, C).astype(np.float64); b = rng.integers(0, 7, C).astype(np.float64)
        a[rng.integers(0, C, 30)] = np.nan
        if case == 3:
            b[:] = np.nan                                      # no finite candidate anywhere: index -1
        if case == 4:
            a[:] = 1e10                                        # every candidate failed: global index 0 wins
        pends = []
        for metrics in (a, b):
            li, lb = ogp.first_strict_min(metrics[mine])
            pends.append(torch.tensor([lb if li >= 0 else float("inf"), float(li)], dtype=torch.float64))
        picks = P._device_pick(cs, pends)
        for (theta, gidx, best), metrics in zip(picks, (a, b)):
            ei, eb = ogp.first_strict_min(metrics)
            assert int(gidx) == ei, (case, int(gidx), ei)
            if ei >= 0:
                assert float(best) == eb and np.array_equal(theta.numpy(), full[ei])
    q.put(rank)
    dist.destroy_process_group()


def test_two_rank_device_pick_both_searches_one_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_device_pick, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
