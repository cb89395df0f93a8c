#!/usr/bin/env python3
"""Where a public-API (host buffers) call spends its time: event timeline of the streamed FRF path and a
cProfile of one grid-search GaussianProcessRegressor.fit."""
import contextlib, cProfile, io, pstats, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
import bench
from b200id import frf, gp as bgp
from b200id.runtime import require_cuda, to_device
from src.frequency_transform.frf_estimator import matlab_freq_grid
from src.gpr import create_kernel, GaussianProcessRegressor
from sklearn.preprocessing import StandardScaler

dev = require_cuda()
n = 10_000_000
(train_h, val_h) = bench.make_records(n, 100, seed=0, dev=dev)
pinned = [torch.from_numpy(x).pin_memory() for x in val_h]
t, u, y = [x.numpy() for x in pinned]
for nd in (100, 150):
    _, w = matlab_freq_grid(nd, -1.0, 2.3)
    w_d = to_device(w, dev)
    span = float(t[-1] - t[0])
    for chunks in (1, 4, 8, 16):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            U, Y = frf.record_coefficients_streamed(t, u, y, w_d, span, w_max=float(w.max()), n_chunks=chunks)
            host_done = time.perf_counter()
            UY = torch.stack([U, Y]).cpu()
            t1 = time.perf_counter()
        print(f"nd={nd} chunks={chunks}: total {1e3*(t1-t0):.3f} ms (host enqueue {1e3*(host_done-t0):.3f} ms)")
    # bulk path for comparison
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, Y = frf.record_coefficients_device(to_device(t, dev), to_device(u, dev), to_device(y, dev), w_d, span, t0=float(t[0]), w_max=float(w.max()))
        UY = torch.stack([U, Y]).cpu(); t1 = time.perf_counter()
    print(f"nd={nd} bulk: total {1e3*(t1-t0):.3f} ms")
    # pure copy time
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a = [to_device(x, dev) for x in (t, u, y)]
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"copy of 3 arrays: {1e3*(t1-t0):.3f} ms = {3*8*n/(t1-t0)/1e9:.1f} GB/s")

# ---- one GPR fit through the mirror
rng = np.random.default_rng(0)
_, omega = matlab_freq_grid(100, -1.0, 2.3); _, omega_v = matlab_freq_grid(150, -1.0, 2.3)
Xs = StandardScaler(); Xn = Xs.fit_transform(np.log10(omega).reshape(-1, 1)); Xv = Xs.transform(np.log10(omega_v).reshape(-1, 1))
target = np.sin(3 * Xn.ravel()) + 0.01 * rng.standard_normal(100); vy = np.sin(3 * Xv.ravel())
def one_fit():
    with contextlib.redirect_stdout(io.StringIO()):
        sc = StandardScaler(); z = sc.fit_transform(target.reshape(-1, 1)).ravel()
        m = GaussianProcessRegressor(kernel=create_kernel("matern52"), noise_variance=1e-6)
        m.fit(Xn, z, optimize=True, use_grid_search=True, max_grid_combinations=5000, validation_X=Xv, validation_y=vy, y_scaler=sc)
    return m
for _ in range(5): one_fit()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): one_fit()
torch.cuda.synchronize(); print("one grid-search fit: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): one_fit()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(28); print(s.getvalue()[:6000])
