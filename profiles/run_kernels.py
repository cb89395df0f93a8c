#!/usr/bin/env python3
"""Launch each hot kernel once at the bench shapes (for `ncu --set full`); prints CUDA-event times.

    python profiles/run_kernels.py [--n 10000000] [--nd 100] [--cands 8192]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import frf, fir, gp, pipeline as P          # noqa: E402
from b200id.runtime import require_cuda, to_device      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--nd", type=int, default=100)
ap.add_argument("--cands", type=int, default=8192)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = require_cuda()
gen = torch.Generator(device=dev); gen.manual_seed(0)
t = torch.arange(a.n, dtype=torch.float64, device=dev) * 0.002
u = torch.randn(a.n, dtype=torch.float64, device=dev, generator=gen)
y = torch.randn(a.n, dtype=torch.float64, device=dev, generator=gen)
w = to_device(P.log_grid_omega(a.nd), dev)
g = torch.randn(1024, dtype=torch.float64, device=dev, generator=gen) / 1024
X = to_device(np.linspace(-1.7, 1.7, a.nd), dev)
yy = torch.randn(a.nd, dtype=torch.float64, device=dev, generator=gen)
th = to_device(gp._theta_block(P.restart_population("matern52", a.cands)), dev)
Xv = to_device(np.linspace(-1.7, 1.7, 150), dev)
yv = torch.randn(150, dtype=torch.float64, device=dev, generator=gen)


def timed(name, fn):
    for _ in range(a.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e):.3f} ms")


timed("frf_project_general", lambda: frf.project_device(t, u, y, w))
timed("frf_project_uniform", lambda: frf.project_uniform_device(t, u, y, w, 0.0, 0.002))
timed("fir_eval", lambda: fir.fir_eval_device(u, y, g, 10, 0.0))
timed("fir_eval_2048taps", lambda: fir.fir_eval_device(u, y, torch.cat([g, g]), 10, 0.0))
timed("fir_eval_direct_100taps", lambda: fir.fir_eval_device(u, y, g[:100].contiguous(), 10, 0.0))
timed("gp_batch_nll", lambda: gp.batch_eval_device(3, None, th, X, yy, 1e-6))
timed("gp_batch_val", lambda: gp.batch_eval_device(3, None, th[:2048], X, yy, 1e-6, gp.METRIC_VAL_RMSE, Xv, yv, 0.0, 1.0))
nq = a.n // 5
timed("frf_project_uniform_fifth_of_record", lambda: frf.project_uniform_device(t, u, y, w, 0.0, 0.002, own=(nq, 2 * nq)))
timed("frf_uniformity_fifth_of_record", lambda: frf.uniformity_device_async(t, 0.0, 0.002, own=(nq, 2 * nq)))
timed("frf_moments", lambda: frf.moments_device(u, y, (0, a.n)))
