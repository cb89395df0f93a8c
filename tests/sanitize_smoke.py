#!/usr/bin/env python3
"""Every kernel once at small, ragged sizes -- meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck python tests/sanitize_smoke.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import frf, fir, gp, pipeline as P, synth      # noqa: E402
from b200id.runtime import require_cuda, to_device          # noqa: E402

dev = require_cuda()
rng = np.random.default_rng(0)
for n, nd in ((2, 1), (3, 5), (4097, 37), (70_001, 150), (9_000, 601)):
    t, u, y = synth.make_record(max(n, 16), min(nd, 100), "multisine", seed=1)
    t, u, y = t[:n], u[:n], y[:n]
    w = P.log_grid_omega(nd)
    frf.demod_pair(t, u, y, w)
    frf.demod(t, u, w, subtract_mean=False)
    tj = np.sort(t + rng.uniform(-0.3, 0.3, n) * 0.002)
    if n > 2:
        frf.demod_pair(tj, u, y, w)
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    if n > 8:
        frf.project_device(t_d, u_d, y_d, w_d, own=(3, n - 2))
        frf.project_uniform_device(t_d, u_d, y_d, w_d, float(t[0]), 0.002, own=(1, n - 1))
for n, taps in ((1, 1), (5, 3), (2049, 1024), (10_000, 2500), (3000, 5000), (70_001, 1024)):
    u = rng.standard_normal(n); g = rng.standard_normal(taps)
    fir.fir_validate(u, u * 0.5, g, want_pred=True)
fir.idft_taps(rng.standard_normal(9) + 1j * rng.standard_normal(9), 17)
fir.idft_taps(rng.standard_normal(1000) + 1j * rng.standard_normal(1000), 1024)
for n in (1, 2, 15, 16, 17, 50, 100, 127, 128, 160, 200):
    X = np.sort(rng.uniform(-1.7, 1.7, n)); y = rng.standard_normal(n)
    Xv = np.linspace(-1.5, 1.5, 33); yv = rng.standard_normal(33)
    for name in ("matern52", "dc", "sshf"):
        th = P.restart_population(name, 5)
        gp.batch_eval(name, th, X, y, 1e-4)
        gp.batch_eval(name, th, X, y, 1e-4, Xv=Xv, yv=yv, y_mean=0.1, y_scale=2.0)
    kid = gp.KERNEL_IDS["matern32"][0]
    a, Kinv, _ = gp.fit(kid, [0.5, 1.0], X, y, 1e-4)
    gp.predict(kid, [0.5, 1.0], X, a, Kinv, Xv, True)
    if n <= 160:
        gp.fit(gp.KERNEL_IDS["ss1"][0], [1.0, 1.0], X, y, 0.0)        # indefinite -> pinv branch
    gp.gram(kid, [0.5, 1.0], X, Xv)
m = torch.tensor([3.0, float("nan"), 1.0, 1.0], dtype=torch.float64, device=dev)
gp.first_strict_min_device(m)
A = torch.from_numpy(np.eye(130) * 2.0 + 0.01).to(dev)
gp.cholesky_blocked_device(A)
torch.cuda.synchronize()
print("sanitize_smoke: all kernels ran")
