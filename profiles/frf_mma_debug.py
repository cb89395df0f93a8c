#!/usr/bin/env python3
"""Per-bin comparison of the matrix-product FRF path with the recurrence path on one synthetic record.
Run once per library / switch and dump; then compare:  python profiles/frf_mma_debug.py dump TAG ; ... compare A B"""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
cases = [(100_000, 50, "multisine", 1.9e6), (100_000, 50, "multisine", 0.0), (65_537, 150, "square", 123.4567), (300_001, 150, "square", 0.0)]
if sys.argv[1] == "dump":
    import torch
    from b200id import frf, synth
    from b200id.runtime import require_cuda, to_device
    from oracle import frf as ofrf
    dev = require_cuda()
    out = {}
    for ci, (n, nd, kind, shift) in enumerate(cases):
        t, u, y = synth.make_record(n, min(nd, 100), kind, seed=2)
        t = t + shift
        _, w = ofrf.freq_grid(nd)
        span = float(t[-1] - t[0])
        t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
        sums = frf.moments_device(u_d, y_d, (0, n))
        part = frf.project_uniform_device(t_d, u_d, y_d, w_d, float(t[0]), span / (n - 1), (0, n), sums, 1.0 / n)
        out[f"c{ci}"] = part.cpu().numpy()
    np.savez(f"gpurun_out/frfdbg_{sys.argv[2]}.npz", **out)
else:
    A = np.load(f"gpurun_out/frfdbg_{sys.argv[2]}.npz"); B = np.load(f"gpurun_out/frfdbg_{sys.argv[3]}.npz")
    for ci, (n, nd, kind, shift) in enumerate(cases):
        a, b = A[f"c{ci}"].reshape(4, nd), B[f"c{ci}"].reshape(4, nd)
        for nm, r in (("u", 0), ("y", 2)):
            mag = np.hypot(b[r], b[r + 1]); err = np.hypot(a[r] - b[r], a[r + 1] - b[r + 1]) / mag
            k = int(err.argmax())
            print(f"case {ci} {kind} n={n} nd={nd} shift={shift:g} {nm}: max rel {err.max():.3e} at bin {k} (|S|={mag[k]:.3e}, median |S|={np.median(mag):.3e}); median rel {np.median(err):.2e}")
