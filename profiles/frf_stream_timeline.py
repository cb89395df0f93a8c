#!/usr/bin/env python3
"""Event timeline (ms from the first enqueue) of the streamed host FRF path at the bench shape."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
import bench
from b200id import frf
from b200id.runtime import require_cuda, to_device
from src.frequency_transform.frf_estimator import matlab_freq_grid
dev = require_cuda()
n = 10_000_000
_, val_h = bench.make_records(n, 100, seed=0, dev=dev)
t, u, y = [torch.from_numpy(x).pin_memory().numpy() for x in val_h]
for nd in (100, 150):
    _, w = matlab_freq_grid(nd, -1.0, 2.3); w_d = to_device(w, dev); span = float(t[-1] - t[0])
    for chunks in (4, 6):
        for rep in range(3):
            tr = []
            frf.record_coefficients_streamed(t, u, y, w_d, span, w_max=float(w.max()), n_chunks=chunks, trace=tr)
            torch.cuda.synchronize()
        start = tr[0][1]
        print(f"nd={nd} chunks={chunks}: " + "  ".join(f"{lab}={start.elapsed_time(ev):.2f}" for lab, ev in sorted(tr[1:], key=lambda p: start.elapsed_time(p[1]))))
