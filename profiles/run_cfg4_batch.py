#!/usr/bin/env python3
"""BASELINE config 4 batch: 16 candidates (NLL) at N_d = 2048 in lockstep; target of `ncu -k regex:cholb_`.

    python profiles/run_cfg4_batch.py [reps]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp, synth                                   # noqa: E402
from b200id.runtime import require_cuda, to_device             # noqa: E402

dev = require_cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, nc = 2048, 16
_, w = synth.log_grid(n)
X = np.log10(w); X = (X - X.mean()) / X.std()
y = np.sin(2.0 * X) + 0.05 * np.random.default_rng(4).standard_normal(n)
X_d, y_d = to_device(X, dev), to_device(y, dev)
th = to_device(gp._theta_block(np.column_stack([np.linspace(0.5, 2.0, nc), np.linspace(0.3, 1.0, nc)])), dev)
for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); m = gp.batch_eval_device(gp.KERNEL_IDS["rbf"][0], None, th, X_d, y_d, 1e-3, gp.METRIC_NLL); e.record()
    torch.cuda.synchronize()
    print(f"{nc} candidates, n = {n}: {s.elapsed_time(e):.3f} ms, NLL[0] = {float(m[0]):.6f}")
