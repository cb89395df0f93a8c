#!/usr/bin/env python3
"""Launch the uniform FRF projection at the bench shape a few times (target of `ncu --set full -k regex:frf_project_uniform`)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import frf, pipeline as P
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
n, nd = 10_000_000, 100
gen = torch.Generator(device=dev); gen.manual_seed(0)
t = torch.arange(n, dtype=torch.float64, device=dev) * 0.002
u = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
y = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
w = to_device(P.log_grid_omega(nd), dev)
for _ in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); frf.project_uniform_device(t, u, y, w, 0.0, 0.002); e.record(); torch.cuda.synchronize()
print(f"frf_project_uniform: {s.elapsed_time(e):.3f} ms")
