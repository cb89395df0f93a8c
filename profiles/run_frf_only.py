#!/usr/bin/env python3
"""Launch the uniform FRF projection at the bench shapes a few times (target of `ncu --set full -k regex:frf_project`).

    python profiles/run_frf_only.py [nd ...]          # default: 100 150
Prints the per-call time of b200id_frf_project_uniform (all its launches) and, with B200ID_FRF_MMA=0 in the
environment, of the recurrence kernels for comparison; also the largest relative difference between the two paths
when B200ID_FRF_COMPARE=1 (spawns nothing: the library reads the switch once, so the comparison is done by the caller
running the script twice and np.save-ing the sums)."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import frf, pipeline as P
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
n = int(os.environ.get("FRF_N", 10_000_000))
nds = [int(a) for a in sys.argv[1:]] or [100, 150]
gen = torch.Generator(device=dev); gen.manual_seed(0)
t = torch.arange(n, dtype=torch.float64, device=dev) * 0.002
u = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
y = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
tag = "mma" if os.environ.get("B200ID_FRF_MMA", "0") == "1" else "recurrence"
for nd in nds:
    w = to_device(P.log_grid_omega(nd), dev)
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = frf.project_uniform_device(t, u, y, w, 0.0, 0.002); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"frf_project_uniform[{tag}] n={n} nd={nd}: {best:.3f} ms  ({best / (n * nd) * 1e9:.3f} ms per 1e9 sample-bins)")
    dump = os.environ.get("FRF_DUMP")
    if dump:
        np.save(f"{dump}_{tag}_{nd}.npy", out.cpu().numpy())
