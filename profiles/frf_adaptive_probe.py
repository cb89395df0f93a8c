#!/usr/bin/env python3
"""Bin-adaptive FRF projection (frf_adaptive.cuh) on the two records of BASELINE config 2: 10^7-sample multisine at
N_d = 100 and 10^7-sample square wave at 150 bins.  Prints the time of b200id_frf_project_uniform per form, the number
of bins that took the rounding term, and the largest per-bin relative difference to the all-recurrence form 1.

    python profiles/frf_adaptive_probe.py                # timings (CUDA events)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv \
        python profiles/frf_adaptive_probe.py once       # one call per form: the per-kernel split
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
from b200id import _lib, frf, synth
from b200id.runtime import require_cuda, to_device

once = "once" in sys.argv
n = 10_000_000
dev = require_cuda()
lib = _lib.load()
out = {}
for nd, kind in ((100, "multisine"), (150, "square")):
    t, u, y = synth.make_record(n, 100, kind, seed=1)
    _, w = synth.log_grid(nd)
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    span = float(t[-1] - t[0])
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    ref = None
    for form, force, tag in ((1, 0, "form1"), (2, 0, "adaptive"), (2, 1, "all_listed")):
        lib.b200id_frf_select_recurrence(C.c_int(form))
        lib.b200id_frf_adaptive_force(C.c_int(force))
        reps = 1 if once else 10
        for _ in range(0 if once else 3):
            p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        e1.record()
        torch.cuda.synchronize()
        key = f"{kind}_nd{nd}_{tag}"
        out[key + "_ms"] = round(e0.elapsed_time(e1) / reps, 4)
        U, Y = frf.finalize_device(p, nd, 2.0 / span)
        if ref is None:
            ref = (U, Y)
        else:
            out[key + "_bins_listed"] = frf.adaptive_diag(nd, dev)[0]
            out[key + "_vs_form1_rel_max"] = float(max(((U - ref[0]).abs() / ref[0].abs()).max(), ((Y - ref[1]).abs() / ref[1].abs()).max()))
    del t_d, u_d, y_d
lib.b200id_frf_select_recurrence(C.c_int(-1))
lib.b200id_frf_adaptive_force(C.c_int(-1))
print(json.dumps(out, indent=1))
