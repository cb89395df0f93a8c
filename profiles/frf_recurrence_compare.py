"""Time the three evaluations of b200id_frf_project_uniform side by side (10^7 samples, N_d = 100 and 150):
three-term recurrence, Goertzel recurrence, DMMA main term + FP32 Goertzel rounding term (default).

    python profiles/frf_recurrence_compare.py            # on the GPU box
"""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]

import numpy as np
import torch

from b200id import _lib, frf, synth
from b200id.runtime import require_cuda, to_device

dev = require_cuda()
lib = _lib.load()
n = 10_000_000
out = {}
for nd in (100, 150):
    t, u, y = synth.make_record(n, 100, "multisine" if nd == 100 else "square", seed=1)
    _, w = synth.log_grid(nd)
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    span = float(t[-1] - t[0])
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    res = {}
    for which, name in ((0, "three_term"), (1, "goertzel"), (2, "dmma_goertzel")):
        lib.b200id_frf_select_recurrence(C.c_int(which))
        for _ in range(3):
            p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms": e0.elapsed_time(e1) / 10, "partial": p.cpu().numpy()}
    a, b = res["three_term"]["partial"], res["goertzel"]["partial"]
    k = 4 * nd
    mag = np.hypot(a[:nd], a[nd:2 * nd])
    magy = np.hypot(a[2 * nd:3 * nd], a[3 * nd:])
    du = np.hypot(a[:nd] - b[:nd], a[nd:2 * nd] - b[nd:2 * nd]) / mag
    dy = np.hypot(a[2 * nd:3 * nd] - b[2 * nd:3 * nd], a[3 * nd:] - b[3 * nd:]) / magy
    c = res["dmma_goertzel"]["partial"]
    du2 = np.hypot(b[:nd] - c[:nd], b[nd:2 * nd] - c[nd:2 * nd]) / mag
    dy2 = np.hypot(b[2 * nd:3 * nd] - c[2 * nd:3 * nd], b[3 * nd:] - c[3 * nd:]) / magy
    out[f"nd{nd}"] = {"three_term_ms": res["three_term"]["ms"], "goertzel_ms": res["goertzel"]["ms"],
                      "dmma_goertzel_ms": res["dmma_goertzel"]["ms"],
                      "max_rel_diff_dmma_vs_goertzel_u": float(du2.max()), "max_rel_diff_dmma_vs_goertzel_y": float(dy2.max()),
                      "ms_per_1e9_sample_bins": {k2: v["ms"] / (n * nd / 1e9) for k2, v in res.items()},
                      "max_rel_diff_u": float(du.max()), "max_rel_diff_y": float(dy.max())}
lib.b200id_frf_select_recurrence(C.c_int(-1))
print(json.dumps(out, indent=1))
