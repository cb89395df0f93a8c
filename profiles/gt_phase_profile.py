#!/usr/bin/env python3
"""Phase timing of gp_batch_tiled_kernel (needs a -DGT_PROFILE build of the library, e.g. B200ID_LIB=variants/gtprof.so):
per-warp clock64 deltas per phase, NLL mode and the bench's validation-RMSE pair mode, for one CTA per SM (the
latency of one candidate alone), one full wave (4 per SM) and many waves."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp, pipeline as P, _lib
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(0)
X = to_device(np.linspace(-1.7, 1.7, n), dev)
y = to_device(rng.standard_normal(n), dev); y2 = to_device(rng.standard_normal(n), dev)
Xv = to_device(np.linspace(-1.6, 1.6, 150), dev); yv = to_device(rng.standard_normal(150), dev); yv2 = to_device(rng.standard_normal(150), dev)
names = ["fill", "sync(fill)", "update_finish", "sync(update)", "diag", "lookahead", "sync(diag)", "panel", "sync(panel)", "backsub", "predict", "-", "bs_part2+sync", "bs_loads(w0)", "bs_steps(w0)", "bs_sync1"]
buf = (C.c_ulonglong * 16)()
def run(label, cands, fn):
    fn(); torch.cuda.synchronize()
    lib.b200id_debug_gt_profile(buf, 1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    lib.b200id_debug_gt_profile(buf, 0)
    tot = sum(buf[:16])
    print(f"== {label}: {cands} candidates, {s.elapsed_time(e):.4f} ms, {tot / (cands * 4):.0f} cycles per warp per candidate")
    print("   " + "  ".join(f"{nm} {buf[k] / (cands * 4):.0f} ({buf[k] / tot * 100:.0f}%)" for k, nm in enumerate(names) if buf[k]))
for cands in (148, 592, 900, 8192):
    th = to_device(gp._theta_block(P.restart_population("matern52", cands)), dev)
    run("NLL", cands, lambda: gp.batch_eval_device(3, None, th, X, y, 1e-6, gp.METRIC_NLL))
    run("val-RMSE pair", cands, lambda: gp.batch_eval_pair_device(3, None, th, X, y, y2, 1e-6, gp.METRIC_VAL_RMSE, Xv, yv, yv2, (0.0, 1.0), (0.0, 1.0)) if hasattr(gp, "batch_eval_pair_device") else None)
grid = P.grid_candidates("matern52")
for reps in (1, 9):
    g = np.tile(grid, (reps, 1)) if reps > 1 else grid
    for cnt in ((148, len(g)) if reps == 1 else (len(g),)):
        cs = P.CandidateSet.build("matern52", g[:cnt], dev)
        if cs.shape_of is None:
            continue
        run("shared-shape val-RMSE pair", cnt, lambda: gp.batch_eval_shared_device(3, cs.theta, cs.shape_of, cs.shape_theta, X, y, y2, 1e-6, gp.METRIC_VAL_RMSE, Xv, yv, yv2))
        run("direct val-RMSE pair (same list)", cnt, lambda: gp.batch_eval_pair_device(3, None, cs.theta, X, y, y2, 1e-6, gp.METRIC_VAL_RMSE, Xv, yv, yv2))
