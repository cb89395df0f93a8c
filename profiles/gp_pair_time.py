"""Time of the bench's GP search (900 Matern-5/2 candidates x 2 targets, validation-RMSE metric, N_d = 100) and of a
4096-candidate NLL scan, for comparing library builds (B200ID_LIB):  python profiles/gp_pair_time.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
import numpy as np, torch
from b200id import gp as bgp, pipeline as P, synth
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 100
gi = P.GridInputs.get(nd, dev)
rng = np.random.default_rng(0)
y = to_device(rng.standard_normal(nd), dev)
cands = P.CandidateSet.build("matern52", P.grid_candidates("matern52"), dev)
Xv = to_device(np.linspace(-1.6, 1.6, 150), dev); yv = to_device(rng.standard_normal(150), dev)
kid = bgp.KERNEL_IDS["matern52"][0]
def timed(fn, reps=10):
    for _ in range(3): out = fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps, out
ms_val, m = timed(lambda: bgp.batch_eval_device(kid, None, cands.theta, gi.X_d, y, 1e-6, bgp.METRIC_VAL_RMSE, Xv, yv, 0.0, 1.0))
th = to_device(bgp._theta_block(P.restart_population("matern52", 4096)), dev)
ms_nll, m2 = timed(lambda: bgp.batch_eval_device(kid, None, th, gi.X_d, y, 1e-6, bgp.METRIC_NLL))
y2 = to_device(rng.standard_normal(nd), dev); yv2 = to_device(rng.standard_normal(150), dev)
ms_pair, (pa, pb) = timed(lambda: bgp.batch_eval_pair_device(kid, None, cands.theta, gi.X_d, y, y2, 1e-6, bgp.METRIC_VAL_RMSE, Xv, yv, yv2))
line = "pair val-RMSE 900 x 2: direct %.4f ms" % ms_pair
if cands.shape_of is not None:
    ms_sh, (sa, sb) = timed(lambda: bgp.batch_eval_shared_device(kid, cands.theta, cands.shape_of, cands.shape_theta, gi.X_d, y, y2, 1e-6, bgp.METRIC_VAL_RMSE, Xv, yv, yv2))
    ms_sh_nll, sn = timed(lambda: bgp.batch_eval_shared_device(kid, cands.theta, cands.shape_of, cands.shape_theta, gi.X_d, y, None, 1e-6, bgp.METRIC_NLL))
    ms_d_nll, dn = timed(lambda: bgp.batch_eval_device(kid, None, cands.theta, gi.X_d, y, 1e-6, bgp.METRIC_NLL))
    line += "   shared shapes (%d) %.4f ms, identical bits: %s   NLL 900: direct %.4f shared %.4f identical: %s" % (
        cands.shape_theta.shape[0], ms_sh, bool(torch.equal(pa, sa) and torch.equal(pb, sb)), ms_d_nll, ms_sh_nll, bool(torch.equal(sn, dn)))
print(line)
print(os.environ.get("B200ID_LIB", "default").split("/")[-1], "val-RMSE 900: %.4f ms   NLL 4096: %.4f ms   checksums %.12e %.12e" % (ms_val, ms_nll, float(m[torch.isfinite(m)].sum()), float(m2[torch.isfinite(m2) & (m2 < 1e9)].sum())))
