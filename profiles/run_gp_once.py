#!/usr/bin/env python3
"""Two launches of gp_batch_tiled_kernel for ncu (`-k regex:gp_batch_tiled|gp_shape_table`): the bench's search (900 Matern-5/2
candidates x 2 targets, validation RMSE, through the shared covariance shapes) and 4096 NLL restart points (direct)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
import numpy as np, torch
from b200id import gp as bgp, pipeline as P
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
nd = 100
gi = P.GridInputs.get(nd, dev)
rng = np.random.default_rng(0)
y, y2 = to_device(rng.standard_normal(nd), dev), to_device(rng.standard_normal(nd), dev)
Xv = to_device(np.linspace(-1.6, 1.6, 150), dev)
yv, yv2 = to_device(rng.standard_normal(150), dev), to_device(rng.standard_normal(150), dev)
cs = P.CandidateSet.build("matern52", P.grid_candidates("matern52"), dev)
th = to_device(bgp._theta_block(P.restart_population("matern52", 4096)), dev)
for _ in range(2):
    bgp.batch_eval_shared_device(3, cs.theta, cs.shape_of, cs.shape_theta, gi.X_d, y, y2, 1e-6, bgp.METRIC_VAL_RMSE, Xv, yv, yv2)
    bgp.batch_eval_device(3, None, th, gi.X_d, y, 1e-6, bgp.METRIC_NLL)
torch.cuda.synchronize()
print("ok")
