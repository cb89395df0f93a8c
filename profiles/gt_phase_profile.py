#!/usr/bin/env python3
"""Phase timing of gp_batch_tiled_kernel (needs a -DGT_PROFILE build: B200ID_NVCC_FLAGS=-DGT_PROFILE)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp, pipeline as P, _lib
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
lib = _lib.load()
n, cands = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 8192
X = to_device(np.linspace(-1.7, 1.7, n), dev)
y = torch.randn(n, dtype=torch.float64, device=dev)
th = to_device(gp._theta_block(P.restart_population("matern52", cands)), dev)
gp.batch_eval_device(3, None, th, X, y, 1e-6); torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
lib.b200id_debug_gt_profile(buf, 1)
gp.batch_eval_device(3, None, th, X, y, 1e-6); torch.cuda.synchronize()
lib.b200id_debug_gt_profile(buf, 0)
names = ["fill", "sync(fill)", "update_finish", "sync(update)", "diag", "lookahead", "sync(diag)", "panel", "sync(panel)"]
tot = sum(buf[:9])
for k, nm in enumerate(names):
    print(f"{nm:16s} {buf[k] / (cands * 4):10.0f} cycles per warp per candidate  {buf[k] / tot * 100:5.1f}%")
print("total per warp per candidate", tot / (cands * 4))
