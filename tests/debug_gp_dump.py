"""Scratch: dump GPU candidate metrics for kernels whose failure pattern differs (run under gpurun)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp
g = np.load(ROOT / "tests/golden/gp.npz")
out = {}
for name in ("ss1", "sshf", "stable_spline", "dc", "di", "matern52", "ss2"):
    C = g[f"cand_{name}"]
    out[f"nll_{name}"] = gp.batch_eval(name, C, g["X"], g["yr"], 1e-6)
    out[f"nll0_{name}"] = gp.batch_eval(name, C, g["X"], g["yr"], 0.0)
    out[f"vrmse_{name}"] = gp.batch_eval(name, C, g["X"], g["yr"], 1e-6, Xv=g["Xv"], yv=g["yv_r"],
                                         y_mean=float(g["yr_mean"][0]), y_scale=float(g["yr_scale"][0]))
np.savez(ROOT / "gpurun_out" / "gp_dump.npz", **out)
print("dumped")
