"""One bin-adaptive projection call of the multisine record (10^7 samples, N_d = 100) for `ncu -k regex:...`:
    python profiles/frf_adaptive_once.py [all] [square]   # all = every bin listed; square = the 150-bin validation record"""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "gauss-process_b200")]
import numpy as np, torch
from b200id import _lib, frf, synth
from b200id.runtime import require_cuda, to_device
dev = require_cuda(); lib = _lib.load()
n = 10_000_000; nd = 150 if "square" in sys.argv else 100
t, u, y = synth.make_record(n, 100, "square" if "square" in sys.argv else "multisine", seed=1)
_, w = synth.log_grid(nd)
t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
span = float(t[-1] - t[0]); t0, dt = float(t[0]), span / (n - 1)
sums = frf.moments_device(u_d, y_d, (0, n))
lib.b200id_frf_select_recurrence(C.c_int(2)); lib.b200id_frf_adaptive_force(C.c_int(1 if "all" in sys.argv else 0))
for _ in range(2):
    p = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
torch.cuda.synchronize()
print("bins listed:", frf.adaptive_diag(nd, dev)[0])
