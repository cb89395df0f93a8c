#!/usr/bin/env python3
"""Time b200id_chol_blocked at n = 2048 (BASELINE config 4) and smaller, check L L^T = A."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
for n in (512, 1024, 2048, 4096):
    x = np.linspace(-1.7, 1.7, n)
    K = gp.gram(gp.KERNEL_IDS["rbf"][0], np.array([1.0, 0.5]), x) + 1e-3 * np.eye(n)
    K_d = to_device(K, dev)
    best = 1e9
    for _ in range(4):
        A = K_d.clone()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); info = gp.cholesky_blocked_device(A); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    L = torch.tril(A)
    err = float((L @ L.T - K_d).abs().max())
    ref = torch.linalg.cholesky(K_d)
    print(f"n={n}: {best:.3f} ms  {n**3/3/(best*1e-3)/1e12:.2f} TFLOP/s  info={info}  |LL^T-A|max={err:.2e}  |L-Lref|max={float((L-ref).abs().max()):.2e}")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record(); torch.linalg.cholesky(K_d); e.record(); torch.cuda.synchronize()
    print(f"      (torch.linalg.cholesky / cuSOLVER for scale: {s.elapsed_time(e):.3f} ms)")
