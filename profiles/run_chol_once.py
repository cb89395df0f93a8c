#!/usr/bin/env python3
"""One b200id_chol_blocked call at n = 2048 (target of the ncu launch list)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import gp
from b200id.runtime import require_cuda, to_device
dev = require_cuda()
n = 2048
x = np.linspace(-1.7, 1.7, n)
K = gp.gram(gp.KERNEL_IDS["rbf"][0], np.array([1.0, 0.5]), x) + 1e-3 * np.eye(n)
A = to_device(K, dev)
print(gp.cholesky_blocked_device(A))
