#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun with N >= 2 ranks, one GPU each):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_check.py

Compares, against the single-GPU result computed on every rank: (i) one record time-sharded over ranks
(FRF partial sums all-reduced), (ii) one record's FIR validation sharded by output range, (iii) a
candidate search sharded round-robin with the argmin gather.  Exit code 0 = all ranks agree."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import dist as bdist, frf, fir, gp, pipeline as P, synth     # noqa: E402
from b200id.runtime import require_cuda, to_device                        # noqa: E402

rank, world, local = bdist.init_from_env()
dev = require_cuda()
n, nd = 1_000_003, 100
t, u, y = synth.make_record(n, nd, "multisine", seed=0)
w = P.log_grid_omega(nd)
w_d = to_device(w, dev)
span = float(t[-1] - t[0])
# (i) FRF of one record, time-sharded
Uref, Yref = frf.record_coefficients_device(to_device(t, dev), to_device(u, dev), to_device(y, dev), w_d, span,
                                            t0=float(t[0]), w_max=float(w.max()))
lo, hi = bdist.shard_contiguous(n, rank, world)
a, b = max(lo - 1, 0), min(hi + 1, n)
U, Y = frf.record_coefficients_sharded(to_device(t[a:b], dev), to_device(u[a:b], dev), to_device(y[a:b], dev), w_d,
                                       lo, hi, n, span, float(w.max()))
e_frf = max(float(((U - Uref).abs() / Uref.abs()).max()), float(((Y - Yref).abs() / Yref.abs()).max()))
# (ii) FIR validation, output-sharded with history
g = np.random.default_rng(3).standard_normal(1024) / 1024
g_d = to_device(g, dev)
yy = np.convolve(u, g)[:n] + 0.01 * np.random.default_rng(4).standard_normal(n)
skip = 10
ref_sums, _ = fir.fir_eval_device(to_device(u, dev), to_device(yy, dev), g_d, skip, float(yy[skip]))
ref = fir.metrics_from_sums(ref_sums.cpu().numpy(), skip)
lead = min(lo, 1023)
got = fir.fir_validate_sharded(to_device(u[lo - lead:hi], dev), to_device(yy[lo:hi], dev), g_d, lo, hi, float(yy[skip]))
e_fir = max(abs(got[k] - ref[k]) / abs(ref[k]) for k in ("rmse", "fit_percent", "r2"))
# (iii) candidate search
X = np.linspace(-1.7, 1.7, nd); yt = np.sin(2 * X)
cands = P.grid_candidates("matern52")
full = P.CandidateSet.build("matern52", cands, dev)
shard = P.CandidateSet.build("matern52", cands, dev, rank=rank, world_size=world)
X_d, y_d = to_device(X, dev), to_device(yt, dev)
pend = P.select_candidates_async(full, X_d, y_d, 1e-6).cpu()
i_ref, m_ref = int(pend[1]), float(pend[0])
pend = P.select_candidates_async(shard, X_d, y_d, 1e-6).cpu()
i_got, m_got = P._finish_selection(shard, pend, dev)
ok = e_frf < 1e-10 and e_fir < 1e-11 and i_got == i_ref and m_got == m_ref
print(f"rank {rank}/{world}: frf {e_frf:.2e} fir {e_fir:.2e} argmin {i_got} vs {i_ref} -> {'ok' if ok else 'MISMATCH'}", flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
torch.distributed.all_reduce(flag) if world > 1 else None
torch.distributed.destroy_process_group() if world > 1 else None
sys.exit(int(flag.item() != 0))
