#!/usr/bin/env python3
"""Timings of BASELINE.json configs 3, 4 and (per-GPU share of) 5 on one B200, CUDA events, inputs resident in HBM.

    python profiles/run_configs.py > profiles/rNN_configs.json

cfg 3: all 11 baseline kernels x 4096 restart start points at N_d = 100, NLL, both targets (Re, Im): 90,112
       candidate evaluations as two mixed-kernel launches (one per target).
cfg 4: N_d = 2048: blocked FP64 Cholesky alone (b200id_chol_blocked), one candidate through fit (Gram, factor,
       alpha, K^-1) and a few candidates through the batched NLL; FIR validation on a 10^7-sample square-wave record.
cfg 5: one rank's share of the 10^9-sample multi-record set (R = 100 records x 10^7 over 8 GPUs -> 13 records):
       per-record coefficients + cross-power sums (device-resident record re-used; HBM holds 4 records at a time in
       the real run), and the FIR validation of the same samples.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
from b200id import _lib, frf, fir, gp, pipeline as P, synth   # noqa: E402
from b200id.runtime import require_cuda, to_device, empty     # noqa: E402
import ctypes as C                                             # noqa: E402

dev = require_cuda()
out = {}


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); r = fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best, r


def fp64_peak():
    sink = empty(8, dev)
    best = 0.0
    for mode in (0, 1):
        fl = C.c_double(0.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.call("b200id_fp64_peak_probe", C.c_int(mode), C.c_int(2000), C.c_void_p(sink.data_ptr()), C.byref(fl), None)
        s.record()
        _lib.call("b200id_fp64_peak_probe", C.c_int(mode), C.c_int(2000), C.c_void_p(sink.data_ptr()), C.byref(fl), None)
        e.record(); torch.cuda.synchronize()
        best = max(best, fl.value / (s.elapsed_time(e) * 1e-3) / 1e12)
    return best


peak = fp64_peak()
out["fp64_peak_tflops_measured"] = peak

# ------------------------------------------------------------------ cfg 3
nd, n_rest = 100, 4096
t, u, y = synth.make_record(400_000, nd, "multisine", seed=3)
_, w = synth.log_grid(nd)
U, Y = frf.demod_pair(t, u, y, w)
G = frf.cross_power([U], [Y])
X = np.log10(w); X = (X - X.mean()) / X.std()
targets = []
for v in (G.real, G.imag):
    targets.append((v - v.mean()) / v.std())
theta, kids, names = P.restart_candidates(n_rest)
X_d = to_device(X, dev)
cs = P.CandidateSet.build(None, theta, dev, kernel_ids=kids)
ys = [to_device(v, dev) for v in targets]


def cfg3():
    return [gp.batch_eval_device(0, cs.kernel_ids, cs.theta, X_d, yv, 1e-6, gp.METRIC_NLL) for yv in ys]


ms, m = timed(cfg3)
ncand = 2 * theta.shape[0]
flops = ncand * (nd ** 3 / 3 + 2 * nd ** 2)
out["cfg3"] = {"kernels": 11, "restarts_per_kernel": n_rest, "n_d": nd, "targets": 2, "candidate_evaluations": ncand,
               "ms": ms, "candidates_per_s": ncand / (ms * 1e-3), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
               "fp64_frac": flops / (ms * 1e-3) / 1e12 / peak,
               "finite_metrics": int(sum(int(torch.isfinite(v).sum()) for v in m))}
per_kernel = {}
for k, name in enumerate(P.BASELINE_KERNELS):
    th_k = to_device(theta[k * n_rest:(k + 1) * n_rest], dev)
    kid = gp.KERNEL_IDS[name][0]
    ms_k, _ = timed(lambda: gp.batch_eval_device(kid, None, th_k, X_d, ys[0], 1e-6, gp.METRIC_NLL), reps=2)
    per_kernel[name] = {"ms": ms_k, "candidates_per_s": n_rest / (ms_k * 1e-3)}
out["cfg3"]["per_kernel_single_target"] = per_kernel

# ------------------------------------------------------------------ cfg 4
n4 = 2048
_, w4 = synth.log_grid(n4)
X4 = np.log10(w4); X4 = (X4 - X4.mean()) / X4.std()
rng = np.random.default_rng(4)
y4 = np.sin(2.0 * X4) + 0.05 * rng.standard_normal(n4)
X4_d, y4_d = to_device(X4, dev), to_device(y4, dev)
K = gp.gram(gp.KERNEL_IDS["rbf"][0], np.array([1.0, 0.5]), X4) + 1e-3 * np.eye(n4)
K_d = to_device(K, dev)


def chol_only():
    A = K_d.clone()
    return gp.cholesky_blocked_device(A)


ms_clone, _ = timed(lambda: K_d.clone())
ms_chol, info = timed(chol_only)
ms_chol -= ms_clone
fl_chol = n4 ** 3 / 3
NC4 = 16
th4 = to_device(gp._theta_block(np.column_stack([np.linspace(0.5, 2.0, NC4), np.linspace(0.3, 1.0, NC4)])), dev)
ms_b, m4 = timed(lambda: gp.batch_eval_device(gp.KERNEL_IDS["rbf"][0], None, th4, X4_d, y4_d, 1e-3, gp.METRIC_NLL), reps=2)
fl_b = NC4 * (n4 ** 3 / 3 + 2 * n4 ** 2)
ms_1, _ = timed(lambda: gp.batch_eval_device(gp.KERNEL_IDS["rbf"][0], None, th4[:1].contiguous(), X4_d, y4_d, 1e-3, gp.METRIC_NLL), reps=2)
NC4b = 64
th4b = to_device(gp._theta_block(np.column_stack([np.linspace(0.5, 2.0, NC4b), np.linspace(0.3, 1.0, NC4b)])), dev)
ms_b64, _ = timed(lambda: gp.batch_eval_device(gp.KERNEL_IDS["rbf"][0], None, th4b, X4_d, y4_d, 1e-3, gp.METRIC_NLL), reps=2)
ms_fit, _ = timed(lambda: P.fit_device("rbf", np.array([1.0, 0.5]), X4_d, y4_d, 1e-3, want_kinv=True), reps=2)
out["cfg4"] = {"n_d": n4,
               "chol_blocked": {"ms": ms_chol, "info": int(info), "tflops": fl_chol / (ms_chol * 1e-3) / 1e12, "fp64_frac": fl_chol / (ms_chol * 1e-3) / 1e12 / peak},
               "batch_nll": {"candidates": NC4, "ms": ms_b, "candidates_per_s": NC4 / (ms_b * 1e-3), "algorithmic_tflops": fl_b / (ms_b * 1e-3) / 1e12,
                             "fp64_frac": fl_b / (ms_b * 1e-3) / 1e12 / peak, "single_candidate_ms": ms_1,
                             "note": "up to 16 candidates advance in lockstep (every launch of the factorisation chain carries all of them); the right-hand side rides along the factorisation",
                             "candidates_64": {"ms": ms_b64, "candidates_per_s": NC4b / (ms_b64 * 1e-3),
                                               "fp64_frac": NC4b * (n4 ** 3 / 3 + 2 * n4 ** 2) / (ms_b64 * 1e-3) / 1e12 / peak},
                             "metrics": [float(v) for v in m4.cpu()]},
               "fit_with_kinv": {"ms": ms_fit, "flops_note": "Gram + n^3/3 factor + 2 n^2 solves + n^3 (K^-1 = L^-T L^-1)"}}
n = 10_000_000
tv, uv, yv = synth.make_record(n, 100, "square", seed=1)
uv_d, yv_d = to_device(uv, dev), to_device(yv, dev)
g = np.exp(-np.arange(1024) / 200.0) * np.sin(np.arange(1024) / 15.0) * 1e-2
g_d = to_device(g, dev)
ms_fir, sums = timed(lambda: fir.fir_eval_device(uv_d, yv_d, g_d, 10, float(yv[10])))
out["cfg4"]["fir_square_wave"] = {"n": n, "taps": 1024, "ms": ms_fir, "samples_per_s": n / (ms_fir * 1e-3),
                                  "hbm_gbs": 16.0 * n / (ms_fir * 1e-3) / 1e9}

# ------------------------------------------------------------------ cfg 5 (one rank's share)
R_total, world = 100, 8
R_rank = (R_total + world - 1) // world
t5, u5, y5 = synth.make_record(n, 100, "multisine", seed=50)
t5_d, u5_d, y5_d = to_device(t5, dev), to_device(u5, dev), to_device(y5, dev)
w_d = to_device(w, dev)
span = float(t5[-1] - t5[0])
path = frf.choose_path(t5_d, span, float(w.max()), float(t5[0]))
sums_all = empty(3 * nd, dev)


def cfg5_frf():
    for r in range(R_rank):
        Ur, Yr = frf.record_coefficients_device(t5_d, u5_d, y5_d, w_d, span, path=path)
        frf.cross_power_sums_device(Ur, Yr, sums_all, accumulate=r > 0)
    return sums_all


def cfg5_fir():
    acc = None
    for r in range(R_rank):
        s = fir.fir_eval_device(u5_d, y5_d, g_d, 10, float(y5[10]))
        acc = s if acc is None else acc + s
    return acc


ms_frf, _ = timed(cfg5_frf, reps=2)
ms_fir5, _ = timed(cfg5_fir, reps=2)
samples = R_rank * n
out["cfg5_one_rank_share"] = {"records_per_rank": R_rank, "samples_per_rank": samples, "n_d": nd, "frf_path": path[0],
                              "frf_ms": ms_frf, "frf_samples_per_s": samples / (ms_frf * 1e-3), "frf_hbm_gbs": 24.0 * samples / (ms_frf * 1e-3) / 1e9,
                              "fir_ms": ms_fir5, "fir_samples_per_s": samples / (ms_fir5 * 1e-3), "fir_hbm_gbs": 16.0 * samples / (ms_fir5 * 1e-3) / 1e9,
                              "projected_8gpu_frf_plus_fir_s_for_1e9_samples": (ms_frf + ms_fir5) * 1e-3,
                              "note": "records are independent: the 8-GPU run adds one 2.4 KB all-reduce (see bench.py --gpus 8)"}
print(json.dumps(out, indent=1))
