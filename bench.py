#!/usr/bin/env python3
"""bench.py -- time-samples/s of the FRF -> GPR-selection -> FIR hot path on N B200s (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input (BASELINE.json configs[1],
scaled weakly over ranks as configs[4] does): per rank one 10^7-sample multisine training record and
one 10^7-sample square-wave validation record (dt = 2 ms, FP64), N_d = 100:
  FRF of the training record (100 bins)  ->  cross-power (all-reduce of per-bin num/den over ranks)
  150-bin FRF of the validation record   ->  grid-search targets (validation-RMSE metric)
  Matern-5/2 grid search, 900 candidates x 2 targets (candidates sharded over ranks, argmin gather)
  final fit x2, GP mean on the 1000-point uniform grid, Hermitian IDFT -> 1024 taps
  FIR validation of the taps over the validation record (all-reduce of the fused error sums)
value = training samples processed by all ranks / step time (device-resident inputs, CUDA events,
max over ranks).  e2e = same metric with HOST (pinned) buffers: H2D of both records and D2H of the
result inside the timed region.  See DESIGN.md section "Measurement" for the roofline arithmetic.

Besides the headline fields the line carries
  parity   -- the step's own results against the CPU oracle on the bench's own inputs (outside the timed region)
  configs  -- BASELINE configs 3 (11 kernels x 4096 restarts, sharded over the ranks), 4 (N_d = 2048, 16 candidates)
              and, with more than one rank, 5 (R = 100 records x 10^7 samples round-robin over the ranks + FIR)
  e2e_full_upload / e2e_pageable -- the same call with the time vectors uploaded every step / from pageable memory

--impl reference times the reference's CPU algorithm (oracle port) on the SAME records at the SAME n; when the run
would not fit the time budget (B200ID_REF_BUDGET_S, default 240 s) the O(n) stages (FRF, FIR) run on a slice and are
scaled linearly while the GP stage is timed in full -- the line says which (config.n_samples_per_record_run).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "gauss-process_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "time-samples/s (FRF+GPR-selection+FIR pipeline)"
UNIT = "samples/s"
KERNEL = "matern52"
NOISE = 1e-6
N_TAPS = 1024
VAL_ND = 150
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
# (profiles/r01_ncu_*.txt), 10^7-sample record: FRF reads t,u,y once (240 MB), FIR reads u,y (160 MB)
NCU_TRAFFIC_BYTES = {"frf_main_ws_kernel": 240.1e6, "frf_main_dmma_kernel": 240.0e6, "frf_round_goertzel_kernel": 240.1e6, "frf_timebase_fit_kernel": 20.0e6,
                     "frf_project_uniform_g_kernel": 245.9e6, "fir_fft16_kernel": 167.5e6, "frf_project_uniform3_kernel": 245.2e6, "frf_project_uniform_kernel": 245.2e6, "frf_project_kernel": 245.3e6, "fir_eval_kernel": 167.3e6,
                     "fir_fft_kernel": 163.6e6, "gp_batch_tiled_kernel": 0.33e6}
FRF_BYTES_PER_SAMPLE = 24        # t, u, y FP64 read once
FIR_BYTES_PER_SAMPLE = 16        # u, y FP64 read once, y_pred not materialised


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000, help="samples per record (per rank)")
    ap.add_argument("--nd", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="samples per record of the CPU-baseline pass (0 = the bench's own n, one pass ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block (cfg 3 / 4 / 5)")
    ap.add_argument("--resident-only", action="store_true",
                    help="profiling aid (ncu launch lists): only the device-resident step, the same launches the "
                         "headline 'value' times; prints a reduced line, not a bench line")
    ap.add_argument("--cfg5", default="auto", choices=["auto", "on", "off"],
                    help="config 5 (100 records x 1e7 samples over the ranks): auto = when WORLD_SIZE > 1")
    ap.add_argument("--cfg5-records", type=int, default=100)
    ap.add_argument("--restarts", type=int, default=4096, help="config 3: start points per kernel")
    return ap.parse_args()


def workload_config(a, world, n_run=None):
    n_run = a.n if n_run is None else n_run
    return {
        "workload": f"cfg2 full pipeline per rank: {a.n}-sample multisine train record + {a.n}-sample square-wave "
                    f"validation record, N_d={a.nd}, validation FRF {VAL_ND} bins, Matern-5/2 grid 900 candidates x 2 "
                    f"targets (val-RMSE), 1000-pt GP mean, Hermitian IDFT -> {N_TAPS} taps, FIR validation",
        "n_samples_per_record": a.n, "n_samples_per_record_run": n_run, "n_d": a.nd, "records_per_rank": 2, "ranks": world,
        "gp_candidates_per_step": 1800, "fir_taps": N_TAPS, "noise_variance": NOISE,
        "l2_policy": "inputs larger than L2 (2 records x 240 MB per rank vs 126 MB L2)",
        "parallelism": f"records x{world} (weak), candidates round-robin over ranks",
    }


# ------------------------------------------------------------------------------------------------
# synthetic inputs (outside every timed region)
def make_records(n, nd, seed, dev=None):
    """(train, val) host records.  The multisine is summed on the GPU when one is present (10^9 sines
    at n = 10^7 take ~15 s in NumPy); the plant is scipy.signal.lfilter on the host either way."""
    from b200id import synth
    if dev is None or n < 1_000_000:
        return synth.make_record(n, nd, "multisine", seed=seed), synth.make_record(n, nd, "square", seed=seed + 1)
    import torch
    rng = np.random.default_rng(seed)
    f, w = synth.log_grid(nd)
    amp = torch.from_numpy(rng.uniform(0.0, 20.0 / nd, size=nd)).to(dev)
    ph = torch.from_numpy(rng.uniform(0.0, 2.0 * np.pi, size=nd)).to(dev)
    w_d = torch.from_numpy(w).to(dev)
    t = np.arange(n, dtype=np.float64) * synth.DT
    u = np.empty(n)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        tc = torch.from_numpy(t[s:e]).to(dev)
        u[s:e] = (torch.sin(tc[:, None] * w_d[None, :] + ph[None, :]) @ amp).cpu().numpy()
    y = synth.apply_plant(u, seed=seed + 7)
    tv, uv = synth.square_wave(n, seed=seed + 1)
    yv = synth.apply_plant(uv, seed=seed + 8)
    return (t, u, y), (tv, uv, yv)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_pipeline_pass(train, val, nd, threads):
    """One pass of the SAME pipeline through the NumPy oracle (the reference's algorithm, oracle/*.py) on the given
    host records.  Returns (seconds per stage, results): the results are what the parity block compares with."""
    from oracle import frf as ofrf, gp as ogp, fir as ofir
    (t, u, y), (tv, uv, yv) = train, val
    st = {}
    t0 = time.perf_counter()
    w, G = ofrf.estimate_frf([(t, u, y)], nd, threads=threads)
    st["frf_train"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    wv, Gv = ofrf.estimate_frf([(tv, uv, yv)], VAL_ND, threads=threads)
    st["frf_val"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    X, xm, xs = ogp.standardize(np.log10(w))
    Xv = (np.log10(wv) - xm) / xs
    fits = []
    with np.errstate(all="ignore"):
        for target, tv_ in ((G.real, Gv.real), (G.imag, Gv.imag)):
            z, m, s = ogp.standardize(target)
            best, best_metric, _, _ = ogp.grid_search(KERNEL, X, z, NOISE, 5000, Xv=Xv, yv=tv_, y_mean=m, y_scale=s)
            alpha, Kinv, _ = ogp.fit(KERNEL, best, X, z, NOISE)
            fits.append((best, alpha, Kinv, m, s, best_metric))
    st["gp"] = time.perf_counter() - t0
    t0 = time.perf_counter()

    def predict(w_new):
        Xn = (np.log10(w_new) - xm) / xs
        parts = [ogp.predict(KERNEL, b, X, a, K, Xn) * s + m for (b, a, K, m, s, _) in fits]
        return parts[0] + 1j * parts[1]

    taps, *_ = ofir.paper_fir(w, G, predict=predict)
    metrics, _ = ofir.fir_validate(uv, yv, taps)
    st["fir"] = time.perf_counter() - t0
    res = {"omega": w, "G": G, "Gv": Gv, "theta_real": np.asarray(fits[0][0]), "theta_imag": np.asarray(fits[1][0]),
           "metric_real": float(fits[0][5]), "metric_imag": float(fits[1][5]), "taps": taps, "fir": metrics}
    return st, res


def scaled_step_seconds(st, n_run, n_full):
    """Step time at n_full from stage seconds measured at n_run: FRF and FIR are O(n) (frf_estimator.py:226-231,
    fir_fitting.py:285), the GP stage does not depend on n."""
    f = n_full / float(n_run)
    return (st["frf_train"] + st["frf_val"] + st["fir"]) * f + st["gp"]


def slice_records(train, val, n_s):
    return tuple(a[:n_s] for a in train), tuple(a[:n_s] for a in val)


def run_reference_arm(a, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is pure Python and
    /root/reference does not exist on the GPU box) on the host cores with all the threads the per-bin loop can use,
    on the bench's own records.  Runs at the bench's own n when (steps + warmup) passes fit the budget; otherwise
    the O(n) stages run on a slice and are scaled (the line carries the n actually run and the stage seconds)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    dev = None
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", 0)       # only to synthesise the multisine quickly; nothing timed runs there
    except Exception:
        dev = None
    train, val = make_records(a.n, a.nd, seed=0, dev=dev)
    budget = float(os.environ.get("B200ID_REF_BUDGET_S", "240"))
    # calibration on a small slice: seconds per sample of the O(n) stages and the fixed GP cost
    n_cal = min(a.n, 200_000)
    st_cal, _ = cpu_pipeline_pass(*slice_records(train, val, n_cal), a.nd, threads)
    passes = a.warmup + a.steps
    est_full = scaled_step_seconds(st_cal, n_cal, a.n)
    if passes * est_full <= budget:
        n_s = a.n
    else:
        per_sample = (st_cal["frf_train"] + st_cal["frf_val"] + st_cal["fir"]) / n_cal
        n_s = int(max(n_cal, min(a.n, (budget / passes - st_cal["gp"]) / per_sample)))
    tr_s, va_s = slice_records(train, val, n_s)
    times, stages = [], []
    for i in range(passes):
        st, _ = cpu_pipeline_pass(tr_s, va_s, a.nd, threads)
        if i >= a.warmup:
            times.append(scaled_step_seconds(st, n_s, a.n))
            stages.append(st)
    sec = float(np.mean(times))
    value = a.n / sec
    stage_mean = {k: float(np.mean([st[k] for st in stages])) for k in stages[0]}
    sample = (f"{n_s} of {a.n} samples per record actually run per step"
              + ("" if n_s == a.n else "; FRF and FIR seconds scaled by n/n_run, GP stage as measured")
              + f"; same bins, same 1800 GP candidates, same {N_TAPS}-tap FIR; per-bin FRF loop on a {threads}-thread pool; "
              f"stage seconds at n_run {json.dumps({k: round(v, 4) for k, v in stage_mean.items()})}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(a, 1, n_run=n_s),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "n_samples_run": n_s, "stage_seconds_at_n_run": stage_mean,
                         "note": "the oracle port spreads the reference's single-threaded per-bin loop over all host cores "
                                 "(conservative: the reference itself uses one core outside BLAS)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def measure_frf_phases(dev_recs, a, dev):
    """{frf_main_ms, frf_main_tflops, frf_round_ms, bins listed} of one step's two b200id_frf_project_uniform calls."""
    import ctypes as C
    import torch
    from b200id import _lib, frf as bfrf, pipeline as P
    from b200id.runtime import to_device
    lib = _lib.load()
    out = {}
    recs = [("train", dev_recs[0], a.nd), ("val", dev_recs[1], VAL_ND)]
    tot = {"full": 0.0, "main": 0.0}
    listed = {}
    for tag, rec, nd in recs:
        t_d, u_d, y_d = rec[0], rec[1], rec[2]
        n = t_d.numel()
        w_d = to_device(P.log_grid_omega(nd), dev)
        t0 = float(t_d[0].item()); span = float(t_d[-1].item()) - t0
        dt = span / (n - 1)
        sums = bfrf.moments_device(u_d, y_d, (0, n))
        for mode, key in ((0, "full"), (2, "main")):
            prev = lib.b200id_frf_adaptive_force(C.c_int(mode))
            try:
                for _ in range(2):
                    bfrf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(5):
                    bfrf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
                e.record()
                torch.cuda.synchronize()
                tot[key] += s.elapsed_time(e) / 5
                if mode == 0:
                    listed[tag] = bfrf.adaptive_diag(nd, dev)[0]
            finally:
                lib.b200id_frf_adaptive_force(C.c_int(prev))
    flops = 8.0 * a.n * (a.nd + VAL_ND)
    out["frf_main_ms"] = tot["main"]
    out["frf_main_tflops"] = flops / (tot["main"] * 1e-3) / 1e12
    out["frf_round_ms"] = tot["full"] - tot["main"]
    out["frf_bins_with_rounding_term"] = listed
    return out


def run_configs(a, rank, world, dev, res, timed, hbm_peak):
    """BASELINE configs 3, 4 and 5 measured in the same run (every rank takes part; CUDA events, max over ranks).
    Returns the `configs` block of the line (fp64 fractions are filled in by the caller, which owns the peak probe)."""
    import torch
    from b200id import fir as bfir, frf as bfrf, gp as bgp, pipeline as P, synth
    from b200id.runtime import empty, to_device
    out = {}
    gi = P.GridInputs.get(a.nd, dev)
    # ---- cfg 3: 11 kernels x R restart start points x 2 targets at N_d = nd, NLL, candidates round-robin over ranks
    Gh = res.G.cpu().numpy()
    targets = []
    for v in (Gh.real, Gh.imag):
        sd = v.std()
        targets.append(to_device((v - v.mean()) / (sd if sd != 0 else 1.0), dev))
    pop = P.Population.build(a.restarts, dev, rank=rank, world_size=world)
    ms3, picks, _, _ = timed(lambda: P.population_search_async(pop, gi.X_d, targets, NOISE), 3, 2)
    picks = picks.cpu().numpy()                                   # [2, n_kernels, 5]
    n_eval = 2 * pop.n_total
    best_k = [P.select_best(pop.names, picks[j, :, 4].tolist()) for j in range(2)]
    out["cfg3"] = {
        "workload": f"{len(pop.names)} kernels x {a.restarts} restart start points (gpr_fitting.py:132-145, np.random.seed(42) per "
                    f"kernel) x 2 targets at N_d={a.nd}, NLL, noise {NOISE}; round-robin over {world} rank(s), one all-gather",
        "candidate_evaluations": n_eval, "ms": ms3, "candidates_per_s": n_eval / (ms3 * 1e-3),
        "algorithmic_flops": n_eval * (a.nd ** 3 / 3.0 + 2.0 * a.nd ** 2),
        "per_kernel_best_nll": {t: {k: float(picks[j, i, 4]) for i, k in enumerate(pop.names)} for j, t in enumerate(("real", "imag"))},
        "cross_kernel_pick": {t: {"kernel": pop.names[best_k[j]], "theta": picks[j, best_k[j], :3].tolist(),
                                  "global_index": int(picks[j, best_k[j], 3])} for j, t in enumerate(("real", "imag"))},
    }
    # ---- cfg 4: N_d = 2048, 16 candidates (NLL) in lockstep + one blocked Cholesky; every rank runs a replica
    n4, nc4 = 2048, 16
    _, w4 = synth.log_grid(n4)
    X4 = np.log10(w4); X4 = (X4 - X4.mean()) / X4.std()
    y4 = np.sin(2.0 * X4) + 0.05 * np.random.default_rng(4).standard_normal(n4)
    X4_d, y4_d = to_device(X4, dev), to_device(y4, dev)
    th4 = to_device(bgp._theta_block(np.column_stack([np.linspace(0.5, 2.0, nc4), np.linspace(0.3, 1.0, nc4)])), dev)
    kid = bgp.KERNEL_IDS["rbf"][0]
    ms4, m4, _, _ = timed(lambda: bgp.batch_eval_device(kid, None, th4, X4_d, y4_d, 1e-3, bgp.METRIC_NLL), 3, 2)
    K4 = to_device(bgp.gram(kid, np.array([1.0, 0.5]), X4) + 1e-3 * np.eye(n4), dev)
    A4 = torch.empty_like(K4)
    info4 = torch.zeros(1, dtype=torch.int32, device=dev)
    import ctypes as C
    from b200id import _lib
    from b200id.runtime import Workspace, ptr, stream_ptr
    ws4 = Workspace.get("cholb", _lib.load().b200id_chol_blocked_workspace_bytes(n4), dev)

    def chol_once():
        A4.copy_(K4)
        _lib.call("b200id_chol_blocked", ptr(A4), C.c_int(n4), C.c_int(n4), ptr(info4), ptr(ws4), C.c_size_t(ws4.numel()), stream_ptr())

    ms_copy, _, _, _ = timed(lambda: A4.copy_(K4), 3, 1)
    ms_chol, _, _, _ = timed(chol_once, 3, 2)
    out["cfg4"] = {
        "workload": f"N_d={n4} RBF candidates: {nc4} in lockstep (Gram + blocked FP64 Cholesky with y riding along + NLL), replicas on every rank",
        "batch_candidates": nc4, "batch_ms": ms4, "candidates_per_s_per_gpu": nc4 / (ms4 * 1e-3),
        "algorithmic_flops": nc4 * (n4 ** 3 / 3.0 + 2.0 * n4 ** 2),
        "single_cholesky_ms": ms_chol - ms_copy, "single_cholesky_flops": n4 ** 3 / 3.0, "info": int(info4.item()),
        "nll_first": float(m4[0]),
    }
    # ---- kernel-regularised FIR (kernel_regularized.py): regressor summaries of a 1e5-sample record at
    # L = 128 and the marginal likelihood at the 5 forward-difference points of one L-BFGS-B gradient (DC kernel), device
    # against the oracle's NumPy / LAPACK evaluation of the same points on this host (rank 0 only)
    if rank == 0:
        try:
            import math
            import time as _time
            from b200id import kr as bkr
            from oracle import kr_fir as okr
            n_kr, L_kr = 100_000, 128
            _, u_kr, y_kr = synth.make_record(n_kr, a.nd, "multisine", seed=21)
            u_kr, y_kr = u_kr - u_kr.mean(), y_kr - y_kr.mean()
            def timed_local(fn, steps, warmup):
                # rank 0 alone runs this block: no barrier, no all-reduce (timed() holds both)
                for _ in range(warmup):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = None
                for _ in range(steps):
                    o = fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / steps, o

            ms_sum, summ = timed_local(lambda: bkr.summaries(u_kr, y_kr, L_kr), 3, 1)
            th0 = np.array([0.0, math.log(0.98 / 0.02), math.atanh(0.8), math.log(1e-2)])
            pts = np.vstack([th0] + [th0 + 1e-8 * np.eye(4)[k] for k in range(4)])
            ms_nll, (v_gpu, info_kr) = timed_local(lambda: bkr.nll_batch("dc", pts, summ), 5, 2)
            G_h, b_h = summ.G.cpu().numpy(), summ.b.cpu().numpy()
            t0_ = _time.perf_counter()
            v_cpu = np.array([okr.nll(p_, G_h, b_h, summ.yy, n_kr, L_kr, "dc") for p_ in pts])
            cpu_ms = (_time.perf_counter() - t0_) * 1e3
            t0_ = _time.perf_counter()
            okr.summaries(u_kr, y_kr, L_kr)
            cpu_sum_ms = (_time.perf_counter() - t0_) * 1e3
            out["kr_fir"] = {
                "workload": f"kernel-regularised FIR, DC kernel, L={L_kr}, {n_kr} samples: summaries (host u, y in) and 5-point NLL batch",
                "summaries_ms": ms_sum, "summaries_cpu_ms": cpu_sum_ms, "nll_batch_ms": ms_nll, "nll_batch_cpu_ms": cpu_ms,
                "nll_rel_err_max": float(np.max(np.abs(v_gpu - v_cpu) / np.abs(v_cpu))), "info": info_kr.tolist(),
            }
        except Exception as exc:
            out["kr_fir"] = {"error": repr(exc)}
    # ---- cfg 5: R records x n samples round-robin over the ranks: per-record FRF + cross-power sums, ONE all-reduce;
    # FIR validation of the step's taps over the same samples, one 4-double all-reduce
    want5 = a.cfg5 == "on" or (a.cfg5 == "auto" and world > 1)
    if want5:
        R = a.cfg5_records
        mine = list(range(rank, R, world))
        n = a.n
        t_h = np.arange(n, dtype=np.float64) * synth.DT
        t_d = to_device(t_h, dev)
        span, t0 = float(t_h[-1] - t_h[0]), 0.0
        f, w = synth.log_grid(a.nd)
        w_d = gi.w_d
        # plant as its truncated impulse response (2048 taps) so that the records are synthesised on the device
        imp = np.zeros(2048); imp[0] = 1.0
        g_true = to_device(synth.apply_plant(imp, noise=0.0), dev)
        recs = []
        for r in mine:
            rng = np.random.default_rng(1000 + r)
            amp = to_device(rng.uniform(0.0, 20.0 / a.nd, size=a.nd), dev)
            ph = to_device(rng.uniform(0.0, 2.0 * np.pi, size=a.nd), dev)
            u_d = empty(n, dev)
            chunk = 1 << 20
            for s0 in range(0, n, chunk):
                e0 = min(n, s0 + chunk)
                u_d[s0:e0] = torch.sin(t_d[s0:e0, None] * w_d[None, :] + ph[None, :]) @ amp
            _, y_d = bfir.fir_eval_device(u_d, None, g_true, 0, 0.0, want_pred=True)
            gen = torch.Generator(device=dev); gen.manual_seed(2000 + r)
            y_d = y_d + 1e-3 * torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
            recs.append((u_d, y_d.contiguous()))
        torch.cuda.synchronize()
        path = bfrf.choose_path(t_d, span, gi.w_max, t0)
        sums = empty(3 * a.nd, dev)
        taps = res.taps
        skip = min(10, taps.numel() // 10)
        y0s = [float(y_d[skip].item()) for _, y_d in recs]

        def frf5():
            if not recs:
                sums.zero_()
            for k, (u_d, y_d) in enumerate(recs):
                U, Y = bfrf.record_coefficients_device(t_d, u_d, y_d, w_d, span, path=path)
                bfrf.cross_power_sums_device(U, Y, sums, accumulate=k > 0)
            bdist_allreduce(sums)
            return bfrf.gp_targets_device(sums, a.nd, standardise=False)[0]

        def fir5():
            acc = torch.zeros(4, dtype=torch.float64, device=dev)
            for (u_d, y_d), y0 in zip(recs, y0s):
                sm, _ = bfir.fir_eval_device(u_d, y_d, taps, skip, y0)
                acc += sm
            bdist_allreduce(acc)
            return acc

        from b200id.dist import allreduce_sum_ as bdist_allreduce
        ms_frf5, G5, _, _ = timed(frf5, 2, 1)
        ms_fir5, acc5, _, _ = timed(fir5, 2, 1)
        total = R * n
        G5h = G5.cpu().numpy()
        Gtrue = synth.plant_response(w)
        out["cfg5"] = {
            "workload": f"R={R} records x {n} samples (distinct multisine records, same plant) round-robin over {world} rank(s): "
                        f"per-record FRF at N_d={a.nd} + cross-power sums, one all-reduce of 3 N_d doubles; FIR validation "
                        f"({taps.numel()} taps) of the same samples, one 4-double all-reduce",
            "records": R, "records_on_rank0": len(mine), "total_samples": total, "frf_path": path[0],
            "frf_ms": ms_frf5, "frf_samples_per_s": total / (ms_frf5 * 1e-3),
            "frf_hbm_gbs_per_gpu": FRF_BYTES_PER_SAMPLE * total / world / (ms_frf5 * 1e-3) / 1e9,
            "frf_hbm_frac": FRF_BYTES_PER_SAMPLE * total / world / (ms_frf5 * 1e-3) / 1e9 / hbm_peak,
            "frf_algorithmic_flops": 8.0 * a.nd * total,
            "fir_ms": ms_fir5, "fir_samples_per_s": total / (ms_fir5 * 1e-3),
            "fir_hbm_gbs_per_gpu": FIR_BYTES_PER_SAMPLE * total / world / (ms_fir5 * 1e-3) / 1e9,
            "fir_hbm_frac": FIR_BYTES_PER_SAMPLE * total / world / (ms_fir5 * 1e-3) / 1e9 / hbm_peak,
            "frf_plus_fir_samples_per_s": total / ((ms_frf5 + ms_fir5) * 1e-3),
            "G_vs_continuous_plant_median_rel": float(np.median(np.abs(G5h - Gtrue) / np.abs(Gtrue))),
            "fir_sums": acc5.cpu().tolist(),
        }
        del recs
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    from b200id import _lib, dist as bdist, pipeline as P
    from b200id.runtime import require_cuda, to_device, empty
    import ctypes as C

    rank, world, local = bdist.init_from_env()
    dev = require_cuda()
    # one process per GPU: keep this rank's host buffers on the NUMA node its GPU hangs off (matters for the e2e
    # figure at N > 1, where every rank pulls 480 MB per step over its own link); B200ID_NUMA_BIND=0 switches it off
    numa = bdist.bind_to_gpu_numa_node(local) if os.environ.get("B200ID_NUMA_BIND", "1") != "0" else {"bound": False}
    sm, major, minor = _lib.device_info()
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    # ---- inputs: generated once, kept on the host (pinned) for e2e and on the device for `value`
    train_h, val_h = make_records(a.n, a.nd, seed=100 * rank, dev=dev)
    pinned = [[torch.from_numpy(x).pin_memory() for x in rec] for rec in (train_h, val_h)]
    spans = [float(rec[0][-1] - rec[0][0]) for rec in (train_h, val_h)]
    t0s = [float(rec[0][0]) for rec in (train_h, val_h)]
    skip = min(10, N_TAPS // 10)
    y0 = float(val_h[2][skip])
    dev_recs = [[x.to(dev) for x in rec] for rec in pinned]
    cands = P.CandidateSet.build(KERNEL, P.grid_candidates(KERNEL), dev, rank=rank, world_size=world)

    def step_resident():
        tr = [(dev_recs[0][0], dev_recs[0][1], dev_recs[0][2], spans[0], t0s[0])]
        va = (dev_recs[1][0], dev_recs[1][1], dev_recs[1][2], spans[1], y0, t0s[1])
        return P.identify_device(tr, va, a.nd, KERNEL, NOISE, cands, n_taps=N_TAPS)

    # e2e: the calls a user of the reference makes (public drop-in surface, HOST NumPy buffers in pinned
    # memory): every FRF / FIR call copies its record host->device, every result comes back as NumPy.
    import contextlib, io

    class StandardScaler:
        """sklearn.preprocessing.StandardScaler's arithmetic and attributes (mean_, scale_ with the zero-variance
        rule) in plain NumPy, like the reference arm's oracle (ogp.standardize): the caller-side scaling of
        gp_pipeline.py:130-132 without sklearn's per-call input validation (0.3-0.8 ms a call, 7 calls a step)."""
        with_mean = with_std = True

        def fit(self, a):
            a = np.asarray(a, dtype=np.float64)
            self.mean_ = a.mean(axis=0)
            sd = a.std(axis=0)
            self.scale_ = np.where(sd == 0.0, 1.0, sd)
            return self

        def transform(self, a):
            return (np.asarray(a, dtype=np.float64) - self.mean_) / self.scale_

        def fit_transform(self, a):
            return self.fit(a).transform(a)

        def inverse_transform(self, a):
            return np.asarray(a, dtype=np.float64) * self.scale_ + self.mean_

    from src.frequency_transform.frf_estimator import (matlab_freq_grid, synchronous_coefficients_trapz_pair,
                                                       frf_cross_power_average)
    from src.gpr import create_kernel, GaussianProcessRegressor
    from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline
    from b200id import fir as bfir
    host = [[x.numpy() for x in rec] for rec in pinned]            # NumPy views of the pinned buffers
    _, omega_h = matlab_freq_grid(a.nd, -1.0, 2.3)
    _, omega_v = matlab_freq_grid(VAL_ND, -1.0, 2.3)
    e2e_dir = ROOT / "gpurun_out" / f"bench_e2e_rank{rank}"

    def step_e2e():
        # ONE public call on host buffers (b200id.pipeline.identify_host, the array-level counterpart of the
        # reference's run_gp_pipeline): each record crosses the link once, uploads queued back to back and
        # overlapped with the projections, the FIR pass reads the resident validation record; the taps and the
        # metrics come back to the host.
        r = P.identify_host([tuple(host[0])], tuple(host[1]), a.nd, KERNEL, NOISE, cands, n_taps=N_TAPS)
        return r.fir, r.taps.cpu().numpy(), r

    def step_e2e_full_upload():
        # the same call with the per-array time-base facts switched off: t goes up with u and y every step
        r = P.identify_host([tuple(host[0])], tuple(host[1]), a.nd, KERNEL, NOISE, cands, n_taps=N_TAPS, reuse_timebase=False)
        return r.fir, r.taps.cpu().numpy(), r

    pageable = None

    def step_e2e_pageable():
        # what a caller holding plain NumPy arrays (loadmat output) gets: the driver stages pageable memory itself
        r = P.identify_host([tuple(pageable[0])], tuple(pageable[1]), a.nd, KERNEL, NOISE, cands, n_taps=N_TAPS)
        return r.fir, r.taps.cpu().numpy(), r

    def step_e2e_calls():
        with contextlib.redirect_stdout(io.StringIO()):
            (t, u, y), (tv, uv, yv) = host
            U, Y = synchronous_coefficients_trapz_pair(t, u, y, omega_h)
            G = frf_cross_power_average([U], [Y])
            Uv, Yv = synchronous_coefficients_trapz_pair(tv, uv, yv, omega_v)
            Gv = frf_cross_power_average([Uv], [Yv])
            Xs = StandardScaler(); Xn = Xs.fit_transform(np.log10(omega_h).reshape(-1, 1))
            Xv = Xs.transform(np.log10(omega_v).reshape(-1, 1))
            gps = []
            for target, vy in ((G.real, Gv.real), (G.imag, Gv.imag)):
                sc = StandardScaler(); z = sc.fit_transform(target.reshape(-1, 1)).ravel()
                gp = GaussianProcessRegressor(kernel=create_kernel(KERNEL), noise_variance=NOISE)
                gp.fit(Xn, z, optimize=True, use_grid_search=True, max_grid_combinations=5000,
                       validation_X=Xv, validation_y=vy, y_scaler=sc)
                gps.append((gp, sc))

            def predict(w_new):
                Xq = Xs.transform(np.log10(w_new).reshape(-1, 1))
                parts = [sc.inverse_transform(gp.predict(Xq).reshape(-1, 1)).ravel() for gp, sc in gps]
                return parts[0] + 1j * parts[1]

            fir_res = gp_to_fir_direct_pipeline(omega_h, G, gp_predict_func=predict, mat_file=None, output_dir=e2e_dir)
            metrics, _ = bfir.fir_validate(uv, yv, fir_res["fir_coefficients"], want_pred=False)
        return metrics, fir_res["fir_coefficients"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_events=False):
        for _ in range(warmup):
            fn()
        barrier()
        _lib.launch_count = 0
        if with_events:
            _lib.event_log = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        log, _lib.event_log = _lib.event_log, None
        launches = _lib.launch_count
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / steps, out, log, launches

    # ---- device-resident timing (value) with clocks sampled during the timed region
    clocks = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        clocks.start()
    ms_step, res, _, launches = timed(step_resident, a.steps, a.warmup)
    clk = clocks.stop() if rank == 0 else {}
    # ---- per-stage kernel times from a second, event-instrumented run (same work, same stream)
    ms_inst, _, log, _ = timed(step_resident, a.steps, 1, with_events=True)
    stage_ms = {k: sum(s.elapsed_time(e) for s, e in v) / a.steps for k, v in (log or {}).items()}
    if a.resident_only:
        if rank == 0:
            print(json.dumps({"resident_only": True, "ms_per_step": ms_step, "steps": a.steps, "warmup": a.warmup,
                              "ms_per_step_by_call": stage_ms}), flush=True)
        return
    # ---- e2e timing: host pinned buffers in, result out
    from b200id.records import FACTS
    FACTS.clear()
    ms_e2e, (metrics_e, taps_e, res_e), _, _ = timed(step_e2e, a.steps, max(a.warmup, 3))
    # per step: train (u, y) + validation (u, y), each once -- the two time vectors were found to be exactly
    # fl(i dt) + t0 during the warm-up steps and are rebuilt on the device (b200id.records.FACTS); back: the final pack
    # (status 2, theta 6, selection 4 + 2, FIR sums 4, time-base verdict 1) and the taps
    exact_tb = all(bool(FACTS.get(torch.from_numpy(rec[0])).get("exact")) for rec in host)
    h2d = 8 * a.n * ((2 + 2) if exact_tb else (3 + 3))
    d2h = 8 * (2 + 6 + 4 + 2 + 4 + 1) + 8 * N_TAPS
    ms_e2e_full, (metrics_f, _, res_f), _, _ = timed(step_e2e_full_upload, a.steps, a.warmup)
    h2d_full = 8 * a.n * (3 + 3)
    # pageable host arrays: plain NumPy copies of the records (outside the timed region)
    pageable = [[np.array(x.numpy(), copy=True) for x in rec] for rec in pinned]
    # first call on fresh pageable arrays: the host layer page-locks them in place (cudaHostRegister) while it uploads
    # them for the first time; every later call on the same arrays runs at the pinned rate (b200id/runtime.py)
    torch.cuda.synchronize()
    t_first = time.perf_counter()
    step_e2e_pageable()
    torch.cuda.synchronize()
    ms_e2e_pg_first = (time.perf_counter() - t_first) * 1e3
    ms_e2e_pg, (metrics_p, _, _), _, _ = timed(step_e2e_pageable, max(2, a.steps // 2), max(a.warmup, 3))
    pageable = None
    # the same work as the reference's call sequence (drop-in surface, one synchronous call per stage, NumPy in
    # and out of every call): the validation record goes up twice (FRF, then FIR) and nothing overlaps across calls
    ms_calls, (metrics_c, taps_c), _, _ = timed(step_e2e_calls, a.steps, a.warmup)
    h2d_calls = 8 * a.n * (3 + 3 + 2) + 8 * (a.nd + VAL_ND) * 2 + 2 * 900 * 3 * 8 + 2 * 8 * (a.nd * 2 + VAL_ND * 2) + 8 * (N_TAPS + 1000 * 2)
    d2h_calls = 16 * 2 * (a.nd + VAL_ND) * 2 + 2 * (900 * 8 + 16) + 2 * 8 * (a.nd + a.nd * a.nd) + 2 * 1000 * 8 + N_TAPS * 8 + 4 * 8

    total_samples = a.n * world
    value = total_samples / (ms_step * 1e-3)
    e2e_value = total_samples / (ms_e2e * 1e-3)

    # ---- FP64 peak probe (roofline denominator for the FP64-bound kernels; not in MEASURED_PEAKS.json)
    fp64 = {}
    if rank == 0:
        sink = empty(1, dev)
        for mode, name in ((0, "dfma"), (1, "dmma"), (2, "dfma_plus_dmma")):
            flops = C.c_double(0.0)
            from b200id.runtime import ptr, stream_ptr
            iters = 4000
            _lib.call("b200id_fp64_peak_probe", C.c_int(mode), C.c_int(iters), ptr(sink), C.byref(flops), stream_ptr())
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.call("b200id_fp64_peak_probe", C.c_int(mode), C.c_int(iters), ptr(sink), C.byref(flops), stream_ptr())
                e.record()
                torch.cuda.synchronize()
                best = max(best, flops.value / (s.elapsed_time(e) * 1e-3) / 1e12)
            fp64[name + "_tflops"] = best

    configs = None
    if not a.no_configs:
        configs = run_configs(a, rank, world, dev, res, timed, hbm_peak)

    if rank != 0:
        return

    # ---- bin-adaptive FRF: time of the main-term phase alone (rounding term switched off), both launches of a step,
    # measured live with CUDA events on the records of this run; the rounding-term phase is the difference
    frf_phases = {}
    try:
        frf_phases = measure_frf_phases(dev_recs, a, dev)
    except Exception as exc:                                  # never lose the bench line over a diagnostic
        frf_phases = {"frf_phases_error": repr(exc)}

    # ---- roofline of the dominant kernel (largest share of the step among the instrumented calls)
    frf_ms = stage_ms.get("b200id_frf_project", 0.0) + stage_ms.get("b200id_frf_project_uniform", 0.0)
    frf_form = os.environ.get("B200ID_FRF_FORM", "0" if os.environ.get("B200ID_FRF_GOERTZEL", "1") == "0" else "2")
    frf_uniform_kernel = {"0": "frf_project_uniform3_kernel", "1": "frf_project_uniform_g_kernel"}.get(frf_form, "frf_main_ws_kernel")
    fir_kernel = "fir_fft_kernel" if os.environ.get("B200ID_FIR_RADIX4", "0") == "1" else "fir_fft16_kernel"
    frf_kernel = frf_uniform_kernel if stage_ms.get("b200id_frf_project_uniform", 0.0) > stage_ms.get("b200id_frf_project", 0.0) else "frf_project_kernel"
    fir_ms = stage_ms.get("b200id_fir_eval", 0.0)
    # Re and Im searches share one factorisation per candidate (b200id_gp_batch_eval_pair); the candidate count and
    # the algorithmic flops below stay those of the reference's 2 x 900 separate evaluations
    gp_ms = (stage_ms.get("b200id_gp_batch_eval", 0.0) + stage_ms.get("b200id_gp_batch_eval_pair", 0.0)
             + stage_ms.get("b200id_gp_batch_eval_shared", 0.0))      # product grids go through their shared covariance shapes
    dominant = max(((frf_kernel, frf_ms), (fir_kernel, fir_ms), ("gp_batch_tiled_kernel", gp_ms)), key=lambda kv: kv[1])[0]
    frf_bytes = FRF_BYTES_PER_SAMPLE * a.n * 2            # two launches per step: train (nd bins) + validation (150 bins)
    frf_flops = 8.0 * a.n * (a.nd + VAL_ND)               # 2 signals x (cos, sin) x FMA, both launches
    fir_bytes = FIR_BYTES_PER_SAMPLE * a.n
    fir_flops = 2.0 * N_TAPS * a.n
    gp_flops = 1800 / world * (a.nd ** 3 / 3.0 + 2.0 * a.nd ** 2)
    fp64_peak = max(fp64.values()) if fp64 else None
    if dominant == fir_kernel:
        ach = fir_bytes / (fir_ms * 1e-3) / 1e9
    else:
        ach = frf_bytes / (frf_ms * 1e-3) / 1e9 if frf_ms > 0 else 0.0
        dominant = frf_kernel
    roofline = {
        "kernel": dominant, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
        "traffic": NCU_TRAFFIC_BYTES.get(dominant), "peak_source": peak_src,
        "launches_per_step": 2 if dominant.startswith("frf") else 1,
        "note": "FP64 bound by construction (8 FP64 flops per sample-bin against 24 bytes per sample).  Bin-adaptive form: "
                "'achieved' is the whole b200id_frf_project_uniform call (least-squares time base on every 64th stamp 20 MB, "
                "main-term DMMA kernel 240 MB, rounding-term kernel for the listed bins 240 MB: 500 MB of DRAM traffic per "
                "launch in three passes, none of them bandwidth bound); the kernel named here is the largest of the three, its own "
                "time and FP64 fraction are in fp64.frf_main_*; see DESIGN.md K1",
        "call_traffic": 500.2e6 if dominant == "frf_main_ws_kernel" else None,
        "fp64": {
            "peak_tflops_measured": fp64_peak, **fp64,
            "frf_algorithmic_tflops": frf_flops / (frf_ms * 1e-3) / 1e12 if frf_ms > 0 else None,
            **frf_phases,
            # direct-form-equivalent rate: the overlap-save FFT path executes ~90 flops per output, not 2*taps
            "fir_direct_form_equivalent_tflops": fir_flops / (fir_ms * 1e-3) / 1e12 if fir_ms > 0 else None,
            "gp_algorithmic_tflops": gp_flops / (gp_ms * 1e-3) / 1e12 if gp_ms > 0 else None,
            "frf_frac": (frf_flops / (frf_ms * 1e-3) / 1e12 / fp64_peak) if (fp64_peak and frf_ms > 0) else None,
            "frf_main_frac": (frf_phases["frf_main_tflops"] / fp64_peak) if (fp64_peak and "frf_main_tflops" in frf_phases) else None,
            "fir_hbm_frac": (fir_bytes / (fir_ms * 1e-3) / 1e9 / hbm_peak) if fir_ms > 0 else None,
            "gp_frac": (gp_flops / (gp_ms * 1e-3) / 1e12 / fp64_peak) if (fp64_peak and gp_ms > 0) else None,
        },
    }
    if configs and fp64_peak:
        c3, c4 = configs["cfg3"], configs["cfg4"]
        c3["algorithmic_tflops"] = c3["algorithmic_flops"] / (c3["ms"] * 1e-3) / 1e12
        c3["fp64_frac"] = c3["algorithmic_tflops"] / (fp64_peak * world)
        c4["algorithmic_tflops_per_gpu"] = c4["algorithmic_flops"] / (c4["batch_ms"] * 1e-3) / 1e12
        c4["fp64_frac"] = c4["algorithmic_tflops_per_gpu"] / fp64_peak
        c4["single_cholesky_tflops"] = c4["single_cholesky_flops"] / (c4["single_cholesky_ms"] * 1e-3) / 1e12
        if "cfg5" in configs:
            c5 = configs["cfg5"]
            c5["frf_fp64_frac"] = c5["frf_algorithmic_flops"] / (c5["frf_ms"] * 1e-3) / 1e12 / (fp64_peak * world)
    stages = {
        "ms_per_step_by_call": stage_ms,
        "frf_samples_per_s_kernel": 2 * a.n / (frf_ms * 1e-3) if frf_ms > 0 else None,
        "fir_samples_per_s_kernel": a.n / (fir_ms * 1e-3) if fir_ms > 0 else None,
        "gpr_candidates_per_s_kernel": (1800 / world) / (gp_ms * 1e-3) if gp_ms > 0 else None,
        "fir_hbm_gbs": fir_bytes / (fir_ms * 1e-3) / 1e9 if fir_ms > 0 else None,
    }

    # ---- CPU baseline = the oracle port on the bench's OWN records (rank 0, N = 1 only), one pass; its results are
    # also what the parity block compares the GPU step with.  --cpu-sample m < n runs both sides on the first m samples.
    cpu, parity = None, None
    if world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = a.n if a.cpu_sample <= 0 else min(a.n, a.cpu_sample)
        tr_s, va_s = slice_records(train_h, val_h, n_s)
        st, ref = cpu_pipeline_pass(tr_s, va_s, a.nd, threads)
        cpu_s = scaled_step_seconds(st, n_s, a.n)
        cpu = {"value": a.n / cpu_s, "unit": UNIT, "cores": threads, "kind": "port", "n_samples_run": n_s,
               "stage_seconds_at_n_run": {k: round(v, 4) for k, v in st.items()},
               "sample": f"{n_s} of {a.n} samples per record of the same records, one pass"
                         + ("" if n_s == a.n else "; FRF and FIR seconds scaled by n/n_run, GP stage as measured")}
        if n_s == a.n:
            res_p = res
        else:
            tr = [tuple(x[:n_s] for x in dev_recs[0]) + (float(tr_s[0][-1] - tr_s[0][0]), float(tr_s[0][0]))]
            va = tuple(x[:n_s] for x in dev_recs[1]) + (float(va_s[0][-1] - va_s[0][0]), float(va_s[2][skip]), float(va_s[0][0]))
            res_p = P.identify_device(tr, va, a.nd, KERNEL, NOISE, cands, n_taps=N_TAPS)
        Gd = res_p.G.cpu().numpy()
        taps_d = res_p.taps.cpu().numpy()
        rel = lambda x, r: abs(x - r) / abs(r) if r != 0 else abs(x - r)
        parity = {
            "against": "oracle/ (NumPy restatement of the reference, pinned to reference-generated golden vectors)",
            "n_samples_compared": n_s,
            "frf_max_rel_err_per_bin": float(np.max(np.abs(Gd - ref["G"]) / np.abs(ref["G"]))),
            "frf_tolerance": 1e-9,
            "theta_identical": bool(np.array_equal(res_p.theta_real, ref["theta_real"]) and np.array_equal(res_p.theta_imag, ref["theta_imag"])),
            "theta_gpu": [res_p.theta_real.tolist(), res_p.theta_imag.tolist()],
            "theta_oracle": [ref["theta_real"].tolist(), ref["theta_imag"].tolist()],
            "selection_metric_rel_err": [rel(res_p.metric_real, ref["metric_real"]), rel(res_p.metric_imag, ref["metric_imag"])],
            "taps_max_rel_err": float(np.max(np.abs(taps_d - ref["taps"])) / np.max(np.abs(ref["taps"]))),
            "fir_rmse_rel_err": rel(res_p.fir["rmse"], ref["fir"]["rmse"]),
            "fir_fit_percent_abs_err": abs(res_p.fir["fit_percent"] - ref["fir"]["fit_percent"]),
            "e2e_theta_identical_to_resident": bool(np.array_equal(res_e.theta_real, res.theta_real) and np.array_equal(res_e.theta_imag, res.theta_imag)),
            "e2e_fir_rmse_rel_err_vs_resident": rel(metrics_e["rmse"], res.fir["rmse"]),
        }
        parity["pass"] = bool(parity["frf_max_rel_err_per_bin"] < 1e-9 and parity["theta_identical"]
                              and parity["taps_max_rel_err"] < 1e-6 and parity["fir_rmse_rel_err"] < 1e-6)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "b200id.pipeline.identify_host (host NumPy records in pinned memory -> taps, metrics, theta on the host)",
                "timebase": ("u and y cross the link every step; both time vectors are exactly fl(i dt) + t0 (checked on the device over "
                             "the whole record in the first warm-up step) and are rebuilt on the device" if exact_tb else
                             "t, u and y cross the link every step")},
        "e2e_full_upload": {"value": total_samples / (ms_e2e_full * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_full,
                            "h2d_bytes_per_step": h2d_full, "d2h_bytes_per_step": d2h,
                            "call": "identify_host(..., reuse_timebase=False): t, u, y uploaded every step"},
        "e2e_pageable": {"value": total_samples / (ms_e2e_pg * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_pg,
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "first_call_ms": ms_e2e_pg_first,
                         "call": "identify_host on plain (pageable) NumPy arrays, as scipy.io.loadmat returns them: the first call "
                                 "page-locks them in place while uploading (first_call_ms, host clock), later calls on the same "
                                 "arrays are the timed ones; B200ID_PIN_HOST=0 restores driver-staged pageable copies"},
        "e2e_call_by_call": {"value": total_samples / (ms_calls * 1e-3), "unit": UNIT, "ms_per_step": ms_calls,
                             "h2d_bytes_per_step": h2d_calls, "d2h_bytes_per_step": d2h_calls,
                             "call": "src.frequency_transform / src.gpr / src.fir_model drop-in functions, one synchronous call per stage"},
        "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "configs": configs, "stages": stages,
        "result": {"theta_real": res.theta_real.tolist(), "theta_imag": res.theta_imag.tolist(), "fir_rmse": res.fir.get("rmse"),
                   "e2e_fir_rmse": metrics_e.get("rmse"), "e2e_theta_real": res_e.theta_real.tolist(),
                   "e2e_theta_imag": res_e.theta_imag.tolist(), "e2e_call_by_call_fir_rmse": metrics_c.get("rmse")},
        "device": {"sm_count": sm, "cc": f"{major}.{minor}"}, "host_numa": numa,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
