#!/usr/bin/env python3
"""run_baseline-style loop (several kernels, same train / validation files) through the public file-level
surface, with the record store on and off: wall time per kernel, parses and uploads.

    python profiles/file_pipeline_reuse.py [--n 3000000] [--kernels matern52 rbf matern32 matern12]
"""
import argparse, contextlib, io, sys, tempfile, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200")); sys.path.insert(0, str(ROOT / "tests"))
from b200id import synth, records
from b200id.runtime import require_cuda
import pipeline_flow

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3_000_000)
ap.add_argument("--nd", type=int, default=50)
ap.add_argument("--kernels", nargs="+", default=["matern52", "rbf", "matern32", "matern12"])
a = ap.parse_args()
require_cuda()
with tempfile.TemporaryDirectory() as tmp:
    tmp = Path(tmp)
    t, u, y = synth.make_record(a.n, a.nd, "multisine", seed=0)
    train = synth.write_mat(tmp / "train.mat", t, u, y)
    tv, uv, yv = synth.make_record(a.n, a.nd, "square", seed=1)
    val = synth.write_mat(tmp / "val.mat", tv, uv, yv)
    results = {}
    for mode in ("warmup", "store_off", "store_on"):
        records.STORE.clear()
        records.ENABLED = mode != "store_off"
        per_kernel, out = [], []
        for k in (a.kernels[:1] if mode == "warmup" else a.kernels):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                r = pipeline_flow.run(train, val, tmp / f"out_{mode}_{k}", kernel=k, nd=a.nd, optimize=True, grid=True)
            torch.cuda.synchronize(); per_kernel.append(time.perf_counter() - t0)
            out.append((r["kernel_params"], r["fir_extraction"]["rmse"]))
        results[mode] = out
        if mode != "warmup":
            print(f"{mode}: per-kernel wall s = {[round(x, 3) for x in per_kernel]}  total {sum(per_kernel):.3f} s  stats {dict(records.STORE.stats)}")
    same = all(pa == pb and abs(ra - rb) <= 1e-12 * abs(rb) for (pa, ra), (pb, rb) in zip(results["store_off"], results["store_on"]))
    print("identical selections and FIR RMSE with the store on and off:", same)
