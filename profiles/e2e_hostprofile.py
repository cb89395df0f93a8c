#!/usr/bin/env python3
"""cProfile of the public-API (e2e) step of bench.py on the host side."""
import cProfile, pstats, io, contextlib, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "gauss-process_b200"))
import bench
from sklearn.preprocessing import StandardScaler
from src.frequency_transform.frf_estimator import matlab_freq_grid, synchronous_coefficients_trapz_pair, frf_cross_power_average
from src.gpr import create_kernel, GaussianProcessRegressor
from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline
from b200id import fir as bfir
from b200id.runtime import require_cuda
dev = require_cuda()
n, nd = 10_000_000, 100
train_h, val_h = bench.make_records(n, nd, seed=0, dev=dev)
pinned = [[torch.from_numpy(x).pin_memory() for x in rec] for rec in (train_h, val_h)]
host = [[x.numpy() for x in rec] for rec in pinned]
_, omega_h = matlab_freq_grid(nd, -1.0, 2.3); _, omega_v = matlab_freq_grid(150, -1.0, 2.3)
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        (t, u, y), (tv, uv, yv) = host
        U, Y = synchronous_coefficients_trapz_pair(t, u, y, omega_h)
        G = frf_cross_power_average([U], [Y])
        Uv, Yv = synchronous_coefficients_trapz_pair(tv, uv, yv, omega_v)
        Gv = frf_cross_power_average([Uv], [Yv])
        Xs = StandardScaler(); Xn = Xs.fit_transform(np.log10(omega_h).reshape(-1, 1)); Xv = Xs.transform(np.log10(omega_v).reshape(-1, 1))
        gps = []
        for target, vy in ((G.real, Gv.real), (G.imag, Gv.imag)):
            sc = StandardScaler(); z = sc.fit_transform(target.reshape(-1, 1)).ravel()
            gp = GaussianProcessRegressor(kernel=create_kernel("matern52"), noise_variance=1e-6)
            gp.fit(Xn, z, optimize=True, use_grid_search=True, max_grid_combinations=5000, validation_X=Xv, validation_y=vy, y_scaler=sc)
            gps.append((gp, sc))
        def predict(w_new):
            Xq = Xs.transform(np.log10(w_new).reshape(-1, 1))
            parts = [sc.inverse_transform(gp.predict(Xq).reshape(-1, 1)).ravel() for gp, sc in gps]
            return parts[0] + 1j * parts[1]
        fir_res = gp_to_fir_direct_pipeline(omega_h, G, gp_predict_func=predict, mat_file=None, output_dir=ROOT / "gpurun_out" / "e2e_prof")
        return bfir.fir_validate(uv, yv, fir_res["fir_coefficients"], want_pred=False)
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print("ms per step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])

# ---- wall time per public call (synchronised), to see where a step goes
def phases():
    out = {}
    def lap(name, t0):
        torch.cuda.synchronize(); out[name] = out.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    with contextlib.redirect_stdout(io.StringIO()):
        (t, u, y), (tv, uv, yv) = host
        t0 = time.perf_counter(); U, Y = synchronous_coefficients_trapz_pair(t, u, y, omega_h); lap("frf_train", t0)
        t0 = time.perf_counter(); Uv, Yv = synchronous_coefficients_trapz_pair(tv, uv, yv, omega_v); lap("frf_val", t0)
        t0 = time.perf_counter(); G = frf_cross_power_average([U], [Y]); Gv = frf_cross_power_average([Uv], [Yv]); lap("cross_power", t0)
        t0 = time.perf_counter()
        Xs = StandardScaler(); Xn = Xs.fit_transform(np.log10(omega_h).reshape(-1, 1)); Xv = Xs.transform(np.log10(omega_v).reshape(-1, 1))
        gps = []
        for target, vy in ((G.real, Gv.real), (G.imag, Gv.imag)):
            sc = StandardScaler(); z = sc.fit_transform(target.reshape(-1, 1)).ravel()
            gp = GaussianProcessRegressor(kernel=create_kernel("matern52"), noise_variance=1e-6)
            gp.fit(Xn, z, optimize=True, use_grid_search=True, max_grid_combinations=5000, validation_X=Xv, validation_y=vy, y_scaler=sc)
            gps.append((gp, sc))
        lap("gp_fit_x2", t0)
        def predict(w_new):
            Xq = Xs.transform(np.log10(w_new).reshape(-1, 1))
            parts = [sc.inverse_transform(gp.predict(Xq).reshape(-1, 1)).ravel() for gp, sc in gps]
            return parts[0] + 1j * parts[1]
        t0 = time.perf_counter()
        fir_res = gp_to_fir_direct_pipeline(omega_h, G, gp_predict_func=predict, mat_file=None, output_dir=ROOT / "gpurun_out" / "e2e_prof")
        lap("gp_to_fir", t0)
        t0 = time.perf_counter(); bfir.fir_validate(uv, yv, fir_res["fir_coefficients"], want_pred=False); lap("fir_validate", t0)
    return out
acc = {}
for _ in range(5):
    for k, v in phases().items():
        acc[k] = acc.get(k, 0.0) + v / 5
print("phase ms:", {k: round(v, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()), 3))
