#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE itself (build container only).

    python oracle/make_golden.py            # writes tests/golden/

Imports /root/reference through oracle/refharness.py (matplotlib stubbed) and records,
for seeded synthetic inputs, the outputs of every reference function on the hot path
(SURVEY.md section 8a).  The vectors pin the NumPy oracle (tests/test_oracle_golden.py) and
are a second, oracle-independent anchor for the CUDA path (tests/test_*_gpu.py).  Installed
versions are stored in each file (``versions``).
"""
from __future__ import annotations

import argparse
import io
import math
import contextlib
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "gauss-process_b200" / "b200id"))
import refharness  # noqa: E402
import synth       # noqa: E402  (host-only synthetic generator, no `src` import)

refharness.load()
import warnings  # noqa: E402
warnings.filterwarnings("ignore")
np.seterr(all="ignore")
import scipy, sklearn, pandas  # noqa: E402,E401
from src.frequency_transform import frf_estimator as R_frf  # noqa: E402
from src.frequency_transform.transform import estimate_frequency_response as R_estimate  # noqa: E402
from src.gpr.kernels import create_kernel as R_create_kernel  # noqa: E402
from src.gpr.gpr_fitting import GaussianProcessRegressor as R_GPR  # noqa: E402
from src.gpr.grid_search import get_default_param_grids as R_grids  # noqa: E402
from src.fir_model import fir_helpers as R_firh  # noqa: E402
from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline as R_fir_pipeline  # noqa: E402
from src.fir_model.fir_validation import validate_fir_with_mat as R_validate  # noqa: E402
from sklearn.preprocessing import StandardScaler  # noqa: E402

VERSIONS = np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}",
                     f"sklearn {sklearn.__version__}", f"pandas {pandas.__version__}"])
KERNEL_NAMES = ["rbf", "matern12", "matern32", "matern52", "exp", "dc", "di", "ss1", "ss2", "sshf",
                "stable_spline", "rq", "tc"]
BASELINE = ["dc", "di", "exp", "matern12", "matern32", "matern52", "rbf", "ss1", "ss2", "sshf", "stable_spline"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def set_theta(kernel, theta):
    kernel.set_params(np.asarray(theta, dtype=float))
    return kernel


# ------------------------------------------------------------------ FRF
def golden_frf(out: Path):
    d = {"versions": VERSIONS}
    n, nd = 6000, 50
    t, u, y = synth.make_record(n, nd, "multisine", seed=0)
    _, w = R_frf.matlab_freq_grid(nd, -1.0, 2.3)
    d.update(t=t, u=u, y=y, w=w)
    d["grid_hz_37"], d["grid_w_37"] = R_frf.matlab_freq_grid(37, -1.0, 2.3)
    d["U"] = R_frf.synchronous_coefficients_trapz(t, u, w, 0.0, True)
    d["Y"] = R_frf.synchronous_coefficients_trapz(t, y, w, 0.0, True)
    d["G"] = R_frf.frf_cross_power_average([d["U"]], [d["Y"]])
    d["U_drop"] = R_frf.synchronous_coefficients_trapz(t, u, w, 1.5, True)
    d["Y_nomean"] = R_frf.synchronous_coefficients_trapz(t, y + 0.3, w, 0.0, False)
    # non-uniform timebase (jittered sampling instants)
    rng = np.random.default_rng(5)
    tj = np.sort(t + rng.uniform(-0.4, 0.4, n) * synth.DT)
    d["t_jit"] = tj
    d["U_jit"] = R_frf.synchronous_coefficients_trapz(tj, u, w, 0.0, True)
    d["Y_jit"] = R_frf.synchronous_coefficients_trapz(tj, y, w, 0.0, True)
    # two records -> cross-power average; second record has a dead input bin set
    t2, u2, y2 = synth.make_record(n, nd, "multisine", seed=3)
    U2 = R_frf.synchronous_coefficients_trapz(t2, u2, w, 0.0, True)
    Y2 = R_frf.synchronous_coefficients_trapz(t2, y2, w, 0.0, True)
    d.update(u2=u2, y2=y2, U2=U2, Y2=Y2)
    d["G2"] = R_frf.frf_cross_power_average([d["U"], U2], [d["Y"], Y2])
    Uz = d["U"].copy(); Uz[[3, 17]] = 0.0
    d["G_nan"] = R_frf.frf_cross_power_average([Uz], [d["Y"]])
    # file-level entry: estimate_frequency_response on .mat files (time_duration window, 2 files)
    with tempfile.TemporaryDirectory() as tmp:
        p1 = synth.write_mat(Path(tmp) / "a.mat", t, u, y)
        p2 = synth.write_mat(Path(tmp) / "b.mat", t2, u2, y2)
        df = quiet(R_estimate, [str(p1), str(p2)], method="frf", nd=nd, n_files=2)
        d["df_2files"] = df.to_numpy()
        df = quiet(R_estimate, [str(p1)], method="frf", nd=30, time_duration=7.0, drop_seconds=1.0, n_files=1)
        d["df_window"] = df.to_numpy()
    np.savez_compressed(out / "frf.npz", **d)
    print("frf.npz", {k: np.shape(v) for k, v in d.items()})


# ------------------------------------------------------------------ GP
THETAS = {
    "rbf": [[0.7, 1.3], [0.05, 10.0], [30.0, 0.02]],
    "matern12": [[0.7, 1.3], [0.05, 10.0]],
    "matern32": [[0.7, 1.3], [3.0, 0.2]],
    "matern52": [[0.7, 1.3], [0.2, 5.0]],
    "exp": [[0.8, 1.5], [5.0, 0.1]],
    "dc": [[0.9, 1.0, 0.5], [0.3, 2.0, -0.7]],
    "di": [[1.0, 0.9], [4.0, 0.2]],
    "ss1": [[1.0, 1.0], [0.02, 30.0]],
    "ss2": [[1.0, 1.0], [0.3, 3.0]],
    "sshf": [[1.0, 1.0], [2.5, 0.4]],
    "stable_spline": [[0.5, 1.0], [2.0, 7.0]],
    "rq": [[0.7, 1.3, 1.0], [2.0, 0.5, 0.05]],
    "tc": [[0.8, 1.5], [5.0, 0.1]],
}


def metric_ref(kernel, theta, X, y, noise, Xv=None, yv=None, y_scaler=None):
    """The reference's per-candidate closure (grid_search.py:91-133) driven with reference kernels."""
    kernel.set_params(np.asarray(theta, dtype=float))
    K = kernel(X)
    K += noise * np.eye(K.shape[0])
    try:
        L = np.linalg.cholesky(K)
        a = np.linalg.solve(L, y)
        a = np.linalg.solve(L.T, a)
        if Xv is not None:
            yp = kernel(Xv, X) @ a
            if y_scaler is not None:
                yp = y_scaler.inverse_transform(yp.reshape(-1, 1)).ravel()
            return np.sqrt(np.mean((yv - yp) ** 2))
        v = 0.5 * (y @ a)
        v += np.sum(np.log(np.diag(L)))
        v += 0.5 * len(y) * np.log(2 * np.pi)
        return v
    except np.linalg.LinAlgError:
        return 1e10


def gp_problem(nd, n_total=40000, nv=150):
    """Standardised GP regression problem as gp_pipeline.py:125-156 builds it."""
    t, u, y = synth.make_record(n_total, nd, "multisine", seed=0)
    _, w = R_frf.matlab_freq_grid(nd, -1.0, 2.3)
    G = R_frf.frf_cross_power_average([R_frf.synchronous_coefficients_trapz(t, u, w)],
                                      [R_frf.synchronous_coefficients_trapz(t, y, w)])
    tv, uv, yv = synth.make_record(n_total, nd, "square", seed=1)
    _, wv = R_frf.matlab_freq_grid(nv, -1.0, 2.3)
    Gv = R_frf.frf_cross_power_average([R_frf.synchronous_coefficients_trapz(tv, uv, wv)],
                                       [R_frf.synchronous_coefficients_trapz(tv, yv, wv)])
    Xs = StandardScaler(); X = Xs.fit_transform(np.log10(w).reshape(-1, 1))
    ys_r = StandardScaler(); yr = ys_r.fit_transform(G.real.reshape(-1, 1)).ravel()
    ys_i = StandardScaler(); yi = ys_i.fit_transform(G.imag.reshape(-1, 1)).ravel()
    Xv = Xs.transform(np.log10(wv).reshape(-1, 1))
    return dict(w=w, G=G, wv=wv, Gv=Gv, X=X, yr=yr, yi=yi, Xv=Xv, ys_r=ys_r, ys_i=ys_i, Xs=Xs)


def golden_gp(out: Path):
    d = {"versions": VERSIONS}
    P = gp_problem(50)
    X, yr, Xv = P["X"], P["yr"], P["Xv"]
    d.update(X=X, yr=yr, yi=P["yi"], Xv=Xv, yv_r=P["Gv"].real, yv_i=P["Gv"].imag,
             w=P["w"], G=P["G"],
             yr_mean=P["ys_r"].mean_, yr_scale=P["ys_r"].scale_,
             yi_mean=P["ys_i"].mean_, yi_scale=P["ys_i"].scale_,
             X_mean=P["Xs"].mean_, X_scale=P["Xs"].scale_)
    # integer-ish inputs so that DC/DI/SSHF rounding and Exp/TC Heaviside have variety
    Xi = np.linspace(-2.6, 6.4, 19).reshape(-1, 1)
    Xj = np.linspace(-1.2, 5.1, 7).reshape(-1, 1)
    d.update(Xi=Xi, Xj=Xj)
    for name in KERNEL_NAMES:
        for q, th in enumerate(THETAS[name]):
            k = set_theta(R_create_kernel(name), th)
            d[f"gram_{name}_{q}"] = k(X[:24])
            d[f"gramx_{name}_{q}"] = k(Xv[:10], X[:24])
            d[f"grami_{name}_{q}"] = k(Xi)
            d[f"gramij_{name}_{q}"] = k(Xj, Xi)
        d[f"theta_{name}"] = np.array(THETAS[name])
        d[f"bounds_{name}"] = np.array(R_create_kernel(name).bounds)
        d[f"params0_{name}"] = R_create_kernel(name).get_params()
        g = R_grids(R_create_kernel(name))
        d[f"grid_{name}"] = np.stack([g[f"param_{i}"] for i in range(len(g))])

    # per-candidate metrics over the reference's own candidate list (scan order), noise 1e-6
    import itertools, random
    for name in BASELINE:
        k = R_create_kernel(name)
        g = R_grids(k)
        combos = list(itertools.product(*[g[f"param_{i}"] for i in range(len(g))]))
        if len(combos) > 1500:
            random.seed(42)
            combos = random.sample(combos, 1500)
        C = np.array(combos)
        d[f"cand_{name}"] = C
        d[f"nll_{name}"] = np.array([metric_ref(k, th, X, yr, 1e-6) for th in C])
        d[f"vrmse_{name}"] = np.array([metric_ref(k, th, X, yr, 1e-6, Xv, P["Gv"].real, P["ys_r"]) for th in C])
        d[f"vrmse_noscaler_{name}"] = np.array([metric_ref(k, th, X, yr, 1e-6, Xv, P["Gv"].real, None) for th in C[:200]])
        d[f"nll0_{name}"] = np.array([metric_ref(k, th, X, yr, 0.0) for th in C])

    # end-to-end selection through GaussianProcessRegressor.fit(use_grid_search=True)
    for name in BASELINE:
        for mode in ("nll", "val"):
            for noise, tag in ((1e-6, "n6"), (0.0, "n0")):
                gp = R_GPR(kernel=R_create_kernel(name), noise_variance=noise)
                kw = dict(validation_X=Xv, validation_y=P["Gv"].real, y_scaler=P["ys_r"]) if mode == "val" else {}
                mx = 1500 if name == "dc" else 5000
                try:
                    quiet(gp.fit, X, yr, optimize=True, use_grid_search=True, max_grid_combinations=mx, **kw)
                    sel = gp.kernel.get_params()
                    d[f"sel_{name}_{mode}_{tag}"] = sel
                    d[f"sel_alpha_{name}_{mode}_{tag}"] = gp.alpha
                    m, s = gp.predict(X, return_std=True)
                    d[f"sel_mean_{name}_{mode}_{tag}"] = m
                    d[f"sel_std_{name}_{mode}_{tag}"] = s
                except Exception as e:   # all-NaN grid -> set_params(None) raises (grid_search.py:203)
                    d[f"sel_{name}_{mode}_{tag}"] = np.array([np.nan])
                    print("  selection raised for", name, mode, tag, type(e).__name__, e)

    # plain fit / predict (no optimisation), noise 1e-6: alpha, K_inv, mean, std at train and new points
    Xs_new = np.linspace(X.min() - 0.2, X.max() + 0.2, 64).reshape(-1, 1)
    d["Xs_new"] = Xs_new
    for name in KERNEL_NAMES:
        th = THETAS[name][0]
        gp = R_GPR(kernel=set_theta(R_create_kernel(name), th), noise_variance=1e-6)
        gp.fit(X, yr, optimize=False)
        d[f"fit_alpha_{name}"] = gp.alpha
        d[f"fit_Kinv_{name}"] = gp.K_inv
        m, s = gp.predict(X, return_std=True)
        d[f"fit_mean_train_{name}"], d[f"fit_std_train_{name}"] = m, s
        m, s = gp.predict(Xs_new, return_std=True)
        d[f"fit_mean_new_{name}"], d[f"fit_std_new_{name}"] = m, s
    # pinv fallback (noise 0, rank-deficient Gram): DC on standardised inputs
    gp = R_GPR(kernel=set_theta(R_create_kernel("dc"), [0.9, 1.0, 0.5]), noise_variance=0.0)
    gp.fit(X, yr, optimize=False)
    d["pinv_alpha_dc"], d["pinv_Kinv_dc"] = gp.alpha, gp.K_inv
    d["pinv_mean_dc"] = gp.predict(Xs_new)

    # L-BFGS-B path (gpr_fitting.py:102-155), seeded like run_baseline.py:129
    for name in ("matern52", "rbf"):
        np.random.seed(42)
        gp = R_GPR(kernel=R_create_kernel(name), noise_variance=1e-6)
        gp.fit(X, yr, optimize=True, n_restarts=3)
        d[f"lbfgs_theta_{name}"] = gp.kernel.get_params()
        d[f"lbfgs_noise_{name}"] = np.array(gp.noise_variance)
        d[f"lbfgs_nll_{name}"] = np.array(metric_ref(gp.kernel, gp.kernel.get_params(), X, yr, gp.noise_variance))
    np.savez_compressed(out / "gp.npz", **d)
    print("gp.npz", len(d), "arrays")


# ------------------------------------------------------------------ FIR
def golden_fir(out: Path):
    d = {"versions": VERSIONS}
    nd = 50
    _, w = R_frf.matlab_freq_grid(nd, -1.0, 2.3)
    G = synth.plant_response(w)
    d.update(w=w, G=G)
    # helpers on a small spectrum and on the paper-size one
    g_small = G[:9]
    Xh = R_firh.build_two_sided_hermitian_from_positive(g_small)
    d["herm_small"] = Xh
    d["taps_small"], d["hfull_small"] = R_firh.idft_to_fir_coeffs_two_sided(Xh, 11)
    w_uni = np.linspace(w.min(), w.max(), 1000)
    G_uni = synth.plant_response(w_uni)
    Xp = R_firh.build_two_sided_hermitian_from_positive(G_uni)
    d["G_uni"] = G_uni
    d["taps_paper"], _ = R_firh.idft_to_fir_coeffs_two_sided(Xp, 1024)
    # full paper-mode pipeline with and without predictor, validated on a square-wave record
    n = 8000
    t, u, y = synth.make_record(n, nd, "square", seed=1)
    d.update(t=t, u=u, y=y)
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            p = synth.write_mat(Path(tmp) / "val.mat", t, u, y)
            r = quiet(R_fir_pipeline, w, G, gp_predict_func=synth.plant_response, mat_file=p,
                      output_dir=Path(tmp) / "o1")
            d["pipe_taps"] = r["fir_coefficients"]
            d["pipe_metrics"] = np.array([r["rmse"], r["fit_percent"], r["r2"], r["transient_skip"], r["Ts"]])
            r2 = quiet(R_fir_pipeline, w, G, gp_predict_func=None, mat_file=p, output_dir=Path(tmp) / "o2",
                       fir_order=300)
            d["pipe_interp_taps"] = r2["fir_coefficients"]
            d["pipe_interp_metrics"] = np.array([r2["rmse"], r2["fit_percent"], r2["r2"], r2["transient_skip"]])
            v = quiet(R_validate, r["fir_coefficients"], p, output_dir=None, detrend=False)
            d["validate_metrics"] = np.array([v["rmse"], v["fit_percent"], v["r2"], v["transient_skip"]])
            vd = quiet(R_validate, r["fir_coefficients"][:64], p, output_dir=None, detrend=True)
            d["validate_detrend_metrics"] = np.array([vd["rmse"], vd["fit_percent"], vd["r2"], vd["transient_skip"]])
        finally:
            os.chdir(cwd)
    d["y_pred"] = np.convolve(u, d["pipe_taps"], mode="full")[:n]
    np.savez_compressed(out / "fir.npz", **d)
    print("fir.npz", {k: np.shape(v) for k, v in d.items()})


# ------------------------------------------------------------------ whole pipeline
def golden_pipeline(out: Path):
    """run_gp_pipeline (gp_pipeline.py:269) on synthetic .mat files, README command line shape."""
    import argparse as ap
    from src.pipeline.gp_pipeline import run_gp_pipeline
    d = {"versions": VERSIONS}
    n, nd = 30000, 50
    d["n"] = np.array(n); d["nd"] = np.array(nd)
    tt, ut, yt = synth.make_record(n, nd, "multisine", seed=0)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            ptrain = synth.write_mat(Path(tmp) / "train.mat", tt, ut, yt)
            pval = synth.write_mat(Path(tmp) / "val.mat", tv, uv, yv)
            for tag, kernel, opt, grid, noise in (("m52_plain", "matern52", False, False, 1e-6),
                                                  ("m52_grid", "matern52", True, True, 1e-6),
                                                  ("rbf_grid_n0", "rbf", True, True, 0.0),
                                                  ("ss2_grid_n0", "ss2", True, True, 0.0)):
                np.random.seed(42)
                cfg = ap.Namespace(
                    mat_files=[str(ptrain)], use_existing=None, n_files=1, time_duration=None, nd=nd,
                    freq_method="frf", kernel=kernel, nu=None, gp_mode="separate", noise_variance=noise,
                    normalize=True, log_frequency=True, optimize=opt, n_restarts=3,
                    out_dir=str(Path(tmp) / tag), extract_fir=True, fir_length=1024,
                    fir_validation_mat=str(pval), method="gp", is_gp=True, use_grid_search=grid,
                    grid_search_max_combinations=5000, validation_mat=None, n_numerator=2, n_denominator=4)
                res = quiet(run_gp_pipeline, cfg)
                frf = pandas.read_csv(Path(tmp) / tag / "unified_frf.csv").to_numpy()
                d[f"{tag}_frf"] = frf
                d[f"{tag}_theta_real"] = np.array(res["kernel_params"]["real"])
                d[f"{tag}_theta_imag"] = np.array(res["kernel_params"]["imag"])
                d[f"{tag}_gp_rmse"] = np.array([res["real"]["rmse"], res["imag"]["rmse"],
                                                res["real"]["r2"], res["imag"]["r2"]])
                f = res["fir_extraction"]
                d[f"{tag}_taps"] = f["fir_coefficients"]
                d[f"{tag}_fir"] = np.array([f["rmse"], f["fit_percent"], f["r2"]])
                print(" ", tag, "theta", res["kernel_params"]["real"], "fir rmse", f["rmse"])
        finally:
            os.chdir(cwd)
    np.savez_compressed(out / "pipeline.npz", **d)
    print("pipeline.npz", len(d), "arrays")


# ------------------------------------------------------------------ Fourier estimator
def golden_fourier(out: Path):
    """method='fourier' (fourier_estimator.py + transform.py:160-181): per-stage vectors on an even- and an
    odd-length record, and the file-level entry with two files, a window, a drop and a duration."""
    from src.frequency_transform import fourier_estimator as R_fe
    d = {"versions": VERSIONS}
    for tag, n in (("even", 6000), ("odd", 4999)):
        t, u, y = synth.make_record(n, 50, "multisine", seed=11)
        # (records are regenerated in the tests from the same seeds: synth.make_record(n, 50, "multisine", seed=11))
        for wname in (("hann", "none", "blackman") if tag == "even" else ("hann",)):
            f, G = R_fe.compute_transfer_function(t, u, y, window=None if wname == "none" else wname)
            d[f"G_{tag}_{wname}"] = G
        d[f"f_{tag}"] = f
        fu, U = R_fe.compute_fourier_transform(t, u, None, "hann")
        d[f"U_{tag}_hann"] = U
    t, u, y = synth.make_record(6000, 50, "multisine", seed=0)
    t2, u2, y2 = synth.make_record(5555, 50, "multisine", seed=3)
    with tempfile.TemporaryDirectory() as tmp:
        p1 = synth.write_mat(Path(tmp) / "a.mat", t, u, y)
        p2 = synth.write_mat(Path(tmp) / "b.mat", t2, u2, y2)
        d["df_2files"] = quiet(R_estimate, [str(p1), str(p2)], method="fourier", nd=40, f_low=0.1, f_up=100.0, n_files=2).to_numpy()
        d["df_window"] = quiet(R_estimate, [str(p1)], method="fourier", nd=25, f_low=0.5, f_up=60.0, time_duration=7.0,
                               drop_seconds=1.0, n_files=1, window="none").to_numpy()
        d["df_top"] = quiet(R_estimate, [str(p1)], method="fourier", nd=30, f_low=0.0, f_up=1e9, n_files=1).to_numpy()
        d["df_blackman"] = quiet(R_estimate, [str(p2)], method="fourier", nd=17, f_low=1.0, f_up=33.0, n_files=1,
                                 window="blackman").to_numpy()
    np.savez_compressed(out / "fourier.npz", **d)
    print("fourier.npz", {k: np.shape(v) for k, v in d.items()})


def golden_pipeline_polar(out: Path):
    """run_gp_pipeline with gp_mode='polar' (gp_pipeline.py:159-193: GPs on log-magnitude and unwrapped phase,
    grid search without validation data, i.e. the NLL metric) on the same synthetic .mat files."""
    import argparse as ap
    from src.pipeline.gp_pipeline import run_gp_pipeline
    d = {"versions": VERSIONS}
    n, nd = 30000, 50
    d["n"] = np.array(n); d["nd"] = np.array(nd)
    tt, ut, yt = synth.make_record(n, nd, "multisine", seed=0)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            ptrain = synth.write_mat(Path(tmp) / "train.mat", tt, ut, yt)
            pval = synth.write_mat(Path(tmp) / "val.mat", tv, uv, yv)
            for tag, kernel, opt, grid, noise in (("m52_polar_plain", "matern52", False, False, 1e-6),
                                                  ("m52_polar_grid", "matern52", True, True, 1e-6),
                                                  ("rbf_polar_grid", "rbf", True, True, 1e-6)):
                np.random.seed(42)
                cfg = ap.Namespace(
                    mat_files=[str(ptrain)], use_existing=None, n_files=1, time_duration=None, nd=nd,
                    freq_method="frf", kernel=kernel, nu=None, gp_mode="polar", noise_variance=noise,
                    normalize=True, log_frequency=True, optimize=opt, n_restarts=3,
                    out_dir=str(Path(tmp) / tag), extract_fir=True, fir_length=1024,
                    fir_validation_mat=str(pval), method="gp", is_gp=True, use_grid_search=grid,
                    grid_search_max_combinations=5000, validation_mat=None, n_numerator=2, n_denominator=4)
                res = quiet(run_gp_pipeline, cfg)
                d[f"{tag}_theta_mag"] = np.array(res["kernel_params"]["magnitude"])
                d[f"{tag}_theta_phase"] = np.array(res["kernel_params"]["phase"])
                d[f"{tag}_gp_rmse"] = np.array([res["magnitude"]["rmse"], res["phase"]["rmse"],
                                                res["magnitude"]["r2"], res["phase"]["r2"]])
                f = res["fir_extraction"]
                d[f"{tag}_taps"] = f["fir_coefficients"]
                d[f"{tag}_fir"] = np.array([f["rmse"], f["fit_percent"], f["r2"]])
                print(" ", tag, "theta", res["kernel_params"]["magnitude"], res["kernel_params"]["phase"],
                      "fir rmse", f["rmse"])
        finally:
            os.chdir(cwd)
    np.savez_compressed(out / "pipeline_polar.npz", **d)
    print("pipeline_polar.npz", len(d), "arrays")


def golden_matern_nu(out: Path):
    """General-nu Matern (kernels.py:131-135, the scipy.special.kv branch of MaternKernel): Gram matrices for several
    nu / length scales on the GP problem's inputs and on widely spaced inputs (kv arguments from 1e-5 to beyond
    the underflow of exp(-x)), per-candidate NLL and validation RMSE over the reference's own 900-point scan list,
    the grid-search selection, fit / predict, and one whole run_gp_pipeline (--kernel matern --nu 0.8)."""
    import argparse as ap
    import itertools
    from src.gpr.kernels import MaternKernel
    from src.pipeline.gp_pipeline import run_gp_pipeline
    d = {"versions": VERSIONS}
    P = gp_problem(50)
    X, yr, Xv = P["X"], P["yr"], P["Xv"]
    d.update(X=X, yr=yr, Xv=Xv, yv_r=P["Gv"].real, yr_mean=P["ys_r"].mean_, yr_scale=P["ys_r"].scale_)
    Xw = np.concatenate([[0.0], np.logspace(-6, 3.2, 40)]).reshape(-1, 1)          # distances from 1e-6 to 1.6e3
    d["Xw"] = Xw
    nus = [0.3, 0.8, 1.0, 2.0, 3.3, 7.5, 0.50001, 12.25]
    d["nus"] = np.array(nus)
    thetas = [(1.0, 1.0), (0.05, 2.5), (30.0, 0.2), (1e-3, 1e3), (1e3, 1e-3)]
    d["thetas"] = np.array(thetas)
    for a, nu in enumerate(nus):
        for q, (ls, var) in enumerate(thetas):
            k = MaternKernel(length_scale=ls, variance=var, nu=nu)
            with np.errstate(all="ignore"):
                d[f"gram_{a}_{q}"] = k(X[:24])
                d[f"gramx_{a}_{q}"] = k(Xv[:10], X[:24])
                d[f"gramw_{a}_{q}"] = k(Xw)
    for a, nu in ((1, 0.8), (4, 3.3)):
        k = MaternKernel(nu=nu)
        g = R_grids(k)
        C = np.array(list(itertools.product(*[g[f"param_{i}"] for i in range(len(g))])))
        d[f"cand_{a}"] = C
        with np.errstate(all="ignore"):
            d[f"nll_{a}"] = np.array([metric_ref(k, th, X, yr, 1e-6) for th in C])
            d[f"vrmse_{a}"] = np.array([metric_ref(k, th, X, yr, 1e-6, Xv, P["Gv"].real, P["ys_r"]) for th in C])
        for mode in ("nll", "val"):
            gp = R_GPR(kernel=MaternKernel(nu=nu), noise_variance=1e-6)
            kw = dict(validation_X=Xv, validation_y=P["Gv"].real, y_scaler=P["ys_r"]) if mode == "val" else {}
            quiet(gp.fit, X, yr, optimize=True, use_grid_search=True, **kw)
            d[f"sel_{a}_{mode}"] = gp.kernel.get_params()
            d[f"sel_alpha_{a}_{mode}"] = gp.alpha
            m, s_ = gp.predict(X, return_std=True)
            d[f"sel_mean_{a}_{mode}"] = m
            d[f"sel_std_{a}_{mode}"] = s_
        gp = R_GPR(kernel=MaternKernel(length_scale=0.7, variance=1.3, nu=nu), noise_variance=1e-6)
        gp.fit(X, yr, optimize=False)
        d[f"fit_alpha_{a}"] = gp.alpha
        Xs_new = np.linspace(X.min() - 0.2, X.max() + 0.2, 64).reshape(-1, 1)
        d["Xs_new"] = Xs_new
        m, s_ = gp.predict(Xs_new, return_std=True)
        d[f"fit_mean_{a}"] = m
        d[f"fit_std_{a}"] = s_
    n, nd = 30000, 50
    d["n"] = np.array(n); d["nd"] = np.array(nd)
    tt, ut, yt = synth.make_record(n, nd, "multisine", seed=0)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            ptrain = synth.write_mat(Path(tmp) / "train.mat", tt, ut, yt)
            pval = synth.write_mat(Path(tmp) / "val.mat", tv, uv, yv)
            np.random.seed(42)
            cfg = ap.Namespace(
                mat_files=[str(ptrain)], use_existing=None, n_files=1, time_duration=None, nd=nd,
                freq_method="frf", kernel="matern", nu=0.8, gp_mode="separate", noise_variance=1e-6,
                normalize=True, log_frequency=True, optimize=True, n_restarts=3,
                out_dir=str(Path(tmp) / "m08"), extract_fir=True, fir_length=1024,
                fir_validation_mat=str(pval), method="gp", is_gp=True, use_grid_search=True,
                grid_search_max_combinations=5000, validation_mat=None, n_numerator=2, n_denominator=4)
            res = quiet(run_gp_pipeline, cfg)
            d["pipe_theta_real"] = np.array(res["kernel_params"]["real"])
            d["pipe_theta_imag"] = np.array(res["kernel_params"]["imag"])
            f = res["fir_extraction"]
            d["pipe_taps"] = f["fir_coefficients"]
            d["pipe_fir"] = np.array([f["rmse"], f["fit_percent"], f["r2"]])
            print("  pipeline matern nu=0.8 theta", res["kernel_params"], "fir rmse", f["rmse"])
        finally:
            os.chdir(cwd)
    np.savez_compressed(out / "matern_nu.npz", **d)
    print("matern_nu.npz", len(d), "arrays")


def golden_run_baseline(out: Path):
    """run_baseline.py's table: all 11 GP kernels through run_gp_pipeline with build_config's settings
    (run_baseline.py:51-86: N_d = 50, noise 0, grid search on the validation RMSE, max 500000 combinations,
    np.random.seed(42) per method, :129) on the synthetic train / validation .mat files."""
    import argparse as ap
    from src.pipeline.gp_pipeline import run_gp_pipeline
    from src.run_baseline import build_config, GP_METHODS
    d = {"versions": VERSIONS}
    n, nd = 30000, 50
    d["n"] = np.array(n); d["nd"] = np.array(nd)
    tt, ut, yt = synth.make_record(n, nd, "multisine", seed=0)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd(); os.chdir(tmp)
        try:
            ptrain = synth.write_mat(Path(tmp) / "train.mat", tt, ut, yt)
            pval = synth.write_mat(Path(tmp) / "val.mat", tv, uv, yv)
            for kernel in GP_METHODS:
                np.random.seed(42)
                cfg = build_config(kernel, [str(ptrain)], str(pval), str(Path(tmp) / kernel))
                res = quiet(run_gp_pipeline, cfg)
                d[f"{kernel}_theta_real"] = np.array(res["kernel_params"]["real"])
                d[f"{kernel}_theta_imag"] = np.array(res["kernel_params"]["imag"])
                d[f"{kernel}_gp_rmse"] = np.array([res["real"]["rmse"], res["imag"]["rmse"], res["real"]["r2"], res["imag"]["r2"]])
                f = res["fir_extraction"]
                d[f"{kernel}_taps"] = f["fir_coefficients"]
                d[f"{kernel}_fir"] = np.array([f["rmse"], f["fit_percent"], f["r2"]])
                print(" ", kernel, "theta", res["kernel_params"]["real"], res["kernel_params"]["imag"], "fir rmse", f["rmse"])
        finally:
            os.chdir(cwd)
    np.savez_compressed(out / "run_baseline.npz", **d)
    print("run_baseline.npz", len(d), "arrays")


def golden_kr(out: Path):
    """Kernel-regularised FIR (src/fir_model/kernel_regularized.py): regressor summaries, the three covariances behind
    the unconstrained parameters, the marginal likelihood at scattered hyper-parameter points (including ones that
    take the 1e-6 fallback jitter), the posterior mean, and whole run_fir_experiment runs (L-BFGS-B multi-start) on the
    reference's own synthetic fallback data and on a .mat record."""
    from src.fir_model import kernel_regularized as R_kr
    d = {"versions": VERSIONS}
    rng = np.random.default_rng(3)
    t, u, y = synth.make_record(6000, 50, "multisine", seed=4)
    u, y = u - u.mean(), y - y.mean()
    d["u"], d["y"] = u, y
    theta0 = {"dc": [math.log(1.0), math.log(0.98 / 0.02), np.arctanh(0.8)], "ss": [math.log(1.0), math.log(0.96 / 0.04)],
              "si": [math.log(0.98 / 0.02), math.log(0.25), math.log(1e-2), math.log(0.96 / 0.04)]}
    for L in (24, 128):
        U = R_kr._build_regressor(u, L)
        G, b, yy = U.T @ U, U.T @ y, float(y @ y)
        d[f"G_{L}"], d[f"b_{L}"], d[f"yy_{L}"] = G, b, np.array([yy])
        for kname, th0 in theta0.items():
            pts = [np.array(th0 + [math.log(1e-2)])]
            for _ in range(7):
                pts.append(pts[0] + np.concatenate([0.6 * rng.standard_normal(len(th0)), [2.0 * rng.standard_normal()]]))
            pts = np.array(pts)
            d[f"theta_{kname}_{L}"] = pts
            d[f"nll_{kname}_{L}"] = np.array([R_kr._nll(p, G, b, yy, len(u), L, kname) for p in pts])
            d[f"K_{kname}_{L}"] = R_kr._build_kernel(L, kname, pts[1][:len(th0)])
            g_hat, sig2, K = R_kr._posterior_mean_g(pts[0], G, b, L, kname)
            d[f"ghat_{kname}_{L}"], d[f"sig2_{kname}_{L}"] = g_hat, np.array([sig2])
    tmp = Path(tempfile.mkdtemp(prefix="b200id_golden_kr_"))
    mat = synth.write_mat(tmp / "io.mat", t, u + 0.3, y - 0.1)
    for kname, L, src in (("dc", 48, "synthetic"), ("ss", 32, "synthetic"), ("si", 24, "synthetic"), ("dc", 64, "mat")):
        cfg = R_kr.FIRConfig(io_mat=(mat if src == "mat" else tmp / "missing.mat"), out_dir=tmp / f"out_{kname}_{L}_{src}",
                             kernel=kname, L=L, multi_starts=2, maxiter=60)
        res = quiet(R_kr.run_fir_experiment, cfg)
        co = np.load(res["coeff_npz"])
        tag = f"run_{kname}_{L}_{src}"
        d[f"{tag}_ghat"], d[f"{tag}_K"] = co["g_hat"], co["K"]
        d[f"{tag}_metrics"] = np.array([res["rmse"], res["nrmse"], res["r2"], res["fit_percent"], res["sigma2"]])
    d["mat_t"], d["mat_u"], d["mat_y"] = t, u + 0.3, y - 0.1
    np.savez_compressed(out / "kr_fir.npz", **d)
    print("kr_fir.npz", len(d), "arrays")


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap_.add_argument("--only", default="")
    a = ap_.parse_args()
    out = Path(a.out); out.mkdir(parents=True, exist_ok=True)
    todo = {"frf": golden_frf, "gp": golden_gp, "fir": golden_fir, "pipeline": golden_pipeline, "fourier": golden_fourier,
            "pipeline_polar": golden_pipeline_polar, "run_baseline": golden_run_baseline, "matern_nu": golden_matern_nu, "kr": golden_kr}
    for name, fn in todo.items():
        if a.only and name not in a.only.split(","):
            continue
        fn(out)


if __name__ == "__main__":
    main()
