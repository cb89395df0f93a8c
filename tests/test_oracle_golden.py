"""Pin the NumPy oracle against vectors produced by the reference itself (tests/golden/)."""
import numpy as np
import pytest

from oracle import frf as ofrf, gp as ogp, fir as ofir

KERNEL_NAMES = [k for k in ogp.KERNELS if k != "matern_nu"]      # general-nu Matern: tests/test_matern_nu*.py


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ------------------------------------------------------------------ FRF
def test_freq_grid(golden):
    g = golden("frf")
    hz, w = ofrf.freq_grid(37, -1.0, 2.3)
    assert np.array_equal(hz, g["grid_hz_37"]) and np.array_equal(w, g["grid_w_37"])
    assert np.array_equal(ofrf.freq_grid(50)[1], g["w"])


def test_demod_bit_exact(golden):
    g = golden("frf")
    t, u, y, w = g["t"], g["u"], g["y"], g["w"]
    assert np.array_equal(ofrf.demod_trapz(t, u, w), g["U"])
    assert np.array_equal(ofrf.demod_trapz(t, y, w, threads=4), g["Y"])
    assert np.array_equal(ofrf.demod_trapz(t, u, w, 1.5, True), g["U_drop"])
    assert np.array_equal(ofrf.demod_trapz(t, y + 0.3, w, 0.0, False), g["Y_nomean"])
    assert np.array_equal(ofrf.demod_trapz(g["t_jit"], u, w), g["U_jit"])


def test_demod_errors():
    t = np.arange(10) * 0.002
    with pytest.raises(ValueError, match="Not enough samples"):
        ofrf.demod_trapz(t, t, np.array([1.0]), drop_seconds=1.0)
    with pytest.raises(ValueError, match="Non-positive record length"):
        ofrf.demod_trapz(np.zeros(4), np.ones(4), np.array([1.0]))


def test_cross_power(golden):
    g = golden("frf")
    assert np.array_equal(ofrf.cross_power([g["U"]], [g["Y"]]), g["G"])
    assert np.array_equal(ofrf.cross_power([g["U"], g["U2"]], [g["Y"], g["Y2"]]), g["G2"])
    Uz = g["U"].copy(); Uz[[3, 17]] = 0.0
    got = ofrf.cross_power([Uz], [g["Y"]])
    assert np.array_equal(np.isnan(got), np.isnan(g["G_nan"])) and np.isnan(got[[3, 17]]).all()
    with pytest.raises(ValueError):
        ofrf.cross_power([], [])


def test_estimate_frf_records(golden):
    g = golden("frf")
    w, G = ofrf.estimate_frf([(g["t"], g["u"], g["y"]), (g["t"], g["u2"], g["y2"])], 50)
    df = g["df_2files"]
    assert np.array_equal(w, df[:, 0]) and np.array_equal(G.real, df[:, 2]) and np.array_equal(G.imag, df[:, 3])
    w, G = ofrf.estimate_frf([(g["t"], g["u"], g["y"])], 30, drop_seconds=1.0, time_duration=7.0)
    df = g["df_window"]
    assert np.array_equal(w, df[:, 0]) and np.array_equal(G.real, df[:, 2]) and np.array_equal(G.imag, df[:, 3])


# ------------------------------------------------------------------ GP
@pytest.mark.parametrize("name", KERNEL_NAMES)
def test_gram_bit_exact(golden, name):
    g = golden("gp")
    X, Xv, Xi, Xj = g["X"], g["Xv"], g["Xi"], g["Xj"]
    for q, th in enumerate(g[f"theta_{name}"]):
        for got, key in ((ogp.gram(name, th, X[:24]), f"gram_{name}_{q}"),
                         (ogp.gram(name, th, Xv[:10], X[:24]), f"gramx_{name}_{q}"),
                         (ogp.gram(name, th, Xi), f"grami_{name}_{q}"),
                         (ogp.gram(name, th, Xj, Xi), f"gramij_{name}_{q}")):
            assert np.array_equal(got, g[key], equal_nan=True), key


@pytest.mark.parametrize("name", KERNEL_NAMES)
def test_bounds_and_grids(golden, name):
    g = golden("gp")
    assert np.array_equal(np.array(ogp.KERNELS[name][2]), g[f"bounds_{name}"])
    assert np.array_equal(np.stack(ogp.default_grids(name)), g[f"grid_{name}"])


@pytest.mark.parametrize("name", ogp.BASELINE_KERNELS)
def test_candidate_list_and_metrics(golden, name):
    g = golden("gp")
    C = ogp.grid_candidates(name, 1500)
    assert np.array_equal(C, g[f"cand_{name}"])
    X, y, Xv, yv = g["X"], g["yr"], g["Xv"], g["yv_r"]
    ym, ys = float(g["yr_mean"][0]), float(g["yr_scale"][0])
    idx = np.unique(np.r_[np.arange(0, len(C), 7), len(C) - 1])
    with np.errstate(all="ignore"):
        nll = np.array([ogp.candidate_metric(name, C[i], X, y, 1e-6) for i in idx])
        vr = np.array([ogp.candidate_metric(name, C[i], X, y, 1e-6, Xv, yv, ym, ys) for i in idx])
        vr0 = np.array([ogp.candidate_metric(name, C[i], X, y, 1e-6, Xv, yv) for i in range(0, 200, 5)])
        nll0 = np.array([ogp.candidate_metric(name, C[i], X, y, 0.0) for i in idx])
    assert np.array_equal(nll, g[f"nll_{name}"][idx], equal_nan=True)
    np.testing.assert_allclose(vr, g[f"vrmse_{name}"][idx], rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(vr0, g[f"vrmse_noscaler_{name}"][0:200:5], rtol=1e-13, equal_nan=True)
    assert np.array_equal(nll0, g[f"nll0_{name}"][idx], equal_nan=True)


@pytest.mark.parametrize("name", ["matern52", "rbf", "ss2", "di", "stable_spline"])
@pytest.mark.parametrize("mode,tag,noise", [("nll", "n6", 1e-6), ("val", "n6", 1e-6), ("val", "n0", 0.0)])
def test_grid_selection(golden, name, mode, tag, noise):
    g = golden("gp")
    kw = {}
    if mode == "val":
        kw = dict(Xv=g["Xv"], yv=g["yv_r"], y_mean=float(g["yr_mean"][0]), y_scale=float(g["yr_scale"][0]))
    with np.errstate(all="ignore"):
        best, _, metrics, _ = ogp.grid_search(name, g["X"], g["yr"], noise, 5000, **kw)
    assert np.array_equal(best, g[f"sel_{name}_{mode}_{tag}"])


def test_first_strict_min_skips_nan():
    assert ogp.first_strict_min(np.array([np.nan, 3.0, np.nan, 3.0, 2.0, 2.0])) == (4, 2.0)
    assert ogp.first_strict_min(np.array([np.nan, np.nan]))[0] == -1


@pytest.mark.parametrize("name", KERNEL_NAMES)
def test_fit_predict(golden, name):
    g = golden("gp")
    th = g[f"theta_{name}"][0]
    X, y = g["X"], g["yr"]
    alpha, Kinv, _ = ogp.fit(name, th, X, y, 1e-6)   # ss1/exp/... legitimately hit the pinv branch
    assert np.array_equal(alpha, g[f"fit_alpha_{name}"]) and np.array_equal(Kinv, g[f"fit_Kinv_{name}"])
    m, s = ogp.predict(name, th, X, alpha, Kinv, g["Xs_new"], True)
    assert np.array_equal(m, g[f"fit_mean_new_{name}"]) and np.array_equal(s, g[f"fit_std_new_{name}"])
    m, s = ogp.predict(name, th, X, alpha, Kinv, X, True)
    assert np.array_equal(m, g[f"fit_mean_train_{name}"]) and np.array_equal(s, g[f"fit_std_train_{name}"])


def test_fit_pinv_fallback(golden):
    """SS1 on standardised inputs is indefinite -> Cholesky fails -> pinv branch (gpr_fitting.py:84-87);
    the rank-5 DC Gram + 1e-10 jitter still factorises (alpha ~ 1e10), i.e. no fallback."""
    g = golden("gp")
    alpha, Kinv, used_pinv = ogp.fit("ss1", g["theta_ss1"][0], g["X"], g["yr"], 1e-6)
    assert used_pinv
    assert np.array_equal(alpha, g["fit_alpha_ss1"]) and np.array_equal(Kinv, g["fit_Kinv_ss1"])
    alpha, Kinv, used_pinv = ogp.fit("dc", [0.9, 1.0, 0.5], g["X"], g["yr"], 0.0)
    assert not used_pinv
    assert np.array_equal(alpha, g["pinv_alpha_dc"]) and np.array_equal(Kinv, g["pinv_Kinv_dc"])


def test_standardize_matches_sklearn(golden):
    g = golden("gp")
    z, m, s = ogp.standardize(g["G"].real)
    np.testing.assert_allclose(z, g["yr"], rtol=0, atol=1e-14)
    np.testing.assert_allclose([m, s], [g["yr_mean"][0], g["yr_scale"][0]], rtol=1e-14)


def test_restart_population_rule():
    P = ogp.restart_population("dc", 64, 1e-6, seed=42)
    b = np.array(ogp.KERNELS["dc"][2])
    assert P.shape == (64, 3) and (P >= b[:, 0]).all() and (P <= b[:, 1]).all()
    assert np.array_equal(P, ogp.restart_population("dc", 64, 1e-6, seed=42))


# ------------------------------------------------------------------ FIR
def test_hermitian_idft(golden):
    g = golden("fir")
    X = ofir.hermitian_extend(g["G"][:9])
    assert np.array_equal(X, g["herm_small"])
    taps, full = ofir.idft_taps(X, 11)
    assert np.array_equal(taps, g["taps_small"]) and np.array_equal(full, g["hfull_small"])
    taps, _ = ofir.idft_taps(ofir.hermitian_extend(g["G_uni"]), 1024)
    assert np.array_equal(taps, g["taps_paper"])
    with pytest.raises(ValueError, match="exceeds available length"):
        ofir.idft_taps(X, 18)


def test_paper_fir_and_validation(golden):
    from b200id import synth
    g = golden("fir")
    taps, w_uni, G_uni, _ = ofir.paper_fir(g["w"], g["G"], predict=synth.plant_response)
    assert np.array_equal(taps, g["pipe_taps"])
    m, yp = ofir.fir_validate(g["u"], g["y"], taps)
    assert np.array_equal(yp, g["y_pred"])
    assert np.array_equal([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]], g["pipe_metrics"][:4])
    assert np.array_equal([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]], g["validate_metrics"])
    taps_i, *_ = ofir.paper_fir(g["w"], g["G"], predict=None, fir_order=300)
    assert np.array_equal(taps_i, g["pipe_interp_taps"])
    m, _ = ofir.fir_validate(g["u"], g["y"], taps_i)
    assert np.array_equal([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]], g["pipe_interp_metrics"])
    m, _ = ofir.fir_validate(g["u"], g["y"], taps[:64], detrend=True)
    assert np.array_equal([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]], g["validate_detrend_metrics"])


# ------------------------------------------------------------------ Fourier estimator
def test_fourier_oracle_bit_exact(golden):
    from oracle import fourier as ofou
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "gauss-process_b200" / "b200id"))
    import synth
    g = golden("fourier")
    for tag, n in (("even", 6000), ("odd", 4999)):
        t, u, y = synth.make_record(n, 50, "multisine", seed=11)
        for wname in (("hann", "none", "blackman") if tag == "even" else ("hann",)):
            f, G = ofou.transfer_function(t, u, y, None if wname == "none" else wname)
            assert np.array_equal(f, g[f"f_{tag}"]) and np.array_equal(G, g[f"G_{tag}_{wname}"]), (tag, wname)
        assert np.array_equal(ofou.spectrum(t, u, "hann")[1], g[f"U_{tag}_hann"])
    r1 = synth.make_record(6000, 50, "multisine", seed=0)
    r2 = synth.make_record(5555, 50, "multisine", seed=3)

    def table(grid, G):
        w = 2.0 * np.pi * grid
        return np.c_[w, w / (2.0 * np.pi), G.real, G.imag, np.abs(G), np.angle(G)]

    assert np.array_equal(table(*ofou.estimate([r1, r2], 40, 0.1, 100.0)), g["df_2files"])
    assert np.array_equal(table(*ofou.estimate([r1], 25, 0.5, 60.0, drop_seconds=1.0, time_duration=7.0, window="none")), g["df_window"])
    assert np.array_equal(table(*ofou.estimate([r1], 30, 0.0, 1e9)), g["df_top"])
    assert np.array_equal(table(*ofou.estimate([r2], 17, 1.0, 33.0, window="blackman")), g["df_blackman"])
