"""General-nu Matern kernel (reference src/gpr/kernels.py:131-135, the scipy.special.kv branch) on the device:
Gram matrices, per-candidate metrics over the reference's own scan list, grid-search selection, fit / predict and a
whole run_gp_pipeline (--kernel matern --nu 0.8), all against vectors produced by the reference itself
(tests/golden/matern_nu.npz, oracle/make_golden.py::golden_matern_nu)."""
import numpy as np
import pytest

from oracle import gp as ogp
from parity_util import rel_each, rel_max, record

pytestmark = pytest.mark.gpu
KID = 13


@pytest.fixture(scope="module")
def gp():
    from b200id import gp as mod
    return mod


def test_gram_against_reference(gp, golden):
    g = golden("matern_nu")
    worst = 0.0
    for a, nu in enumerate(g["nus"]):
        for q, (ls, var) in enumerate(g["thetas"]):
            th = [ls, var, nu]
            for got, key in ((gp.gram(KID, th, g["X"][:24]), f"gram_{a}_{q}"),
                             (gp.gram(KID, th, g["Xv"][:10], g["X"][:24]), f"gramx_{a}_{q}"),
                             (gp.gram(KID, th, g["Xw"]), f"gramw_{a}_{q}")):
                ref = g[key]
                np.testing.assert_allclose(got, ref, rtol=2e-13, atol=1e-290, err_msg=key)
                nz = ref != 0
                worst = max(worst, float(np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz]))))
    record("gp.gram.matern_nu.rel_each", worst)


@pytest.mark.parametrize("a", [1, 4])
def test_candidate_metrics_and_selection(gp, golden, a):
    """NLL and validation RMSE of the 900 grid candidates (1e-8 where cond(K) < 1e6, 100 cond eps beyond) and the
    selection of GaussianProcessRegressor.fit(use_grid_search=True) through the drop-in classes."""
    import contextlib, io
    from src.gpr.kernels import MaternKernel, create_kernel
    from src.gpr.gpr_fitting import GaussianProcessRegressor
    g = golden("matern_nu")
    nu = float(g["nus"][a])
    C, X, y = g[f"cand_{a}"], g["X"], g["yr"]
    th3 = np.column_stack([C, np.full(len(C), nu)])
    with np.errstate(all="ignore"):
        cond = np.array([np.linalg.cond(ogp.gram("matern_nu", t, X) + 1e-6 * np.eye(len(X))) for t in th3])
    for tag, got, ref in (("nll", gp.batch_eval(KID, th3, X, y, 1e-6), g[f"nll_{a}"]),
                          ("vrmse", gp.batch_eval(KID, th3, X, y, 1e-6, Xv=g["Xv"], yv=g["yv_r"], y_mean=float(g["yr_mean"][0]),
                                                  y_scale=float(g["yr_scale"][0])), g[f"vrmse_{a}"])):
        ok_g, ok_r = np.isfinite(got) & (got != 1e10), np.isfinite(ref) & (ref != 1e10)
        assert int(np.sum((ok_g != ok_r) & (cond < 1e13))) == 0, tag
        both = ok_g & ok_r
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        well = both & (cond < 1e6)
        record(f"gp.metric.matern_nu{nu:g}.{tag}.rel_max_well_conditioned", float(err[well].max()) if well.any() else 0.0)
        assert not well.any() or err[well].max() < 1e-8, (tag, float(err[well].max()))
        assert np.sum(both & (err > np.maximum(1e-8, 100.0 * cond * 2.2e-16))) <= max(1, int(0.01 * both.sum())), tag
    for mode in ("nll", "val"):
        gpr = GaussianProcessRegressor(kernel=create_kernel("matern", nu=nu), noise_variance=1e-6)
        assert isinstance(gpr.kernel, MaternKernel) and gpr.kernel.kernel_id == KID
        kw = dict(validation_X=g["Xv"], validation_y=g["yv_r"], y_scaler=_Scaler(g)) if mode == "val" else {}
        with contextlib.redirect_stdout(io.StringIO()):
            gpr.fit(X, y, optimize=True, use_grid_search=True, **kw)
        assert np.array_equal(gpr.kernel.get_params(), g[f"sel_{a}_{mode}"]), (mode, gpr.kernel.get_params(), g[f"sel_{a}_{mode}"])
        mean, std = gpr.predict(X, return_std=True)
        scale = np.sqrt(max(gpr.kernel.params["variance"], 1e-300))
        assert record(f"gp.sel.matern_nu{nu:g}.{mode}.mean", rel_max(mean, g[f"sel_mean_{a}_{mode}"])) < 1e-6
        assert float(np.max(np.abs(std - g[f"sel_std_{a}_{mode}"]))) / scale < 1e-4
    gpr = GaussianProcessRegressor(kernel=MaternKernel(length_scale=0.7, variance=1.3, nu=nu), noise_variance=1e-6)
    gpr.fit(X, y, optimize=False)
    assert record(f"gp.fit.matern_nu{nu:g}.alpha", rel_max(gpr.alpha, g[f"fit_alpha_{a}"])) < 1e-7
    mean, std = gpr.predict(g["Xs_new"], return_std=True)
    assert record(f"gp.fit.matern_nu{nu:g}.mean", rel_max(mean, g[f"fit_mean_{a}"])) < 1e-8
    assert float(np.max(np.abs(std - g[f"fit_std_{a}"]))) / np.sqrt(1.3) < 1e-5


class _Scaler:
    """StandardScaler stand-in carrying the golden statistics (mean_, scale_)."""
    with_mean = with_std = True

    def __init__(self, g):
        self.mean_, self.scale_ = g["yr_mean"], g["yr_scale"]

    def inverse_transform(self, a):
        return np.asarray(a) * self.scale_ + self.mean_


def test_whole_pipeline_matern_nu08(tmp_path, golden):
    """run_gp_pipeline --kernel matern --nu 0.8 --optimize --grid-search (gp_pipeline.py:282-286) through the drop-in
    surface: selected hyperparameters identical, taps and FIR metrics as the reference's run."""
    import contextlib, io
    import pipeline_flow
    from b200id import synth
    g = golden("matern_nu")
    n, nd = int(g["n"]), int(g["nd"])
    tt, ut, yt = synth.make_record(n, nd, "multisine", seed=0)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=1)
    ptrain = synth.write_mat(tmp_path / "train.mat", tt, ut, yt)
    pval = synth.write_mat(tmp_path / "val.mat", tv, uv, yv)
    np.random.seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        res = pipeline_flow.run(ptrain, pval, tmp_path / "m08", kernel="matern", nd=nd, noise=1e-6, optimize=True, grid=True,
                                kernel_kwargs={"nu": 0.8})
    assert np.array_equal(np.array(res["kernel_params"]["real"]), g["pipe_theta_real"])
    assert np.array_equal(np.array(res["kernel_params"]["imag"]), g["pipe_theta_imag"])
    f = res["fir_extraction"]
    e_taps = record("dropin.pipeline.matern_nu08.taps.rel_max", rel_max(f["fir_coefficients"], g["pipe_taps"]))
    e_fir = record("dropin.pipeline.matern_nu08.fir_metrics.rel_each",
                   rel_each(np.array([f["rmse"], f["fit_percent"], f["r2"]]), g["pipe_fir"]))
    assert e_taps < 1e-7 and e_fir < 1e-7
