"""The device-resident hot path (b200id.pipeline.identify_device -- what bench.py times) against the oracle."""
import numpy as np
import pytest

from oracle import frf as ofrf, gp as ogp, fir as ofir
from parity_util import record

pytestmark = pytest.mark.gpu


def test_smoke_entry_runs_green():
    import __graft_entry__
    __graft_entry__.smoke()


@pytest.mark.parametrize("kernel,n,nd", [("matern52", 30_000, 50), ("rbf", 12_345, 40)])
def test_identify_device_validation_metric_against_oracle(kernel, n, nd):
    """FRF -> validation-RMSE grid search (150-bin validation FRF) -> fit -> FIR -> validation, all resident."""
    import torch
    from b200id import pipeline as P, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    noise = 1e-6
    t, u, y = synth.make_record(n, nd, "multisine", seed=3)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=4)
    skip = min(10, 1024 // 10)
    train = [(to_device(t, dev), to_device(u, dev), to_device(y, dev), float(t[-1] - t[0]), float(t[0]))]
    val = (to_device(tv, dev), to_device(uv, dev), to_device(yv, dev), float(tv[-1] - tv[0]), float(yv[skip]), float(tv[0]))
    cands = P.CandidateSet.build(kernel, P.grid_candidates(kernel), dev)
    res = P.identify_device(train, val, nd, kernel, noise, cands, use_validation_metric=True)
    torch.cuda.synchronize()

    w, G = ofrf.estimate_frf([(t, u, y)], nd)
    wv, Gv = ofrf.estimate_frf([(tv, uv, yv)], 150)
    assert record(f"pipeline.{kernel}.frf", float(np.max(np.abs(res.G.cpu().numpy() - G) / np.abs(G)))) < 1e-9
    X, xm, xs = ogp.standardize(np.log10(w))
    Xv = (np.log10(wv) - xm) / xs
    sel = []
    for target, vy in ((G.real, Gv.real), (G.imag, Gv.imag)):
        z, m, s = ogp.standardize(target)
        best, metric, _, _ = ogp.grid_search(kernel, X, z, noise, Xv=Xv, yv=vy, y_mean=m, y_scale=s)
        sel.append((best, metric, z, m, s))
    assert np.array_equal(sel[0][0], res.theta_real) and np.array_equal(sel[1][0], res.theta_imag)
    assert abs(res.metric_real - sel[0][1]) <= 1e-8 * abs(sel[0][1]) and abs(res.metric_imag - sel[1][1]) <= 1e-8 * abs(sel[1][1])
    fits = [ogp.fit(kernel, b, X, z, noise) + (b, m, s) for b, _, z, m, s in sel]

    def predict(w_new):
        Xn = (np.log10(w_new) - xm) / xs
        parts = [ogp.predict(kernel, b, X, a, K, Xn) * s + m for a, K, _, b, m, s in fits]
        return parts[0] + 1j * parts[1]

    taps, *_ = ofir.paper_fir(w, G, predict=predict)
    m, _ = ofir.fir_validate(uv, yv, taps)
    e_taps = record(f"pipeline.{kernel}.taps", float(np.max(np.abs(res.taps.cpu().numpy() - taps)) / np.max(np.abs(taps))))
    # the fitted mean is only defined to cond(K) * eps (RBF Gram matrices at noise 1e-6 reach cond ~ 1e9)
    cond = max(float(np.linalg.cond(ogp.gram(kernel, b, X) + (noise + 1e-10) * np.eye(len(X)))) for b, *_ in sel)
    tol = max(1e-6, 100.0 * cond * 2.2e-16)
    record(f"pipeline.{kernel}.cond", cond)
    assert e_taps < tol and abs(res.fir["rmse"] - m["rmse"]) <= tol * m["rmse"]
