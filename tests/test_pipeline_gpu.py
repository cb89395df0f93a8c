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


def _host_vs_device(t, u, y, tv, uv, yv, nd, kernel="matern52", n_chunks=6):
    import torch
    from b200id import pipeline as P
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    skip = min(10, 1024 // 10)
    train = [(to_device(t, dev), to_device(u, dev), to_device(y, dev), float(t[-1] - t[0]), float(t[0]))]
    val = (to_device(tv, dev), to_device(uv, dev), to_device(yv, dev), float(tv[-1] - tv[0]), float(yv[skip]), float(tv[0]))
    cands = P.CandidateSet.build(kernel, P.grid_candidates(kernel), dev)
    ref = P.identify_device(train, val, nd, kernel, 1e-6, cands, use_validation_metric=True)
    got = P.identify_host([(t, u, y)], (tv, uv, yv), nd, kernel, 1e-6, cands, use_validation_metric=True, n_chunks=n_chunks)
    torch.cuda.synchronize()
    return ref, got


@pytest.mark.parametrize("n,pinned", [(1 << 21, True), (50_001, False)])
def test_identify_host_matches_identify_device(n, pinned):
    """One call on HOST records (each uploaded once, projected piece by piece while it lands, FIR pass on the
    resident copy) gives what the device-resident pass gives: same selection, FRF / taps / metrics to the order
    of the final additions."""
    import torch
    from b200id import synth
    nd = 40
    t, u, y = synth.make_record(n, nd, "multisine", seed=11)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=12)
    if pinned:
        keep = [torch.from_numpy(a).pin_memory() for a in (t, u, y, tv, uv, yv)]
        t, u, y, tv, uv, yv = [k.numpy() for k in keep]
    ref, got = _host_vs_device(t, u, y, tv, uv, yv, nd)
    assert np.array_equal(ref.theta_real, got.theta_real) and np.array_equal(ref.theta_imag, got.theta_imag)
    Gr, Gg = ref.G.cpu().numpy(), got.G.cpu().numpy()
    # the resident pass projects whole records (bin-adaptive form: the rounding term of strongly excited bins is skipped
    # within a 4e-10 budget per spectrum), the host pass projects pieces as they land (form 1, every bin's rounding term)
    assert record(f"identify_host.frf.{n}", float(np.max(np.abs(Gg - Gr) / np.abs(Gr)))) < 1e-9
    tr, tg = ref.taps.cpu().numpy(), got.taps.cpu().numpy()
    assert record(f"identify_host.taps.{n}", float(np.max(np.abs(tg - tr)) / np.max(np.abs(tr)))) < 1e-8
    assert abs(got.fir["rmse"] - ref.fir["rmse"]) <= 1e-8 * ref.fir["rmse"]
    assert abs(got.metric_real - ref.metric_real) <= 1e-8 * abs(ref.metric_real)


def test_identify_host_recovers_from_a_wrong_timebase_guess():
    """The head of the record sits on the grid, a later stretch does not: the uniform-kernel guess made on the
    host is overturned by the device check over the whole record and the pass is redone with the general kernel,
    i.e. the result is the device-resident pass's (which probes the whole record first)."""
    from b200id import synth
    n, nd = 300_000, 30
    t, u, y = synth.make_record(n, nd, "multisine", seed=21)
    tv, uv, yv = synth.make_record(n, nd, "square", seed=22)
    t = t.copy()
    rng = np.random.default_rng(5)
    t[n // 2: n // 2 + 5000] += rng.uniform(0.0, 2e-4, 5000)        # w_max * delta ~ 0.25 rad, still monotone
    assert np.all(np.diff(t) > 0)
    ref, got = _host_vs_device(t, u, y, tv, uv, yv, nd)
    Gr, Gg = ref.G.cpu().numpy(), got.G.cpu().numpy()
    assert float(np.max(np.abs(Gg - Gr) / np.abs(Gr))) < 1e-12
    assert np.array_equal(ref.theta_real, got.theta_real) and np.array_equal(ref.theta_imag, got.theta_imag)
    w, G = ofrf.estimate_frf([(t, u, y)], nd)
    assert float(np.max(np.abs(Gg - G) / np.abs(G))) < 1e-9


def test_identify_host_two_training_records_and_no_search():
    """Cross-power average over two host records (transform.py:124-150) and the no-search branch (fixed theta,
    validation record used by the FIR pass only, i.e. uploaded but not projected)."""
    import torch
    from b200id import pipeline as P, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    n, nd = 120_000, 30
    recs = [synth.make_record(n, nd, "multisine", seed=s) for s in (31, 32)]
    tv, uv, yv = synth.make_record(n, nd, "square", seed=33)
    skip = min(10, 1024 // 10)
    train_d = [(to_device(t, dev), to_device(u, dev), to_device(y, dev), float(t[-1] - t[0]), float(t[0])) for t, u, y in recs]
    val_d = (to_device(tv, dev), to_device(uv, dev), to_device(yv, dev), float(tv[-1] - tv[0]), float(yv[skip]), float(tv[0]))
    cands = P.CandidateSet.build("matern52", P.grid_candidates("matern52"), dev)
    for kwargs in ({"candidates": cands}, {"candidates": None, "theta0": [0.7, 1.3]}):
        ref = P.identify_device(train_d, val_d, nd, "matern52", 1e-6, **kwargs)
        got = P.identify_host(recs, (tv, uv, yv), nd, "matern52", 1e-6, **kwargs)
        torch.cuda.synchronize()
        Gr, Gg = ref.G.cpu().numpy(), got.G.cpu().numpy()
        assert float(np.max(np.abs(Gg - Gr) / np.abs(Gr))) < 1e-11
        assert np.array_equal(ref.theta_real, got.theta_real) and np.array_equal(ref.theta_imag, got.theta_imag)
        # the two passes add the FRF pieces in a different order (1e-11 in G); the fitted mean amplifies that by
        # cond(K + noise I), as in test_identify_device_validation_metric_against_oracle
        assert abs(got.fir["rmse"] - ref.fir["rmse"]) <= 1e-6 * ref.fir["rmse"]
    # no validation record at all: FRF -> fit -> taps only
    got = P.identify_host(recs, None, nd, "matern52", 1e-6, candidates=None, theta0=[0.7, 1.3])
    ref = P.identify_device(train_d, None, nd, "matern52", 1e-6, candidates=None, theta0=[0.7, 1.3])
    tr, tg = ref.taps.cpu().numpy(), got.taps.cpu().numpy()
    assert float(np.max(np.abs(tg - tr)) / np.max(np.abs(tr))) < 1e-6 and got.fir == {}


def test_argmin_pack_and_pick_kernels_follow_the_scan_rule():
    """b200id_argmin_pack / b200id_argmin_pick (the two launches either side of the ranks' argmin exchange) against
    b200id.dist.pick_first_strict_min, the host statement of grid_search.py:186-196 over a sharded list: ties go to the
    smallest global index, NaN metrics and empty shards are skipped, nothing finite anywhere gives index -1."""
    import ctypes as C
    import torch
    from b200id import _lib, dist as bdist
    from b200id.runtime import require_cuda, to_device, ptr, stream_ptr, empty
    dev = require_cuda()
    rng = np.random.default_rng(9)
    n_cand, W, S = 57, 4, 3
    full = rng.uniform(0.1, 5.0, (n_cand, 3))
    full_d = to_device(full, dev)
    for case in range(40):
        # pack: local (best, index) -> (best, global index) through a rank's index map
        gmap = np.sort(rng.choice(n_cand, 15, replace=False)).astype(np.int64)
        lidx = rng.integers(-1, 15, S).astype(np.float64)
        lbest = np.where(lidx >= 0, rng.integers(0, 4, S).astype(np.float64), np.inf)
        pend = to_device(np.c_[lbest, lidx].ravel(), dev)
        mine = empty(2 * S, dev)
        _lib.call("b200id_argmin_pack", ptr(pend), C.c_int(S), ptr(torch.as_tensor(gmap, device=dev)), ptr(mine), stream_ptr())
        got = mine.cpu().numpy().reshape(S, 2)
        want_g = np.where(lidx >= 0, gmap[np.maximum(lidx.astype(int), 0)], -1)
        assert np.array_equal(got[:, 0], lbest) and np.array_equal(got[:, 1], want_g)
        _lib.call("b200id_argmin_pack", ptr(pend), C.c_int(S), ptr(None), ptr(mine), stream_ptr())
        assert np.array_equal(mine.cpu().numpy().reshape(S, 2)[:, 1], lidx)
        # pick: gathered [W][S][2] -> winner per search
        m = rng.integers(0, 3, (W, S)).astype(np.float64)
        g = rng.permuted(np.arange(n_cand))[:W * S].reshape(W, S).astype(np.float64)
        m[rng.random((W, S)) < 0.2] = np.nan
        g[rng.random((W, S)) < 0.2] = -1.0
        if case % 7 == 0:
            g[:, 0] = -1.0                                   # search 0: every shard empty
        if case % 5 == 0:
            m[:, 1] = 1e10                                   # search 1: every candidate failed -> smallest index wins
        tab = to_device(np.stack([m, g], axis=2).ravel(), dev)
        out = empty(5 * S, dev)
        _lib.call("b200id_argmin_pick", ptr(tab), C.c_int(W), C.c_int(S), ptr(full_d), ptr(out), stream_ptr())
        out = out.cpu().numpy().reshape(S, 5)
        for j in range(S):
            ei, eb = bdist.pick_first_strict_min([(float(m[r, j]), int(g[r, j])) for r in range(W)])
            assert int(out[j, 3]) == ei, (case, j, out[j], ei)
            assert out[j, 4] == eb or (ei < 0 and np.isinf(out[j, 4]))
            assert np.array_equal(out[j, :3], full[max(ei, 0)])
