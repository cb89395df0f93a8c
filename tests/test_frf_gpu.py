"""GPU parity of K1 (FRF projection) through the C ABI, against the oracle and the reference's
golden vectors.  Tolerance from BASELINE.json north_star: FRF G(jw_k) within 1e-9 relative."""
import numpy as np
import pytest

from oracle import frf as ofrf
from parity_util import rel_each, rel_max, record

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def frf():
    from b200id import frf as mod
    return mod


def test_golden_uniform(frf, golden):
    g = golden("frf")
    U, Y = frf.demod_pair(g["t"], g["u"], g["y"], g["w"])
    eU = record("frf.golden.U.rel_each", rel_each(U, g["U"]))
    eY = record("frf.golden.Y.rel_each", rel_each(Y, g["Y"]))
    assert eU < TOL and eY < TOL
    G = frf.cross_power([U], [Y])
    assert record("frf.golden.G.rel_each", rel_each(G, g["G"])) < TOL


def test_golden_single_signal_drop_and_nomean(frf, golden):
    g = golden("frf")
    assert rel_each(frf.demod(g["t"], g["u"], g["w"]), g["U"]) < TOL
    assert record("frf.golden.U_drop.rel_each", rel_each(frf.demod(g["t"], g["u"], g["w"], 1.5, True), g["U_drop"])) < TOL
    assert record("frf.golden.Y_nomean.rel_each",
                  rel_each(frf.demod(g["t"], g["y"] + 0.3, g["w"], 0.0, False), g["Y_nomean"])) < TOL


def test_golden_nonuniform_timebase(frf, golden):
    g = golden("frf")
    U, Y = frf.demod_pair(g["t_jit"], g["u"], g["y"], g["w"])
    assert record("frf.golden.U_jit.rel_each", rel_each(U, g["U_jit"])) < TOL
    assert rel_each(Y, g["Y_jit"]) < TOL


def test_cross_power_records_and_nan(frf, golden):
    g = golden("frf")
    G2 = frf.cross_power([g["U"], g["U2"]], [g["Y"], g["Y2"]])
    assert rel_each(G2, g["G2"]) < 1e-14
    Uz = g["U"].copy(); Uz[[3, 17]] = 0.0
    Gn = frf.cross_power([Uz], [g["Y"]])
    assert np.array_equal(np.isnan(Gn), np.isnan(g["G_nan"]))
    ok = ~np.isnan(Gn)
    assert rel_each(Gn[ok], g["G_nan"][ok]) < 1e-14
    with pytest.raises(ValueError):
        frf.cross_power([], [])
    Gm, _, _ = frf.estimate_records([(g["t"], g["u"], g["y"]), (g["t"], g["u2"], g["y2"])], g["w"])
    assert rel_each(Gm.real, g["df_2files"][:, 2]) < TOL and rel_each(Gm.imag, g["df_2files"][:, 3]) < TOL


def test_errors_match_reference(frf):
    t = np.arange(10) * 0.002
    with pytest.raises(ValueError, match="Not enough samples"):
        frf.demod(t, t, np.array([1.0]), drop_seconds=1.0)
    with pytest.raises(ValueError, match="Non-positive record length"):
        frf.demod(np.zeros(4), np.ones(4), np.array([1.0]))


@pytest.mark.parametrize("n,nd,kind", [(200_000, 100, "multisine"), (300_001, 150, "square"), (4097, 7, "multisine"),
                                       (50_000, 600, "multisine")])
def test_oracle_seeded(frf, n, nd, kind):
    """Ragged sizes, nd above one CTA's bin group (600 > 512), square-wave (broadband) input."""
    from b200id import synth
    t, u, y = synth.make_record(n, min(nd, 100), kind, seed=11)
    _, w = ofrf.freq_grid(nd)
    U, Y = frf.demod_pair(t, u, y, w)
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e = max(rel_each(U, Ur), rel_each(Y, Yr))
    record(f"frf.oracle.{kind}.n{n}.nd{nd}.rel_each", e)
    assert e < TOL
    G, Gr = frf.cross_power([U], [Y]), ofrf.cross_power([Ur], [Yr])
    assert record(f"frf.oracle.{kind}.n{n}.nd{nd}.G.rel_each", rel_each(G, Gr)) < TOL


def test_long_record_late_time_phases(frf):
    """Phases up to ~1e9 rad (time origin far from zero): the Cody-Waite path must hold."""
    from b200id import synth
    n, nd = 100_000, 50
    t, u, y = synth.make_record(n, nd, "multisine", seed=2)
    t = t + 1.9e6                       # as if this were the tail of a 1e9-sample record
    _, w = ofrf.freq_grid(nd)
    U, Y = frf.demod_pair(t, u, y, w)
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e = record("frf.oracle.late_time.rel_each", max(rel_each(U, Ur), rel_each(Y, Yr)))
    assert e < TOL


def test_linearity_and_segment_additivity_full_size(frf):
    """Size-independent properties at the BASELINE cfg-2 shape (1e7 samples, nd=100):
    projecting [0,n) equals the sum of the projections of two owned segments (what the time-segment
    shards of the multi-GPU path compute), and the projection is linear in the signal."""
    import torch
    from b200id.runtime import require_cuda
    dev = require_cuda()
    n, nd = 10_000_000, 100
    _, w = ofrf.freq_grid(nd)
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    t = torch.arange(n, dtype=torch.float64, device=dev) * 0.002
    u = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    y = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    w_d = torch.from_numpy(w).to(dev)
    full = frf.project_device(t, u, y, w_d)
    cut = 3_333_333
    a = frf.project_device(t, u, y, w_d, own=(0, cut))
    b = frf.project_device(t, u, y, w_d, own=(cut, n))
    scale = float(full.abs().max())
    e = record("frf.property.segment_additivity", float((a + b - full).abs().max()) / scale)
    assert e < 1e-12
    lin = frf.project_device(t, 2.0 * u + y, y, w_d)
    expect = 2.0 * full[:2 * nd] + full[2 * nd:]
    e = record("frf.property.linearity", float((lin[:2 * nd] - expect).abs().max()) / scale)
    assert e < 1e-12
    again = frf.project_device(t, u, y, w_d)
    assert torch.equal(again, full), "projection must be deterministic run to run"


@pytest.mark.parametrize("n,nd,kind,t_shift", [(200_000, 100, "multisine", 0.0), (300_001, 150, "square", 0.0),
                                               (100_000, 50, "multisine", 1.9e6), (65_537, 150, "square", 123.4567),
                                               (5_000, 600, "multisine", 0.0)])
def test_uniform_fast_path_matches_oracle_and_general(frf, n, nd, kind, t_shift):
    """The uniform-timebase kernel (rotation recurrence + first-order rounding correction) against the
    oracle (per-bin 1e-9, including weakly excited bins of the broadband square-wave record) and against
    the general kernel; with time origins far from zero (phases ~ 1e9 rad) and off the sampling grid."""
    import torch
    from b200id import synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    t, u, y = synth.make_record(n, min(nd, 100), kind, seed=21)
    t = t + t_shift
    _, w = ofrf.freq_grid(nd)
    span = float(t[-1] - t[0])
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    kind_sel, t0, dt = frf.choose_path(t_d, span, float(w.max()), float(t[0]))
    assert kind_sel == "uniform"
    Uu, Yu = frf.record_coefficients_device(t_d, u_d, y_d, w_d, span, path=("uniform", t0, dt))
    Ug, Yg = frf.record_coefficients_device(t_d, u_d, y_d, w_d, span, path=("general", None, None))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e_o = max(rel_each(Uu.cpu().numpy(), Ur), rel_each(Yu.cpu().numpy(), Yr))
    e_g = max(rel_each(Uu.cpu().numpy(), Ug.cpu().numpy()), rel_each(Yu.cpu().numpy(), Yg.cpu().numpy()))
    record(f"frf.uniform.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_oracle", e_o)
    record(f"frf.uniform.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_general", e_g)
    assert e_o < TOL and e_g < TOL


def test_jittered_timebase_takes_general_path(frf, golden):
    import torch
    from b200id.runtime import require_cuda, to_device
    g = golden("frf")
    dev = require_cuda()
    tj = g["t_jit"]
    kind_sel, _, _ = frf.choose_path(to_device(tj, dev), float(tj[-1] - tj[0]), float(g["w"].max()), float(tj[0]))
    assert kind_sel == "general"


@pytest.mark.parametrize("n,nd", [(2, 1), (3, 5), (17, 3), (513, 513)])
def test_minimum_sizes(frf, n, nd):
    """Smallest legal records (n = 2) and bin counts around one CTA's bin group."""
    rng = np.random.default_rng(n * 1000 + nd)
    t = np.arange(n) * 0.002
    u, y = rng.standard_normal(n), rng.standard_normal(n)
    _, w = ofrf.freq_grid(nd)
    U, Y = frf.demod_pair(t, u, y, w)
    Ur, Yr = ofrf.demod_trapz(t, u, w), ofrf.demod_trapz(t, y, w)
    np.testing.assert_allclose(U, Ur, rtol=1e-9, atol=1e-13 * np.max(np.abs(Ur)))
    np.testing.assert_allclose(Y, Yr, rtol=1e-9, atol=1e-13 * np.max(np.abs(Yr)))


def test_dead_input_gives_nan_like_reference(frf):
    """u identically constant -> mean-removed input is zero -> denominator <= eps -> NaN+NaNj bins."""
    t = np.arange(1000) * 0.002
    u = np.full(1000, 3.0); y = np.sin(t)
    _, w = ofrf.freq_grid(10)
    U, Y = frf.demod_pair(t, u, y, w)
    G = frf.cross_power([U], [Y])
    Gr = ofrf.cross_power([ofrf.demod_trapz(t, u, w)], [ofrf.demod_trapz(t, y, w)])
    assert np.isnan(G).all() and np.isnan(Gr).all()


def test_all_kernels_ragged_sizes_run_clean():
    """tests/sanitize_smoke.py (every kernel at ragged / minimum sizes) must finish without a CUDA fault.
    (compute-sanitizer is closed on this GPU pool; this is the plain run it would wrap.)"""
    import subprocess, sys
    from pathlib import Path
    script = Path(__file__).resolve().parent / "sanitize_smoke.py"
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "all kernels ran" in out.stdout, out.stderr[-3000:]


def test_uniform_path_ill_conditioned_bins_use_rotation_kernel(frf):
    """Bins whose chain step lands on a multiple of pi (or is tiny) are unsafe for the three-term
    recurrence; they must be routed to the compact rotation kernel and still meet the 1e-9 parity."""
    from b200id import synth
    n = 50_000
    t, u, y = synth.make_record(n, 50, "square", seed=5)
    R, S, dt = 512 // 5, 2, 0.002
    w = np.array([1e-4, 1.0, np.pi / (R * S * dt), 2 * np.pi / (R * S * dt) + 1e-9, 300.0])
    U, Y = frf.demod_pair(t, u, y, w)
    Ur, Yr = ofrf.demod_trapz(t, u, w), ofrf.demod_trapz(t, y, w)
    e = record("frf.uniform.ill_conditioned_bins.rel_each", max(rel_each(U, Ur), rel_each(Y, Yr)))
    assert e < TOL


@pytest.mark.parametrize("n,chunks,kind,t_shift", [(300_001, 8, "multisine", 0.0), (200_003, 5, "square", 123.4567),
                                                   (65_537, 3, "multisine", 1.9e6), (1000, 8, "multisine", 0.0)])
def test_streamed_host_record_matches_oracle_and_resident(frf, n, chunks, kind, t_shift):
    """The copy-overlapped host path (signals first, time stamps in pieces, one projection per piece)
    against the oracle at the FRF tolerance and against the single-launch resident path."""
    from b200id import synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    nd = 60
    t, u, y = synth.make_record(n, nd, kind, seed=33)
    t = t + t_shift
    _, w = ofrf.freq_grid(nd)
    span, w_d = float(t[-1] - t[0]), to_device(w, dev)
    Us, Ys = frf.record_coefficients_streamed(t, u, y, w_d, span, w_max=float(w.max()), n_chunks=chunks)
    Ud, Yd = frf.record_coefficients_device(to_device(t, dev), to_device(u, dev), to_device(y, dev), w_d, span,
                                            t0=float(t[0]), w_max=float(w.max()))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e_o = max(rel_each(Us.cpu().numpy(), Ur), rel_each(Ys.cpu().numpy(), Yr))
    e_d = max(rel_each(Us.cpu().numpy(), Ud.cpu().numpy()), rel_each(Ys.cpu().numpy(), Yd.cpu().numpy()))
    record(f"frf.streamed.{kind}.n{n}.c{chunks}.vs_oracle", e_o)
    record(f"frf.streamed.{kind}.n{n}.c{chunks}.vs_resident", e_d)
    assert e_o < TOL and e_d < TOL
    # one signal, no mean subtraction
    U1, none = frf.record_coefficients_streamed(t, u, None, w_d, span, subtract_mean=False, w_max=float(w.max()),
                                                n_chunks=chunks)
    assert none is None
    assert rel_each(U1.cpu().numpy(), ofrf.demod_trapz(t, u, w, subtract_mean=False, threads=8)) < TOL


def test_streamed_record_nonuniform_and_late_jitter(frf, golden):
    """(i) a jittered time base is projected piecewise by the general kernel; (ii) a record whose FIRST piece
    sits on the grid but a later one does not is detected on the device and re-projected (no silent use of
    the uniform kernel off its validity range)."""
    from b200id import synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    g = golden("frf")
    tj, uj, yj, w = g["t_jit"], g["u"], g["y"], g["w"]
    w_d = to_device(w, dev)
    U, Y = frf.record_coefficients_streamed(tj, uj, yj, w_d, float(tj[-1] - tj[0]), w_max=float(w.max()), n_chunks=4)
    assert max(rel_each(U.cpu().numpy(), g["U_jit"]), rel_each(Y.cpu().numpy(), g["Y_jit"])) < TOL
    n, nd = 120_000, 40
    t, u, y = synth.make_record(n, nd, "multisine", seed=5)
    rng = np.random.default_rng(0)
    t = t.copy()
    t[n // 2:-1] += rng.uniform(-2e-4, 2e-4, size=n - n // 2 - 1)      # second half off the grid, still monotone
    assert np.all(np.diff(t) > 0)
    _, w = ofrf.freq_grid(nd)
    w_d = to_device(w, dev)
    U, Y = frf.record_coefficients_streamed(t, u, y, w_d, float(t[-1] - t[0]), w_max=float(w.max()), n_chunks=8)
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e = record("frf.streamed.late_jitter.vs_oracle", max(rel_each(U.cpu().numpy(), Ur), rel_each(Y.cpu().numpy(), Yr)))
    assert e < TOL


def test_public_demod_streams_large_host_records(frf, monkeypatch):
    """demod_pair on a host record above the streaming threshold = the same call with streaming disabled."""
    from b200id import synth
    n, nd = 1_200_000, 30
    t, u, y = synth.make_record(n, nd, "multisine", seed=8)
    _, w = ofrf.freq_grid(nd)
    assert n >= frf.STREAM_MIN_SAMPLES
    Us, Ys = frf.demod_pair(t, u, y, w)
    monkeypatch.setattr(frf, "_STREAM", False)
    Ub, Yb = frf.demod_pair(t, u, y, w)
    e = record("frf.streamed.public.vs_bulk", max(rel_each(Us, Ub), rel_each(Ys, Yb)))
    assert e < 1e-11
    Ur = ofrf.demod_trapz(t, u, w, threads=16)
    assert rel_each(Us, Ur) < TOL


@pytest.mark.parametrize("nd,R", [(100, 1), (150, 3), (7, 2), (1000, 1)])
def test_cross_power_sums_and_gp_targets(frf, nd, R):
    """b200id_cross_power_sums + b200id_gp_targets (G, StandardScaler statistics, standardised targets in one
    kernel) against NumPy: frf_estimator.py:239-252 and gp_pipeline.py:130-132."""
    import torch
    from b200id.runtime import require_cuda, empty
    dev = require_cuda()
    rng = np.random.default_rng(nd + R)
    U = rng.standard_normal((R, nd)) + 1j * rng.standard_normal((R, nd))
    Y = rng.standard_normal((R, nd)) + 1j * rng.standard_normal((R, nd))
    if nd > 5:
        U[:, 3] = 0.0                                        # dead bin -> NaN + NaN j
    sums = empty(3 * nd, dev)
    Ud, Yd = torch.from_numpy(U).to(dev), torch.from_numpy(Y).to(dev)
    frf.cross_power_sums_device(Ud[:1], Yd[:1], sums)
    if R > 1:
        frf.cross_power_sums_device(Ud[1:], Yd[1:], sums, accumulate=True)
    num, den = np.sum(Y * np.conj(U), axis=0), np.sum(np.abs(U) ** 2, axis=0)
    got = sums.cpu().numpy()
    np.testing.assert_allclose(got[:nd], num.real, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(got[nd:2 * nd], num.imag, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(got[2 * nd:], den, rtol=1e-13, atol=1e-300)
    with np.errstate(all="ignore"):
        Gref = np.where(den > 1e-15, num / den, np.nan + 1j * np.nan)
    G, stats, y_re, y_im = frf.gp_targets_device(sums, nd, standardise=False)
    assert stats is None
    np.testing.assert_allclose(G.cpu().numpy(), Gref, rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(y_re.cpu().numpy(), Gref.real, rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(y_im.cpu().numpy(), Gref.imag, rtol=1e-13, equal_nan=True)
    if nd > 5:
        sums[2 * nd + 3] = 1.0                               # revive the dead bin so that the statistics are finite
        Gref[3] = num[3]
    G, stats, y_re, y_im = frf.gp_targets_device(sums, nd, standardise=True)
    st = stats.cpu().numpy()
    ref = np.array([Gref.real.mean(), Gref.imag.mean(), Gref.real.std(), Gref.imag.std()])
    np.testing.assert_allclose(st, ref, rtol=1e-12)
    np.testing.assert_allclose(y_re.cpu().numpy(), (Gref.real - ref[0]) / ref[2], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(y_im.cpu().numpy(), (Gref.imag - ref[1]) / ref[3], rtol=1e-10, atol=1e-12)
    # zero variance -> scale 1 (sklearn)
    sums[:] = torch.cat([torch.full((nd,), 2.0), torch.full((nd,), -3.0), torch.ones(nd)]).to(dev)
    _, stats, y_re, _ = frf.gp_targets_device(sums, nd, standardise=True)
    assert stats.cpu().tolist() == [2.0, -3.0, 1.0, 1.0] and float(y_re.abs().max()) == 0.0


BUDGET = 4e-10          # DA_BUDGET of csrc/frf_adaptive.cuh


@pytest.mark.parametrize("n,nd,kind,t_shift", [(200_000, 100, "multisine", 0.0), (300_001, 150, "square", 0.0),
                                               (100_000, 50, "multisine", 1.9e6), (65_537, 150, "square", 123.4567),
                                               (4_097, 7, "multisine", 0.0), (70_000, 160, "square", -33.3)])
def test_matrix_product_path_matches_recurrence_and_oracle(frf, n, nd, kind, t_shift):
    """The experimental block-phasor evaluation (FP64 mma.sync main term + fixed-point rounding correction,
    b200id_frf_project_uniform_mma) against the recurrence kernels and the oracle: weak bins of the square-wave
    record, phases ~1e9 rad, ragged tiles, nd not a multiple of 4, the largest supported nd, negative times."""
    from b200id import synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    t, u, y = synth.make_record(n, min(nd, 100), kind, seed=5)
    t = t + t_shift
    _, w = ofrf.freq_grid(nd)
    span = float(t[-1] - t[0])
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    a = frf.project_uniform_mma_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
    b = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
    Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
    Ub, Yb = (x.cpu().numpy() for x in frf.finalize_device(b, nd, 2.0 / span))
    e_r = record(f"frf.mma.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_recurrence", max(rel_each(Ua, Ub), rel_each(Ya, Yb)))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e_o = record(f"frf.mma.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr)))
    assert e_r < TOL and e_o < TOL
    # owned sub-ranges add up (the streamed upload and the time-segment shards use them)
    cut = n // 3 + 5
    a1 = frf.project_uniform_mma_device(t_d, u_d, y_d, w_d, t0, dt, (0, cut), sums, 1.0 / n)
    a2 = frf.project_uniform_mma_device(t_d, u_d, y_d, w_d, t0, dt, (cut, n), sums, 1.0 / n)
    scale = float(a.abs().max())
    assert record(f"frf.mma.{kind}.n{n}.segments", float((a1 + a2 - a).abs().max()) / scale) < 1e-12
    again = frf.project_uniform_mma_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
    import torch
    assert torch.equal(again, a), "the matrix-product path must be deterministic run to run"


@pytest.mark.parametrize("n,nd,kind,t_shift", [(200_000, 100, "multisine", 0.0), (300_001, 150, "square", 0.0),
                                               (100_000, 50, "multisine", 1.9e6), (65_537, 150, "square", 123.4567),
                                               (4_097, 7, "multisine", 0.0), (70_000, 160, "square", -33.3),
                                               (1_000_003, 100, "square", 0.0), (50_000, 601, "multisine", 0.0)])
def test_goertzel_form_matches_three_term_and_oracle(frf, n, nd, kind, t_shift):
    """The Goertzel form of the uniform path (recurrence on the accumulators, rounding term in packed FP32,
    frf_project_uniform_g_kernel) against the three-term kernel and the oracle, on the records that stress the
    rounding term: weak bins of the square-wave record, phases ~1e9 rad, ragged tiles, odd bin counts, more than
    one bin group, negative times."""
    import ctypes as C
    import torch
    from b200id import _lib, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    t, u, y = synth.make_record(n, min(nd, 100), kind, seed=5)
    t = t + t_shift
    _, w = ofrf.freq_grid(nd)
    span = float(t[-1] - t[0])
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    lib = _lib.load()
    prev = lib.b200id_frf_select_recurrence(C.c_int(1))
    try:
        a = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        cut = n // 3 + 5
        a1 = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, cut), sums, 1.0 / n)
        a2 = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (cut, n), sums, 1.0 / n)
        again = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        only_u = frf.project_uniform_device(t_d, u_d, None, w_d, t0, dt, (0, n), sums, 1.0 / n)
        lib.b200id_frf_select_recurrence(C.c_int(0))
        b = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
    finally:
        lib.b200id_frf_select_recurrence(C.c_int(prev))
    Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
    Ub, Yb = (x.cpu().numpy() for x in frf.finalize_device(b, nd, 2.0 / span))
    e_r = record(f"frf.goertzel.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_three_term", max(rel_each(Ua, Ub), rel_each(Ya, Yb)))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e_o = record(f"frf.goertzel.{kind}.n{n}.nd{nd}.shift{t_shift:g}.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr)))
    assert e_r < TOL and e_o < TOL
    scale = float(a.abs().max())
    assert record(f"frf.goertzel.{kind}.n{n}.segments", float((a1 + a2 - a).abs().max()) / scale) < 1e-12
    assert torch.equal(again, a), "deterministic run to run"
    assert torch.equal(only_u[:2 * nd], a[:2 * nd]), "u sums do not depend on whether y rides along"


@pytest.mark.parametrize("n,nd,kind,t_shift", [(200_000, 100, "multisine", 0.0), (300_001, 150, "square", 0.0),
                                               (100_000, 50, "multisine", 1.9e6), (65_537, 150, "square", 123.4567),
                                               (4_097, 7, "multisine", 0.0), (70_000, 160, "square", -33.3),
                                               (1_000_003, 100, "square", 0.0), (50_000, 601, "multisine", 0.0),
                                               (1_537, 37, "square", 0.0), (1_000, 1, "multisine", 5.0)])
def test_adaptive_form_matches_goertzel_and_oracle(frf, n, nd, kind, t_shift):
    """The default evaluation of the uniform path (frf_adaptive.cuh: exact-phase main term of every bin as block-phasor
    FP64 matrix products on the FP64 tensor pipe; the reference's per-sample phase-rounding term as a packed-FP32
    Goertzel recurrence ONLY for the bins a device-side test finds weakly excited) against the all-recurrence Goertzel
    kernel and the oracle, on the records that stress it: weak bins of the square-wave record, phases ~1e9 rad, ragged
    tiles, bin counts that are not a multiple of 4, several bin groups, a single bin, negative times.  Run twice: with
    the per-bin decision and with every bin listed (the decision may only ever skip what does not matter); run-to-run
    determinism; sub-range calls (form 1) still add up to the whole-record call."""
    import ctypes as C
    import torch
    from b200id import _lib, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    t, u, y = synth.make_record(n, min(nd, 100), kind, seed=5)
    t = t + t_shift
    _, w = ofrf.freq_grid(nd)
    span = float(t[-1] - t[0])
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    lib = _lib.load()
    prev = lib.b200id_frf_select_recurrence(C.c_int(2))
    prev_force = lib.b200id_frf_adaptive_force(C.c_int(0))
    try:
        a = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        listed, worst_skipped = frf.adaptive_diag(nd, dev)
        cut = n // 3 + 5
        a1 = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, cut), sums, 1.0 / n)
        a2 = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (cut, n), sums, 1.0 / n)
        again = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        lib.b200id_frf_adaptive_force(C.c_int(1))
        full = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        listed_full, _ = frf.adaptive_diag(nd, dev)
        lib.b200id_frf_select_recurrence(C.c_int(1))
        b = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
    finally:
        lib.b200id_frf_select_recurrence(C.c_int(prev))
        lib.b200id_frf_adaptive_force(C.c_int(prev_force))
    tag = f"frf.adaptive.{kind}.n{n}.nd{nd}.shift{t_shift:g}"
    record(f"{tag}.bins_with_rounding_term", listed)
    record(f"{tag}.worst_skipped_bound_over_budget", worst_skipped / BUDGET)
    assert listed_full == nd and 0 <= listed <= nd and worst_skipped <= BUDGET
    Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
    Uf, Yf = (x.cpu().numpy() for x in frf.finalize_device(full, nd, 2.0 / span))
    Ub, Yb = (x.cpu().numpy() for x in frf.finalize_device(b, nd, 2.0 / span))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e_r = record(f"{tag}.vs_goertzel", max(rel_each(Ua, Ub), rel_each(Ya, Yb)))
    e_o = record(f"{tag}.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr)))
    e_f = record(f"{tag}.all_listed.vs_oracle", max(rel_each(Uf, Ur), rel_each(Yf, Yr)))
    # what skipping cost: the adaptive result against the same kernels with every bin listed
    record(f"{tag}.skipped_vs_all_listed", max(rel_each(Ua, Uf), rel_each(Ya, Yf)))
    assert e_r < TOL and e_o < TOL and e_f < TOL
    # sub-range calls run form 1 (every bin with its rounding term): against the whole-record call the sum differs by
    # what the decision skipped, i.e. by less than the budget relative to the largest sum
    scale = float(a.abs().max())
    assert record(f"{tag}.segments", float((a1 + a2 - a).abs().max()) / scale) < BUDGET
    assert torch.equal(again, a), "deterministic run to run"


def test_adaptive_form_lists_by_excitation(frf):
    """What the device-side decision does on the two kinds of record of BASELINE config 2: on a multisine record
    whose lines sit on the bins only the weak output bins need the rounding term; on the square-wave record the weakly
    excited input bins do as well.  A real timebase deviation (jitter far above the rounding level of t, here 5e-11 s against
    ulp(t) ~ 1e-13 s) is never treated as noise: its first-order bound w_k sum |a_i delta_i| lists every bin but the few
    lowest-frequency ones."""
    import ctypes as C
    from b200id import _lib, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    lib = _lib.load()
    n = 500_000
    out = {}
    prev = lib.b200id_frf_select_recurrence(C.c_int(2))
    prev_force = lib.b200id_frf_adaptive_force(C.c_int(0))
    try:
        for kind, nd, jitter in (("multisine", 100, 0.0), ("square", 150, 0.0), ("multisine", 100, 5e-11)):
            t, u, y = synth.make_record(n, 100, kind, seed=11)
            if jitter:
                t = t + jitter * np.random.default_rng(3).standard_normal(n)
            _, w = ofrf.freq_grid(nd)
            span = float(t[-1] - t[0])
            t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
            t0, dt = float(t[0]), span / (n - 1)
            sums = frf.moments_device(u_d, y_d, (0, n))
            a = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
            listed, _ = frf.adaptive_diag(nd, dev)
            out[(kind, jitter)] = listed
            Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
            Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
            e = record(f"frf.adaptive.listing.{kind}.jitter{jitter:g}.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr)))
            lib.b200id_frf_select_recurrence(C.c_int(1))
            b = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
            lib.b200id_frf_select_recurrence(C.c_int(2))
            Ub, Yb = (x.cpu().numpy() for x in frf.finalize_device(b, nd, 2.0 / span))
            record(f"frf.adaptive.listing.{kind}.jitter{jitter:g}.form1_vs_oracle", max(rel_each(Ub, Ur), rel_each(Yb, Yr)))
            record(f"frf.adaptive.listing.{kind}.jitter{jitter:g}.bins_with_rounding_term", listed)
            assert e < TOL
    finally:
        lib.b200id_frf_select_recurrence(C.c_int(prev))
        lib.b200id_frf_adaptive_force(C.c_int(prev_force))
    # the multisine's lines are all excited in u; what is listed are the OUTPUT's bins above the plant's corner
    # (|Y_k| rolls off as 1/f^4 there), about a third of the grid; the square wave adds its own weak input bins
    assert 0 < out[("multisine", 0.0)] <= 45
    assert out[("multisine", 0.0)] < out[("square", 0.0)] < 150
    assert out[("multisine", 5e-11)] >= 80


def test_adaptive_form_ill_conditioned_chain_step(frf):
    """Listed bins whose Goertzel chain step is ill-conditioned (|sin(w R dt)| < 1e-3) take the FP32 rotation inside
    the rounding-term kernel.  With every one of 50 bins listed its threads would walk R = 20 samples per step and the
    kernel lowers R (up to five times) to get away from such bins; bins planted at w = m pi / (R dt) for EVERY R it
    may try (20 ... 15) leave it no way out, so the rotation path runs; neighbours just inside and outside the
    threshold ride along."""
    import ctypes as C
    from b200id import _lib, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    n, nd = 120_000, 50
    t, u, y = synth.make_record(n, 100, "square", seed=9)
    _, w = ofrf.freq_grid(nd)
    w = w.copy()
    dt_nom = 0.002
    for i, R in enumerate(range(20, 14, -1)):
        w[10 + i] = (1 + i % 3) * np.pi / (R * dt_nom)
    w[30] = 3.0 * np.pi / (15 * dt_nom) * (1.0 - 2e-4)
    w[49] = 2.0 * np.pi / (15 * dt_nom) * (1.0 + 1e-5)
    span = float(t[-1] - t[0])
    t_d, u_d, y_d, w_d = (to_device(a, dev) for a in (t, u, y, w))
    t0, dt = float(t[0]), span / (n - 1)
    sums = frf.moments_device(u_d, y_d, (0, n))
    lib = _lib.load()
    prev = lib.b200id_frf_select_recurrence(C.c_int(2))
    prev_force = lib.b200id_frf_adaptive_force(C.c_int(1))
    try:
        a = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        listed, _ = frf.adaptive_diag(nd, dev)
    finally:
        lib.b200id_frf_select_recurrence(C.c_int(prev))
        lib.b200id_frf_adaptive_force(C.c_int(prev_force))
    assert listed == nd
    Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    e = record("frf.adaptive.ill_conditioned_step.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr)))
    assert e < TOL


def test_adaptive_form_unaligned_arrays(frf):
    """Views that start 8 bytes into an allocation are not 16-byte aligned: the TMA-staged main-term kernel does not
    apply and the plain build of the same arithmetic runs (frf_main_dmma_kernel); same result as on aligned copies up
    to the order of the tile sums (and, for a bin at the edge of the decision, the skipped rounding term), and within
    tolerance of the oracle."""
    import ctypes as C
    import torch
    from b200id import _lib, synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    n, nd = 150_001, 60
    t, u, y = synth.make_record(n + 1, 60, "multisine", seed=13)
    _, w = ofrf.freq_grid(nd)
    big = [to_device(a, dev) for a in (t, u, y)]
    t_d, u_d, y_d = (b[1:] for b in big)                                  # data_ptr() % 16 == 8
    assert all(v.data_ptr() % 16 == 8 for v in (t_d, u_d, y_d))
    w_d = to_device(w, dev)
    t1, u1, y1 = t[1:], u[1:], y[1:]
    span = float(t1[-1] - t1[0])
    t0, dt = float(t1[0]), span / (n - 1)
    lib = _lib.load()
    prev = lib.b200id_frf_select_recurrence(C.c_int(2))
    try:
        sums = frf.moments_device(u_d, y_d, (0, n))
        a = frf.project_uniform_device(t_d, u_d, y_d, w_d, t0, dt, (0, n), sums, 1.0 / n)
        tc, uc, yc = (v.clone() for v in (t_d, u_d, y_d))                 # aligned copies -> warp-specialised kernel
        b = frf.project_uniform_device(tc, uc, yc, w_d, t0, dt, (0, n), sums, 1.0 / n)
    finally:
        lib.b200id_frf_select_recurrence(C.c_int(prev))
    Ua, Ya = (x.cpu().numpy() for x in frf.finalize_device(a, nd, 2.0 / span))
    Ub, Yb = (x.cpu().numpy() for x in frf.finalize_device(b, nd, 2.0 / span))
    Ur, Yr = ofrf.demod_trapz(t1, u1, w, threads=8), ofrf.demod_trapz(t1, y1, w, threads=8)
    assert record("frf.adaptive.unaligned.vs_oracle", max(rel_each(Ua, Ur), rel_each(Ya, Yr))) < TOL
    assert record("frf.adaptive.unaligned.vs_aligned", max(rel_each(Ua, Ub), rel_each(Ya, Yb))) < BUDGET
