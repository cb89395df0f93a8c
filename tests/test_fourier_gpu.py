"""GPU parity of the Fourier (FFT-ratio) estimator behind ``estimate_frequency_response(method='fourier')``
 against the reference's golden vectors and the oracle.

Tolerance: the reference's FFT carries an absolute error of ~eps * ||x|| * sqrt(log n) in EVERY bin, so a bin is
only defined to that absolute level; the comparisons below are relative to the largest spectrum magnitude for the
spectra and 1e-9 relative to the largest |G| for the interpolated transfer function."""
import numpy as np
import pytest

from oracle import fourier as ofou
from parity_util import record

pytestmark = pytest.mark.gpu


def _rel_to_max(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def _point_tolerance(f, U, Y, grid_hz, G_ref, base=1e-9):
    """Per grid point: 1e-9, widened by the conditioning of Y/U at the two bracketing bins (an FFT bin is only
    defined to ~1e-13 of the spectral peak, in the reference as much as here)."""
    j = np.clip(np.searchsorted(f, grid_hz, side="right") - 1, 0, f.size - 2)
    weakest = np.minimum(np.abs(U[j]), np.abs(U[j + 1]))
    strongest_y = np.maximum(np.abs(Y[j]), np.abs(Y[j + 1]))
    cond = 1e-12 * (np.abs(U).max() + np.abs(Y).max()) * (1.0 + strongest_y / weakest) / weakest
    with np.errstate(all="ignore"):
        return np.maximum(base, cond / np.maximum(np.abs(G_ref), 1e-300))


@pytest.fixture(scope="module")
def fou():
    from b200id import fourier
    return fourier


@pytest.mark.parametrize("tag,n", [("even", 6000), ("odd", 4999)])
def test_selected_bins_match_reference_fft(fou, golden, tag, n):
    from b200id import synth
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    g = golden("fourier")
    t, u, y = synth.make_record(n, 50, "multisine", seed=11)
    dt = float(np.mean(np.diff(t)))
    f_ref = g[f"f_{tag}"]
    assert fou.positive_bin_count(n) == f_ref.size and fou.top_frequency(n, dt) == f_ref[-1]
    bins = np.unique(np.r_[0, 1, 2, 17, 500, 501, f_ref.size - 2, f_ref.size - 1, np.arange(40, 400, 7)]).astype(np.int64)
    assert np.array_equal(bins * (1.0 / (n * dt)), f_ref[bins])
    U, Y = fou.spectra_at_bins_device(to_device(u, dev), to_device(y, dev), bins, dt, "hann")
    e = record(f"fourier.U_bins.{tag}", _rel_to_max(U, g[f"U_{tag}_hann"][bins]))
    assert e < 1e-13
    for wname in (("hann", "none", "blackman") if tag == "even" else ("hann",)):
        win = None if wname == "none" else wname
        U, Y = fou.spectra_at_bins_device(to_device(u, dev), to_device(y, dev), bins, dt, win)
        G = Y / (U + 1e-10)
        G[np.abs(U) < 1e-9] = 0.0
        ref = g[f"G_{tag}_{wname}"][bins]
        strong = np.abs(ofou.spectrum(t, u, win)[1][bins]) > 1e-6 * np.max(np.abs(ofou.spectrum(t, u, win)[1]))
        e = record(f"fourier.G_bins.{tag}.{wname}", float(np.max(np.abs(G[strong] - ref[strong]) / np.abs(ref[strong]))))
        assert e < 1e-8, wname                 # bins carrying signal: ratio of two spectra each good to ~1e-13 of the peak


def test_bracket_bins_cover_np_interp(fou):
    rng = np.random.default_rng(0)
    for n, dt in ((6000, 0.002), (4999, 0.0020000000001), (12, 0.1), (1_000_003, 0.002)):
        n_pos = fou.positive_bin_count(n)
        f = np.arange(n_pos) * (1.0 / (n * dt))
        grid = np.r_[np.linspace(0.0, f[-1] * 1.2, 57), f[[0, n_pos // 2, -1]], rng.uniform(0, f[-1], 40)]
        bins = fou.bracket_bins(grid, n, dt)
        vals = rng.standard_normal(n_pos)
        assert np.array_equal(np.interp(grid, f, vals), np.interp(grid, f[bins], vals[bins])), (n, dt)
        assert bins.size <= 2 * grid.size + 2


def test_estimate_frequency_response_fourier_golden(tmp_path, golden):
    from b200id import synth
    from src.frequency_transform import estimate_frequency_response
    g = golden("fourier")
    t, u, y = synth.make_record(6000, 50, "multisine", seed=0)
    t2, u2, y2 = synth.make_record(5555, 50, "multisine", seed=3)
    p1 = synth.write_mat(tmp_path / "a.mat", t, u, y)
    p2 = synth.write_mat(tmp_path / "b.mat", t2, u2, y2)
    cases = {"df_2files": dict(mat_files=[str(p1), str(p2)], nd=40, f_low=0.1, f_up=100.0, n_files=2),
             "df_window": dict(mat_files=[str(p1)], nd=25, f_low=0.5, f_up=60.0, time_duration=7.0, drop_seconds=1.0, n_files=1, window="none"),
             "df_top": dict(mat_files=[str(p1)], nd=30, f_low=0.0, f_up=1e9, n_files=1),
             "df_blackman": dict(mat_files=[str(p2)], nd=17, f_low=1.0, f_up=33.0, n_files=1, window="blackman")}
    recs = {"a": (t, u, y), "b": (t2, u2, y2)}
    for key, kw in cases.items():
        df = estimate_frequency_response(method="fourier", **kw)
        got, ref = df.to_numpy(), g[key]
        assert list(df.columns) == ["omega_rad_s", "freq_Hz", "ReG", "ImG", "absG", "phase_rad"]
        assert np.array_equal(got[:, :2], ref[:, :2]), key                   # grid bits identical
        Gg, Gr = got[:, 2] + 1j * got[:, 3], ref[:, 2] + 1j * ref[:, 3]
        # G = Y / U at a bin is defined to (FFT rounding ~ 1e-13 of the spectral peak) / |U_bin|: grid points whose
        # bracketing bins hold only leakage (between the multisine lines, above its band) are ill-conditioned
        # in the reference too.  Allow 1e-9 plus that conditioning term.
        tol = np.full(len(Gr), 1e-9)
        for name in ("a", "b")[:len(kw["mat_files"])] if len(kw["mat_files"]) == 2 else ("a" if "a.mat" in kw["mat_files"][0] else "b",):
            tt, uu, yy = ofou.apply_time_window(*recs[name], kw.get("drop_seconds", 0.0), kw.get("time_duration"))
            win = None if kw.get("window", "hann") == "none" else kw.get("window", "hann")
            f, U = ofou.spectrum(tt, uu, win)
            _, Y = ofou.spectrum(tt, yy, win)
            tol = np.maximum(tol, _point_tolerance(f, U, Y, ref[:, 1], Gr))
        err = np.abs(Gg - Gr) / np.maximum(np.abs(Gr), 1e-300)
        record(f"fourier.dropin.{key}.rel_each_max", float(err.max()))
        record(f"fourier.dropin.{key}.well_conditioned_points", int((tol <= 1e-9).sum()))
        assert np.all(err <= tol), (key, float((err / tol).max()))
        if key in ("df_2files", "df_window"):          # in-band grids: also tight relative to the largest |G|
            assert record(f"fourier.dropin.{key}.rel_to_max", _rel_to_max(Gg, Gr)) < 1e-9


def test_fourier_full_size_against_oracle(tmp_path):
    """2^20 + 1 samples (odd, not a power of two): the direct bins against the oracle's full FFT."""
    from b200id import synth, fourier as fou
    n, nd = (1 << 20) + 1, 64
    t, u, y = synth.make_record(n, 50, "multisine", seed=5)
    grid = np.linspace(0.2, 120.0, nd)
    G = fou.transfer_function_on_grid(t, u, y, grid, window="hann")
    f, Gref = ofou.transfer_function(t, u, y, "hann")
    ref = ofou.interpolate(f, Gref, grid)
    tol = _point_tolerance(f, ofou.spectrum(t, u, "hann")[1], ofou.spectrum(t, y, "hann")[1], grid, ref)
    err = np.abs(G - ref) / np.maximum(np.abs(ref), 1e-300)
    record("fourier.full_size.rel_each_max", float(err.max()))
    record("fourier.full_size.points_at_1e-9", int((tol <= 1e-9).sum()))
    assert np.all(err <= tol), float((err / tol).max())
