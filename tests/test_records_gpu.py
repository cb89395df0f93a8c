"""Resident-record reuse on the GPU: the reference's own call sequences on
one file -- FRF of u then of y (data_loader.py:170-185), paper-mode FIR validation (fir_fitting.py:267-306),
Wave.mat validation (fir_validation.py:163-190) -- cost ONE parse and ONE upload, and return exactly what the
un-cached calls return."""
import contextlib
import io

import numpy as np
import pytest

from oracle import frf as ofrf, fir as ofir
from parity_util import rel_each, rel_max, record

pytestmark = pytest.mark.gpu


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture()
def store():
    from b200id import records
    records.STORE.clear()
    yield records.STORE
    records.STORE.clear()


@pytest.mark.parametrize("n", [6000, 1_100_000])          # bulk upload / streamed first pass that is then adopted
def test_reference_call_sequence_runs_on_one_upload(tmp_path, store, n):
    from b200id import synth, _lib
    from src.frequency_transform.frf_estimator import (load_time_u_y, synchronous_coefficients_trapz,
                                                       frf_cross_power_average, matlab_freq_grid)
    from src.frequency_transform import estimate_frequency_response
    from src.fir_model.fir_validation import validate_fir_with_mat
    nd = 40
    t, u, y = synth.make_record(n, nd, "multisine", seed=12)
    p = synth.write_mat(tmp_path / "val.mat", t, u, y)
    _, w = matlab_freq_grid(nd, -1.0, 2.3)
    # data_loader.py:170-188 verbatim in structure
    tt, uu, yy = load_time_u_y(p, y_col=0)
    U = synchronous_coefficients_trapz(tt, uu, w, 0.0, True)
    launches_after_u = _lib.launch_count
    Y = synchronous_coefficients_trapz(tt, yy, w, 0.0, True)
    assert _lib.launch_count == launches_after_u, "the y coefficients come from the same fused pass"
    G = frf_cross_power_average([U], [Y])
    Ur, Yr = ofrf.demod_trapz(t, u, w, threads=8), ofrf.demod_trapz(t, y, w, threads=8)
    assert record(f"records.frf.n{n}.vs_oracle", max(rel_each(U, Ur), rel_each(Y, Yr))) < 1e-9
    assert store.stats["uploads"] == 1 and store.stats["mat_misses"] == 1
    # results handed out are private copies
    U[0] = 0
    assert synchronous_coefficients_trapz(tt, uu, w, 0.0, True)[0] != 0
    # the file-level entry on the same file: no parse, no upload, no projection
    before = _lib.launch_count
    df = estimate_frequency_response([str(p)], method="frf", nd=nd)
    assert store.stats["uploads"] == 1 and store.stats["mat_misses"] == 1
    assert _lib.launch_count - before == 1, "only the cross-power kernel runs"
    assert np.allclose(df["ReG"].to_numpy() + 1j * df["ImG"].to_numpy(), G, rtol=1e-13, atol=0)
    # another grid / a transient drop: same resident copy, new projection
    _, w2 = matlab_freq_grid(25, -1.0, 2.0)
    U2 = synchronous_coefficients_trapz(tt, uu, w2, 1.0, True)
    assert rel_each(U2, ofrf.demod_trapz(t, u, w2, drop_seconds=1.0, threads=8)) < 1e-9
    assert store.stats["uploads"] == 1
    # FIR validation of the same file: u, y already in HBM
    g = np.random.default_rng(0).standard_normal(300) / 300
    v = quiet(validate_fir_with_mat, g, p, output_dir=None)
    mr, _ = ofir.fir_validate(u, y, g)
    for k in ("rmse", "fit_percent", "r2"):
        assert abs(v[k] - mr[k]) <= 1e-11 * max(1.0, abs(mr[k])), k
    assert store.stats["mat_misses"] == 1
    assert store.stats["uploads"] == 1          # both loaders pull the same vectors: one shared record, one HBM copy


def test_cached_and_uncached_paths_agree(tmp_path, store, monkeypatch):
    from b200id import synth, records
    from src.frequency_transform import estimate_frequency_response
    from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline
    n, nd = 50_000, 30
    t, u, y = synth.make_record(n, nd, "multisine", seed=2)
    p = synth.write_mat(tmp_path / "rec.mat", t, u, y)
    df1 = estimate_frequency_response([str(p)], method="frf", nd=nd, drop_seconds=0.5)
    r1 = quiet(gp_to_fir_direct_pipeline, df1["omega_rad_s"].to_numpy(), df1["ReG"].to_numpy() + 1j * df1["ImG"].to_numpy(),
               gp_predict_func=synth.plant_response, mat_file=p, output_dir=tmp_path / "o1")
    monkeypatch.setattr(records, "ENABLED", False)
    df2 = estimate_frequency_response([str(p)], method="frf", nd=nd, drop_seconds=0.5)
    r2 = quiet(gp_to_fir_direct_pipeline, df2["omega_rad_s"].to_numpy(), df2["ReG"].to_numpy() + 1j * df2["ImG"].to_numpy(),
               gp_predict_func=synth.plant_response, mat_file=p, output_dir=tmp_path / "o2")
    assert rel_each(df1.to_numpy()[:, 2:], df2.to_numpy()[:, 2:]) < 1e-12
    for k in ("rmse", "fit_percent", "r2", "Ts"):
        assert abs(r1[k] - r2[k]) <= 1e-12 * max(1.0, abs(r2[k])), k
    assert np.array_equal(r1["fir_coefficients"], r2["fir_coefficients"])


def test_non_monotone_record_with_drop_takes_host_path(tmp_path, store):
    """A stored record whose time stamps are not increasing cannot drop a prefix by index: host path."""
    from scipy.io import savemat
    from src.frequency_transform.frf_estimator import load_time_u_y, synchronous_coefficients_trapz, matlab_freq_grid
    rng = np.random.default_rng(1)
    n = 4000
    t = np.arange(n) * 0.002
    perm = np.arange(n); perm[100], perm[200] = 200, 100
    t = t[perm]
    u, y = rng.standard_normal(n), rng.standard_normal(n)
    savemat(str(tmp_path / "nm.mat"), {"output": np.vstack([t, y, u])})
    tt, uu, yy = load_time_u_y(tmp_path / "nm.mat")
    _, w = matlab_freq_grid(10, -1.0, 2.0)
    got = synchronous_coefficients_trapz(tt, uu, w, 0.5, True)
    assert rel_each(got, ofrf.demod_trapz(t, u, w, drop_seconds=0.5)) < 1e-9
