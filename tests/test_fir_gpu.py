"""GPU parity of K5/K6 (Hermitian IDFT -> FIR taps, FIR convolution + fused error sums) through the
C ABI, against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from oracle import fir as ofir
from parity_util import rel_each, rel_max, record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fir():
    from b200id import fir as mod
    return mod


def test_idft_golden(fir, golden):
    g = golden("fir")
    taps = fir.idft_taps(g["G"][:9], 11)
    assert record("fir.idft.small.rel_max", rel_max(taps, g["taps_small"])) < 1e-13
    taps = fir.idft_taps(g["G_uni"], 1024)
    assert record("fir.idft.paper.rel_max", rel_max(taps, g["taps_paper"])) < 1e-12
    full = fir.idft_taps(g["G_uni"], 1999)
    _, ref_full = ofir.idft_taps(ofir.hermitian_extend(g["G_uni"]), 1999)
    assert rel_max(full, ref_full) < 1e-12
    with pytest.raises(ValueError, match="exceeds available length"):
        fir.idft_taps(g["G"][:9], 18)


def test_convolution_and_metrics_golden(fir, golden):
    g = golden("fir")
    m, yp = fir.fir_validate(g["u"], g["y"], g["pipe_taps"])
    assert record("fir.conv.golden.y_pred.rel_max", rel_max(yp, g["y_pred"])) < 1e-13
    ref = g["validate_metrics"]
    got = np.array([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]])
    assert record("fir.metrics.golden.rel_each", rel_each(got, ref)) < 1e-12
    m, _ = fir.fir_validate(g["u"], g["y"], g["pipe_interp_taps"])
    got = np.array([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]])
    assert rel_each(got, g["pipe_interp_metrics"]) < 1e-12
    m, _ = fir.fir_validate(g["u"], g["y"], g["pipe_taps"][:64], detrend=True)
    got = np.array([m["rmse"], m["fit_percent"], m["r2"], m["transient_skip"]])
    assert rel_each(got, g["validate_detrend_metrics"]) < 1e-12


@pytest.mark.parametrize("n,taps", [(1, 1), (5, 3), (2047, 1024), (2048, 1024), (2049, 7), (100_003, 1024),
                                    (50_000, 300), (4096, 1999), (10_000, 2500),
                                    # overlap-save FFT path (n >= 32768, 128 <= taps <= 2048), block-edge cases
                                    (32_768, 128), (40_000, 2048), (3073 * 2 * 6, 1024), (3073 * 2 * 6 + 1, 1024),
                                    (250_001, 1025), (70_000, 127), (70_000, 2049)])
def test_convolution_oracle_ragged(fir, n, taps):
    rng = np.random.default_rng(n + taps)
    u = rng.standard_normal(n); g = rng.standard_normal(taps) / taps
    got = fir.fir_predict(u, g)
    ref = ofir.fir_predict(u, g)
    e = record(f"fir.conv.oracle.n{n}.taps{taps}.rel_max", rel_max(got, ref))
    assert e < 1e-13
    if n > 20:
        y = ref + 0.01 * rng.standard_normal(n)
        m, _ = fir.fir_validate(u, y, g)
        mr, _ = ofir.fir_validate(u, y, g)
        for k in ("rmse", "fit_percent", "r2"):
            assert abs(m[k] - mr[k]) <= 1e-11 * max(1.0, abs(mr[k])), k
        assert m["transient_skip"] == mr["transient_skip"]


def test_shard_history_matches_whole(fir):
    """A slice with `lead` samples of history reproduces the same outputs and additive sums as the
    whole record (what the FIR shards of the multi-GPU path compute); run at the cfg-2 size."""
    import torch
    from b200id.runtime import require_cuda
    dev = require_cuda()
    n, taps = 10_000_000, 1024
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    u = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    y = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
    g = torch.randn(taps, dtype=torch.float64, device=dev, generator=gen) / taps
    skip, y0 = 10, float(y[10].item())
    whole, pred = fir.fir_eval_device(u, y, g, skip, y0, want_pred=True)
    cut = 4_321_987
    a, pa = fir.fir_eval_device(u[:cut], y[:cut], g, skip, y0, want_pred=True)
    b, pb = fir.fir_eval_device(u[cut:], y[cut:], g, 0, y0, lead=taps - 1, want_pred=True)
    # the overlap-save path blocks the record from each slice's own origin: equal to rounding, not bitwise
    scale = float(pred.abs().max())
    assert float((pa - pred[:cut]).abs().max()) < 1e-13 * scale and float((pb - pred[cut:]).abs().max()) < 1e-13 * scale
    e = record("fir.property.shard_additivity", float(((a + b - whole).abs() / whole.abs()).max()))
    assert e < 1e-12
    # linearity in the taps, and idempotence (deterministic run to run)
    w2, p2 = fir.fir_eval_device(u, y, 2.0 * g, skip, y0, want_pred=True)
    assert float((p2 - 2.0 * pred).abs().max()) < 1e-12
    again, _ = fir.fir_eval_device(u, y, g, skip, y0)
    assert torch.equal(again, whole)
    # spot-check 64 outputs against a direct dot product
    idx = torch.randint(taps, n, (64,), generator=torch.Generator().manual_seed(0))
    for i in idx.tolist():
        ref = float((u[i - taps + 1:i + 1].flip(0) * g).sum().item())
        assert abs(float(pred[i].item()) - ref) < 1e-12


@pytest.mark.parametrize("n,taps,chunks", [(300_007, 1024, 6), (70_001, 300, 3), (150_000, 2048, 8), (41, 7, 4)])
def test_streamed_host_validation_matches_oracle(fir, n, taps, chunks):
    """Piecewise upload + per-piece FIR (history from the pieces already resident) = the oracle's whole-record
    convolution and metrics; both with and without the prediction coming back."""
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    rng = np.random.default_rng(n)
    u = rng.standard_normal(n); g = rng.standard_normal(taps) / taps
    ref = ofir.fir_predict(u, g)
    y = ref + 0.01 * rng.standard_normal(n)
    skip = min(10, taps // 10)
    sums, pred = fir.fir_eval_streamed(u, y, to_device(g, dev), skip, float(y[skip]), n, want_pred=True, n_chunks=chunks)
    assert rel_max(pred.cpu().numpy(), ref) < 1e-13
    m = fir.metrics_from_sums(sums.cpu().numpy(), skip)
    mr, _ = ofir.fir_validate(u, y, g)
    for k in ("rmse", "fit_percent", "r2"):
        assert abs(m[k] - mr[k]) <= 1e-11 * max(1.0, abs(mr[k])), k
    sums2, none = fir.fir_eval_streamed(u, y, to_device(g, dev), skip, float(y[skip]), n, want_pred=False, n_chunks=chunks)
    assert none is None
    np.testing.assert_allclose(sums2.cpu().numpy(), sums.cpu().numpy(), rtol=1e-13)


def test_public_validate_streams_large_host_records(fir, monkeypatch):
    n, taps = 1_100_000, 1024
    rng = np.random.default_rng(4)
    u = rng.standard_normal(n); g = rng.standard_normal(taps) / taps
    y = rng.standard_normal(n)
    assert n >= fir.STREAM_MIN_SAMPLES
    ms, ps = fir.fir_validate(u, y, g)
    monkeypatch.setattr(fir, "_STREAM", False)
    mb, pb = fir.fir_validate(u, y, g)
    assert rel_max(ps, pb) < 1e-13
    for k in ("rmse", "fit_percent", "r2"):
        assert abs(ms[k] - mb[k]) <= 1e-12 * max(1.0, abs(mb[k])), k


@pytest.mark.parametrize("n,taps,lead", [(100_003, 1024, 0), (65_536, 128, 0), (70_001, 2048, 0), (250_000, 1024, 1023),
                                         (40_000, 777, 300)])
def test_radix16_passes_reproduce_radix4_passes(fir, n, taps, lead):
    """The radix-16 organisation of the overlap-save transform (fir_fft16_kernel: two radix-4 stages per pass in
    registers) runs the radix-4 kernel's butterflies in the same order, so the predictions are the radix-4
    kernel's bit for bit -- ragged last unit, shortest / longest tap counts, history before the slice."""
    import ctypes as C
    import torch
    from b200id import _lib
    from b200id.runtime import require_cuda, to_device
    dev = require_cuda()
    rng = np.random.default_rng(n + taps)
    u_full, y_full, g = rng.standard_normal(n + lead), rng.standard_normal(n + lead), rng.standard_normal(taps) * 0.02
    u_d, y_d, g_d = to_device(u_full, dev), to_device(y_full, dev), to_device(g, dev)
    skip, y0 = 10, float(y_full[lead + 10])
    lib = _lib.load()
    out = {}
    prev = lib.b200id_fir_select_fft(C.c_int(0))
    try:
        for radix4 in (0, 1):
            lib.b200id_fir_select_fft(C.c_int(radix4))
            sums, pred = fir.fir_eval_device(u_d[lead:], y_d[lead:], g_d, skip, y0, want_pred=True, lead=lead)
            torch.cuda.synchronize()
            out[radix4] = (sums.cpu().numpy(), pred.cpu().numpy())
    finally:
        lib.b200id_fir_select_fft(C.c_int(-1 if prev == 0 else prev))
    assert np.array_equal(out[0][1], out[1][1])                      # predictions: bit for bit
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-13)     # sums: other thread -> element map, other addition order
    ref = np.convolve(u_full, g)[lead:lead + n]
    assert record(f"fir.fft16.n{n}.taps{taps}.lead{lead}", float(np.max(np.abs(out[0][1] - ref)) / np.max(np.abs(ref)))) < 1e-12
