"""The kernel-regularised FIR oracle (oracle/kr_fir.py) against vectors produced by the reference itself
(tests/golden/kr_fir.npz, oracle/make_golden.py::golden_kr)."""
import numpy as np
import pytest

from oracle import kr_fir as okr

KERNELS = ("dc", "ss", "si")


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("L", [24, 128])
def test_summaries_and_kernels(golden, L):
    g = golden("kr_fir")
    G, b, yy = okr.summaries(g["u"], g["y"], L)
    assert rel(G, g[f"G_{L}"]) < 1e-13 and rel(b, g[f"b_{L}"]) < 1e-13 and abs(yy - g[f"yy_{L}"][0]) <= 1e-13 * yy
    for k in KERNELS:
        th = g[f"theta_{k}_{L}"][1]
        assert rel(okr.build_kernel(L, k, th[:okr.KDIM[k]]), g[f"K_{k}_{L}"]) < 1e-13, k


@pytest.mark.parametrize("L", [24, 128])
@pytest.mark.parametrize("kernel", KERNELS)
def test_nll_and_posterior(golden, kernel, L):
    g = golden("kr_fir")
    G, b, yy, n = g[f"G_{L}"], g[f"b_{L}"], float(g[f"yy_{L}"][0]), len(g["u"])
    got = np.array([okr.nll(th, G, b, yy, n, L, kernel) for th in g[f"theta_{kernel}_{L}"]])
    ref = g[f"nll_{kernel}_{L}"]
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-9, (got, ref)
    g_hat, sig2, _ = okr.posterior_mean_g(g[f"theta_{kernel}_{L}"][0], G, b, L, kernel)
    assert rel(g_hat, g[f"ghat_{kernel}_{L}"]) < 1e-8 and abs(sig2 - g[f"sig2_{kernel}_{L}"][0]) <= 1e-15 * sig2
