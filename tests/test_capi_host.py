"""CPU-side checks of the C ABI: the library loads, exports every symbol include/b200id.h declares,
and the scalar device math (compiled for the host too) matches NumPy / the oracle.  No GPU calls."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from b200id import _lib
from oracle import gp as ogp

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200id_build", ROOT / "gauss-process_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "b200id.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200id_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200id.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.b200id_version() == 100


def test_kernel_table_matches_oracle():
    from b200id import gp
    for name, (kid, names, _) in ogp.KERNELS.items():
        assert gp.KERNEL_IDS[name] == (kid, len(names))


def test_sincos_host_large_arguments(lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 20000), rng.uniform(-1e5, 1e5, 20000),
                        rng.uniform(-3e9, 3e9, 40000), np.array([0.0, np.pi / 4, -np.pi / 4, 1e-300, 2.5e9])])
    s = np.empty_like(x); c = np.empty_like(x)
    lib.b200id_selftest_sincos_host(x.ctypes.data_as(C.c_void_p), C.c_int64(x.size),
                                    s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    # absolute error a few 1e-16: the FRF sums sample * cos/sin, so absolute accuracy is what matters
    assert np.max(np.abs(s - np.sin(x))) < 4e-16
    assert np.max(np.abs(c - np.cos(x))) < 4e-16


@pytest.mark.parametrize("name", [k for k in ogp.KERNELS if k != "matern_nu"])
def test_gram_host_matches_oracle(lib, golden, name):
    g = golden("gp")
    kid = ogp.KERNELS[name][0]
    for th in g[f"theta_{name}"]:
        for X1, X2 in ((g["X"][:24], g["X"][:24]), (g["Xv"][:10], g["X"][:24]), (g["Xi"], g["Xi"]), (g["Xj"], g["Xi"])):
            x1 = np.ascontiguousarray(X1.ravel()); x2 = np.ascontiguousarray(X2.ravel())
            th3 = np.zeros(3); th3[:len(th)] = th
            K = np.empty((x1.size, x2.size))
            lib.b200id_selftest_gram_host(C.c_int(kid), th3.ctypes.data_as(C.c_void_p), x1.ctypes.data_as(C.c_void_p),
                                          C.c_int(x1.size), x2.ctypes.data_as(C.c_void_p), C.c_int(x2.size),
                                          K.ctypes.data_as(C.c_void_p))
            with np.errstate(all="ignore"):
                ref = ogp.gram(name, th, X1, X2)
            # same operation order as NumPy; only libm exp/pow may differ in the last bit
            np.testing.assert_allclose(K, ref, rtol=1e-15, atol=0, equal_nan=True)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_general_nu_matern_host_matches_reference(lib, golden):
    """The device Bessel-K (csrc/gp_bessel.cuh: Temme series / Steed CF2 + recurrence), compiled for the host, against
    the reference's scipy.special.kv Gram matrices (golden matern_nu.npz, kernels.py:131-135): orders 0.3 ... 12.25,
    kv arguments from 1e-5 to beyond the underflow of exp(-x).  scipy's AMOS and this implementation are both good to
    a few 1e-14 in that range; the oracle restatement itself must be bit-identical to the reference."""
    g = golden("matern_nu")
    worst = 0.0
    for a, nu in enumerate(g["nus"]):
        for q, (ls, var) in enumerate(g["thetas"]):
            th = np.array([ls, var, nu])
            for key, X1, X2 in ((f"gram_{a}_{q}", g["X"][:24], g["X"][:24]), (f"gramx_{a}_{q}", g["Xv"][:10], g["X"][:24]),
                                (f"gramw_{a}_{q}", g["Xw"], g["Xw"])):
                x1 = np.ascontiguousarray(X1.ravel()); x2 = np.ascontiguousarray(X2.ravel())
                K = np.empty((x1.size, x2.size))
                lib.b200id_selftest_gram_host(C.c_int(13), th.ctypes.data_as(C.c_void_p), x1.ctypes.data_as(C.c_void_p),
                                              C.c_int(x1.size), x2.ctypes.data_as(C.c_void_p), C.c_int(x2.size),
                                              K.ctypes.data_as(C.c_void_p))
                ref = g[key]
                with np.errstate(all="ignore"):
                    assert np.array_equal(ogp.gram("matern_nu", th, X1, X2), ref), key      # oracle = reference, bit for bit
                np.testing.assert_allclose(K, ref, rtol=2e-13, atol=1e-290, err_msg=key)
                nz = ref != 0
                worst = max(worst, float(np.max(np.abs(K[nz] - ref[nz]) / np.abs(ref[nz]))))
    assert worst < 2e-13
