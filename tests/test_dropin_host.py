"""CPU-side checks of the drop-in surface that need no GPU: names, parameter bookkeeping, host I/O,
error behaviour, and (build container only) that the overlay hands the reference's own callers our
three packages."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_kernel_surface_matches_reference_tables(golden):
    from src.gpr.kernels import create_kernel, Kernel
    g = golden("gp")
    for name in ["rbf", "matern12", "matern32", "matern52", "exp", "dc", "di", "ss1", "ss2", "sshf", "stable_spline", "rq", "tc"]:
        k = create_kernel(name)
        assert isinstance(k, Kernel)
        assert np.array_equal(np.array(k.bounds), g[f"bounds_{name}"])
        assert np.array_equal(k.get_params(), g[f"params0_{name}"])
        assert k.n_params == len(k.bounds)
        k.set_params(g[f"theta_{name}"][0])
        assert np.array_equal(k.get_params(), g[f"theta_{name}"][0])
        assert name.split("_")[0][:2].lower() in repr(k).lower() or "Kernel" in repr(k)
    assert create_kernel("matern").params["nu"] == 1.5 and create_kernel("matern", nu=2.5).kernel_id == 3
    assert create_kernel("matern52", nu=0.5).params["nu"] == 2.5      # nu kwarg ignored like the reference
    with pytest.raises(ValueError, match="Unknown kernel type"):
        create_kernel("nope")


def test_default_grids(golden):
    from src.gpr.kernels import create_kernel
    from src.gpr.grid_search import get_default_param_grids
    g = golden("gp")
    for name in ["rbf", "dc", "di", "stable_spline"]:
        grids = get_default_param_grids(create_kernel(name))
        assert list(grids) == [f"param_{i}" for i in range(len(grids))]
        assert np.array_equal(np.stack([grids[f"param_{i}"] for i in range(len(grids))]), g[f"grid_{name}"])


def test_host_io_and_grid(tmp_path, golden):
    from b200id import synth
    from src.frequency_transform.frf_estimator import (load_time_u_y, matlab_freq_grid, describe_timebase,
                                                       resolve_mat_files, try_load_frequency_from_mat)
    from scipy.io import savemat
    g = golden("frf")
    hz, w = matlab_freq_grid(37, -1.0, 2.3)
    assert np.array_equal(hz, g["grid_hz_37"]) and np.array_equal(w, g["grid_w_37"])
    p = synth.write_mat(tmp_path / "rec.mat", g["t"], g["u"], g["y"])
    t, u, y = load_time_u_y(p)
    assert np.array_equal(t, g["t"]) and np.array_equal(u, g["u"]) and np.array_equal(y, g["y"])
    savemat(str(tmp_path / "named.mat"), {"t": g["t"][:100], "u": g["u"][:100], "y": np.c_[g["y"][:100], g["u"][:100]]})
    t2, u2, y2 = load_time_u_y(tmp_path / "named.mat", y_col=1)
    assert np.array_equal(y2, g["u"][:100])
    with pytest.raises(ValueError, match="out of range"):
        load_time_u_y(tmp_path / "named.mat", y_col=5)
    savemat(str(tmp_path / "bad.mat"), {"z": np.zeros((4, 4))})
    with pytest.raises(RuntimeError, match="Could not locate"):
        load_time_u_y(tmp_path / "bad.mat")
    assert describe_timebase(g["t"])[2] is True and describe_timebase(g["t"][::-1].copy())[2] is False
    assert [q.name for q in resolve_mat_files([str(tmp_path / "*.mat")])] == ["bad.mat", "named.mat", "rec.mat"]
    savemat(str(tmp_path / "freq.mat"), {"frequency": np.array([1.0, 2.0])})
    assert np.allclose(try_load_frequency_from_mat(tmp_path / "freq.mat"), 2 * np.pi * np.array([1.0, 2.0]))


def test_fir_helpers_host(golden):
    from src.fir_model.fir_helpers import build_two_sided_hermitian_from_positive, build_uniform_omega_linear, _load_timebase_from_mat
    from src.fir_model.fir_fitting import gp_to_fir_direct_pipeline
    g = golden("fir")
    assert np.array_equal(build_two_sided_hermitian_from_positive(g["G"][:9]), g["herm_small"])
    wu, Gu = build_uniform_omega_linear(g["w"], g["G"], 7)
    assert wu.shape == (7,) and Gu[0] == g["G"][0] and Gu[-1] == g["G"][-1]
    with pytest.raises(NotImplementedError, match="legacy"):
        gp_to_fir_direct_pipeline(g["w"], g["G"], paper_mode=False, output_dir=Path("/tmp/_b200id_unused"))


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: on a machine without CUDA the hot-path calls raise instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from src.frequency_transform.frf_estimator import synchronous_coefficients_trapz
    from src.gpr import create_kernel
    t = np.arange(100) * 0.002
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        synchronous_coefficients_trapz(t, t, np.array([1.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        create_kernel("rbf")(t[:4])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gauss-process_b200/ may import it."""
    for py in (ROOT / "gauss-process_b200").rglob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py


@pytest.mark.needs_reference
def test_overlay_serves_reference_callers():
    """With gauss-process_b200 first on sys.path the reference's unchanged run_gp_pipeline / run_baseline
    import OUR src.frequency_transform / src.gpr / src.fir_model (matplotlib stubbed here)."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from oracle import refharness; refharness.install_matplotlib_stub()\n"
        "import src; from src.pipeline import gp_pipeline; import src.run_baseline as rb\n"
        "import src.gpr.gpr_fitting as gf, src.frequency_transform.transform as tr, src.fir_model.fir_fitting as ff\n"
        "assert '/root/reference' in gp_pipeline.__file__ and '/root/reference' in rb.__file__\n"
        "for m in (gf, tr, ff): assert 'gauss-process_b200' in m.__file__, m.__file__\n"
        "assert gp_pipeline.GaussianProcessRegressor is gf.GaussianProcessRegressor\n"
        "print('overlay ok')\n" % (str(ROOT), str(ROOT / "gauss-process_b200")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "overlay ok" in out.stdout, out.stderr[-2000:]
