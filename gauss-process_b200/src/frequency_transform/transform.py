"""Unified entry point ``estimate_frequency_response`` (mirror of reference
src/frequency_transform/transform.py:45-193).

``method='frf'`` runs on the GPU: per file one fused pass over (t, u, y) for both signals, then one
cross-power average over the records.  ``method='fourier'`` (the FFT estimator on a linear grid) and
the optional disk cache are outside the hot path: they are taken from the
reference checkout when one is on ``src.__path__`` and raise NotImplementedError otherwise.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional

import numpy as np
import pandas as pd

from src.frequency_transform.frf_estimator import (  # noqa: F401  (re-exported like the reference)
    resolve_mat_files,
    load_time_u_y,
    try_load_frequency_from_mat,
    matlab_freq_grid,
    describe_timebase,
    synchronous_coefficients_trapz,
    synchronous_coefficients_trapz_pair,
    frf_cross_power_average,
)

DEFAULT_ND = 100
DEFAULT_F_LOW_LOG10 = -1.0
DEFAULT_F_UP_LOG10 = 2.3
DEFAULT_DROP_SECONDS = 0.0
DEFAULT_WINDOW = "hann"


def estimate_frequency_response(
    mat_files: List[str],
    method: str = "frf",
    nd: int = DEFAULT_ND,
    f_low: float = DEFAULT_F_LOW_LOG10,
    f_up: float = DEFAULT_F_UP_LOG10,
    drop_seconds: float = DEFAULT_DROP_SECONDS,
    time_duration: Optional[float] = None,
    n_files: Optional[int] = None,
    y_col: int = 0,
    subtract_mean: bool = True,
    window: str = DEFAULT_WINDOW,
    use_cache: bool = False,
    cache_dir: Optional[Path] = None,
) -> pd.DataFrame:
    """DataFrame [omega_rad_s, freq_Hz, ReG, ImG, absG, phase_rad] from .mat time series (reference :45-113)."""
    paths = resolve_mat_files(mat_files)
    if not paths:
        raise FileNotFoundError("No MAT files found for the given input.")
    if n_files is not None:
        paths = paths[:n_files]

    cache = key = None
    if use_cache:
        try:
            from src.frequency_transform.cache import FrequencyDataCache     # reference host module
        except ImportError as exc:
            raise NotImplementedError("the FRF disk cache is host I/O outside the b200id hot path; "
                                      "put a reference checkout on GP_REFERENCE_ROOT to use it") from exc
        cache = FrequencyDataCache(cache_dir) if cache_dir else FrequencyDataCache()
        key = cache.get_cache_key([str(p) for p in paths], len(paths), time_duration, nd, method)
        hit = cache.get(key)
        if hit is not None:
            return hit

    if method == "frf":
        df = _estimate_frf(paths, nd, f_low, f_up, drop_seconds, time_duration, y_col, subtract_mean)
    elif method == "fourier":
        df = _estimate_fourier(paths, nd, f_low, f_up, drop_seconds, time_duration, y_col, window)
    else:
        raise ValueError(f"Unknown method '{method}'. Choose 'frf' or 'fourier'.")
    if cache is not None:
        cache.put(key, df)
    return df


def _estimate_frf(paths, nd, f_low, f_up, drop_seconds, time_duration, y_col, subtract_mean):
    """Per file: window -> monotone order -> fused GPU demodulation of u and y; then the cross-power
    average over files and the (defensive) sort by omega (reference :120-157)."""
    _, omega = matlab_freq_grid(nd, f_low, f_up)
    U_list, Y_list = [], []
    for mat_path in paths:
        t, u, y = load_time_u_y(mat_path, y_col=y_col)
        drop = drop_seconds
        if time_duration is not None:
            t_start = t[0] + drop_seconds
            t_end = t_start + time_duration
            keep = (t >= t_start) & (t <= t_end)
            if not np.any(keep):
                raise ValueError(f"No data in [{t_start:.2f}, {t_end:.2f}]")
            t, u, y = t[keep], u[keep], y[keep]
            drop = 0.0
        if not describe_timebase(t)[2]:
            order = np.argsort(t)
            t, u, y = t[order], u[order], y[order]
        U, Y = synchronous_coefficients_trapz_pair(t, u, y, omega, drop, subtract_mean)
        U_list.append(U)
        Y_list.append(Y)
    G = frf_cross_power_average(U_list, Y_list)
    order = np.argsort(omega)
    return _build_dataframe(omega[order], G[order])


def _estimate_fourier(paths, nd, f_min, f_max, drop_seconds, time_duration, y_col, window):
    try:
        from src.frequency_transform import fourier_estimator as fe          # reference host module
    except ImportError as exc:
        raise NotImplementedError("method='fourier' is outside the b200id hot path (SURVEY.md section 2 row 3); "
                                  "put a reference checkout on GP_REFERENCE_ROOT to use it") from exc
    win = None if window == "none" else window
    estimates, f_top = [], 0.0
    for mat_path in paths:
        t, u, y = fe.load_time_data(mat_path, y_col=y_col)
        t, u, y = fe.apply_time_window(t, u, y, drop_seconds, time_duration)
        freq, G = fe.compute_transfer_function(t, u, y, window=win)
        f_top = max(f_top, freq[-1])
        estimates.append((freq, G))
    f_grid = fe.linear_frequency_grid(f_min, min(f_max, f_top) if f_max is not None else f_top, nd)
    on_grid = [fe.interpolate_to_grid(freq, G, f_grid) for freq, G in estimates]
    G_avg = fe.average_multiple_estimates(on_grid) if len(on_grid) > 1 else on_grid[0]
    return _build_dataframe(2.0 * np.pi * f_grid, G_avg)


def _build_dataframe(omega: np.ndarray, G: np.ndarray) -> pd.DataFrame:
    return pd.DataFrame({
        "omega_rad_s": omega,
        "freq_Hz": omega / (2.0 * np.pi),
        "ReG": np.real(G),
        "ImG": np.imag(G),
        "absG": np.abs(G),
        "phase_rad": np.angle(G),
    })
