"""Synthetic records for the FRF -> GPR -> FIR path (tests, golden vectors, bench).

The reference's recordings are not in its checkout, so every workload here is a
stand-in with the same layout:

* timebase ``t = i * dt`` with dt = 0.002 s (500 Hz), FP64;
* multisine excitation on the ``matlab_freq_grid(N_d, -1, 2.3)`` frequencies with
  amplitudes ``U(0, 20/N_d)`` and phases ``U(0, 2*pi)``
  (reference: flexible_link/repeated_daqne.m:28,43-45);
* random-pulse square wave for validation (reference: flexible_link/square/pulse.m:13-20);
* plant = bilinear-discretised 2nd-order low-pass x lightly damped resonance
  (n_b = 2 / n_a = 4 in continuous time), applied with scipy.signal.lfilter,
  plus small Gaussian output noise;
* ``.mat`` layout: one variable ``output`` = 3 x N ``[t; y; u]``
  (reference: data/sample_data/README.md:25-34).

Pure host NumPy/SciPy. Nothing here is on the timed path.
"""
from __future__ import annotations

from pathlib import Path
from typing import Tuple

import numpy as np

DT = 0.002
F_LOW_LOG10 = -1.0
F_UP_LOG10 = 2.3


def log_grid(nd: int, f_low: float = F_LOW_LOG10, f_up: float = F_UP_LOG10) -> Tuple[np.ndarray, np.ndarray]:
    """(freqs_hz, omega) on the reference's MATLAB-style log grid (upper end excluded)."""
    step = (f_up - f_low) / nd
    logs = f_low + step * np.arange(nd, dtype=np.float64)
    f = np.power(10.0, logs)
    return f, 2.0 * np.pi * f


def plant_tf():
    """Continuous-time numerator/denominator of the flexible-link-like plant."""
    w1, z1 = 2.0 * np.pi * 2.0, 0.7      # low-pass section
    w2, z2 = 2.0 * np.pi * 9.0, 0.04     # lightly damped resonance
    num = np.array([w1 * w1 * w2 * w2])
    den = np.polymul([1.0, 2 * z1 * w1, w1 * w1], [1.0, 2 * z2 * w2, w2 * w2])
    return num, den


def plant_response(omega: np.ndarray) -> np.ndarray:
    """Continuous-time G(j*omega) of :func:`plant_tf`."""
    num, den = plant_tf()
    s = 1j * np.asarray(omega, dtype=np.float64)
    return np.polyval(num, s) / np.polyval(den, s)


def apply_plant(u: np.ndarray, dt: float = DT, noise: float = 1e-3, seed: int = 7) -> np.ndarray:
    from scipy.signal import bilinear, lfilter
    num, den = plant_tf()
    b, a = bilinear(num, den, fs=1.0 / dt)
    y = lfilter(b, a, u)
    if noise > 0:
        y = y + noise * np.random.default_rng(seed).standard_normal(u.shape[0])
    return np.ascontiguousarray(y, dtype=np.float64)


def multisine(n: int, nd: int, dt: float = DT, seed: int = 0, chunk: int = 1 << 18) -> Tuple[np.ndarray, np.ndarray]:
    """(t, u) multisine record of ``n`` samples on the ``nd``-bin log grid."""
    rng = np.random.default_rng(seed)
    f, _ = log_grid(nd)
    amp = rng.uniform(0.0, 20.0 / nd, size=nd)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=nd)
    t = np.arange(n, dtype=np.float64) * dt
    u = np.empty(n, dtype=np.float64)
    w = 2.0 * np.pi * f
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        u[s:e] = np.sin(np.outer(t[s:e], w) + ph[None, :]) @ amp
    return t, u


def square_wave(n: int, dt: float = DT, seed: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """(t, u) random-amplitude / period / duty pulse train."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) * dt
    u = np.zeros(n, dtype=np.float64)
    pos = 0
    while pos < n:
        amp = rng.uniform(0.1, 5.0)
        period = rng.uniform(0.1, 10.0)
        duty = rng.uniform(0.1, 0.9)
        reps = int(rng.integers(2, 8))
        p = max(2, int(round(period / dt)))
        hi = max(1, int(round(duty * p)))
        for _ in range(reps):
            if pos >= n:
                break
            u[pos:min(n, pos + hi)] = amp
            pos += p
    return t, u


def make_record(n: int, nd: int, kind: str = "multisine", seed: int = 0, noise: float = 1e-3):
    """(t, u, y) synthetic record."""
    if kind == "multisine":
        t, u = multisine(n, nd, seed=seed)
    elif kind == "square":
        t, u = square_wave(n, seed=seed)
    else:
        raise ValueError(f"unknown record kind {kind!r}")
    y = apply_plant(u, noise=noise, seed=seed + 7)
    return t, u, y


def write_mat(path, t: np.ndarray, u: np.ndarray, y: np.ndarray) -> Path:
    """Write the reference's ``output`` = 3 x N ``[t; y; u]`` layout."""
    from scipy.io import savemat
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    savemat(str(path), {"output": np.vstack([t, y, u])})
    return path
