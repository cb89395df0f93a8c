"""GP kernel classes with the reference's surface; Gram matrices are built on the GPU.

Mirror of reference src/gpr/kernels.py (Kernel ABC :28, concrete kernels :90-382, factory :386-411).
Instances keep ``.params`` (dict), ``.bounds`` (list of (lo, hi)), ``get_params()`` -> ndarray in the
reference's parameter order, ``set_params(arr)``, ``n_params`` and ``__call__(X1, X2=None)``.  The
covariance arithmetic lives in csrc/gp_kernels.cuh (b200id_gp_gram); ``kernel_id`` / ``theta()`` are
what the batched selection and fit/predict entry points consume.
"""
from abc import ABC, abstractmethod
from typing import List, Optional, Sequence, Tuple

import numpy as np

from b200id import gp as _gp

_WIDE = (1e-3, 1e3)


class Kernel(ABC):
    """Base class: parameter bookkeeping on the host, evaluation on the device."""

    #: names of the optimisable parameters, in get_params() order
    PARAM_ORDER: Tuple[str, ...] = ()
    #: default search bounds, same order
    DEFAULT_BOUNDS: Tuple[Tuple[float, float], ...] = ()

    def __init__(self, **kwargs):
        self.params = kwargs
        self.bounds = self._get_default_bounds()

    # -- device hand-off ---------------------------------------------------------------------
    @property
    @abstractmethod
    def kernel_id(self) -> int:
        """C ABI kernel id (include/b200id.h)."""

    def theta(self) -> np.ndarray:
        """What the device receives for this kernel: the optimisable parameters followed by any fixed shape parameter."""
        return self.with_fixed(np.asarray(self.get_params(), dtype=np.float64))

    def with_fixed(self, thetas: np.ndarray) -> np.ndarray:
        """Candidate rows [..., n_params] -> device rows (fixed shape parameters appended; none by default)."""
        return np.asarray(thetas, dtype=np.float64)

    def __call__(self, X1: np.ndarray, X2: Optional[np.ndarray] = None) -> np.ndarray:
        """K(X1, X2) (X2 None -> K(X1, X1)) for 1-D inputs given as (n, 1) or (n,) arrays."""
        return _gp.gram(self.kernel_id, self.theta(), X1, X2)

    # -- reference surface -------------------------------------------------------------------
    def _get_default_bounds(self) -> List[Tuple[float, float]]:
        return [tuple(b) for b in self.DEFAULT_BOUNDS]

    def get_params(self) -> np.ndarray:
        return np.array([self.params[name] for name in self.PARAM_ORDER])

    def set_params(self, params: Sequence[float]) -> None:
        for name, value in zip(self.PARAM_ORDER, params):
            self.params[name] = value

    @property
    def n_params(self) -> int:
        return len(self.PARAM_ORDER)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.params})"


class CombinedKernel(Kernel):
    """Parameter container for a list of kernels (unused by the pipeline, kept for the surface; reference :59-86)."""

    def __init__(self, kernels: List[Kernel]):
        self.kernels = kernels
        super().__init__()

    @property
    def kernel_id(self) -> int:
        raise NotImplementedError("CombinedKernel has no device evaluation (it has none in the reference either)")

    def _get_default_bounds(self):
        return [b for k in self.kernels for b in k.bounds]

    def get_params(self):
        return np.array([p for k in self.kernels for p in k.get_params()])

    def set_params(self, params):
        pos = 0
        for k in self.kernels:
            k.set_params(params[pos:pos + k.n_params])
            pos += k.n_params

    @property
    def n_params(self):
        return sum(k.n_params for k in self.kernels)


def _simple_kernel(name: str, cid: int, order: Tuple[str, ...], bounds, defaults: dict, doc: str):
    """Class factory for kernels whose optimisable parameters are exactly their constructor arguments."""

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(order)
        if unknown:
            raise TypeError(f"{name}.__init__() got an unexpected keyword argument '{sorted(unknown)[0]}'")
        Kernel.__init__(self, **{**defaults, **kwargs})

    cls = type(name, (Kernel,), {
        "__init__": __init__, "__doc__": doc, "PARAM_ORDER": order, "DEFAULT_BOUNDS": tuple(bounds),
        "kernel_id": property(lambda self: cid),
    })
    return cls


RBFKernel = _simple_kernel(
    "RBFKernel", 0, ("length_scale", "variance"), (_WIDE, _WIDE), {"length_scale": 1.0, "variance": 1.0},
    "variance * exp(-d^2 / (2 length_scale^2))  (reference :90-110)")
RationalQuadraticKernel = _simple_kernel(
    "RationalQuadraticKernel", 11, ("length_scale", "variance", "alpha"), (_WIDE, _WIDE, _WIDE),
    {"length_scale": 1.0, "variance": 1.0, "alpha": 1.0},
    "variance * (1 + d^2 / (2 alpha length_scale^2))^-alpha  (reference :150-171)")
ExponentialKernel = _simple_kernel(
    "ExponentialKernel", 4, ("omega", "variance"), (_WIDE, _WIDE), {"omega": 1.0, "variance": 1.0},
    "variance * H(x) H(x') exp(-omega (x + x'))  (reference :175-198)")
TCKernel = _simple_kernel(
    "TCKernel", 12, ("omega", "variance"), (_WIDE, _WIDE), {"omega": 1.0, "variance": 1.0},
    "variance * H(x) H(x') exp(-omega max(x, x'))  (reference :201-224)")
DCKernel = _simple_kernel(
    "DCKernel", 5, ("alpha", "beta", "rho"), ((1e-3, 0.999), _WIDE, (-0.999, 0.999)),
    {"alpha": 0.9, "beta": 1.0, "rho": 0.5},
    "beta * alpha^((i+j)/2) * rho^|i-j| on rounded inputs  (reference :227-250)")
DIKernel = _simple_kernel(
    "DIKernel", 6, ("beta", "alpha"), (_WIDE, (1e-3, 0.999)), {"beta": 1.0, "alpha": 0.9},
    "beta * alpha^i on the rounded diagonal  (reference :253-278)")
FirstOrderStableSplineKernel = _simple_kernel(
    "FirstOrderStableSplineKernel", 7, ("beta", "variance"), (_WIDE, _WIDE), {"beta": 1.0, "variance": 1.0},
    "variance * exp(-beta min(s, t))  (reference :282-301)")
SecondOrderStableSplineKernel = _simple_kernel(
    "SecondOrderStableSplineKernel", 8, ("beta", "variance"), (_WIDE, _WIDE), {"beta": 1.0, "variance": 1.0},
    "variance * (exp(-beta (s+t+max))/2 - exp(-3 beta max)/6)  (reference :304-329)")
HighFrequencyStableSplineKernel = _simple_kernel(
    "HighFrequencyStableSplineKernel", 9, ("beta", "variance"), (_WIDE, _WIDE), {"beta": 1.0, "variance": 1.0},
    "variance * (-1)^(i+j) max(exp(-beta s), exp(-beta t))  (reference :332-355)")
StableSplineKernel = _simple_kernel(
    "StableSplineKernel", 10, ("beta", "variance"), (_WIDE, _WIDE), {"beta": 0.5, "variance": 1.0},
    "variance * r^2 (R - r/3) / 2 with r/R = min/max of exp(-beta x)  (reference :358-382)")


class MaternKernel(Kernel):
    """Matern kernel; nu is a fixed shape parameter, not optimised (reference :113-147).

    nu in {0.5, 1.5, 2.5} uses the closed forms (:124-130); any other nu > 0 the Bessel form
    dists**nu * K_nu(sqrt(2 nu) dists) * 2**(1-nu) / Gamma(nu) (:131-135) with K_nu evaluated on the device
    (csrc/gp_bessel.cuh: Temme series / Steed continued fraction + recurrence); nu then rides along as a third,
    fixed entry of the device parameter row."""
    PARAM_ORDER = ("length_scale", "variance")
    DEFAULT_BOUNDS = (_WIDE, _WIDE)
    _IDS = {0.5: 1, 1.5: 2, 2.5: 3}
    GENERAL_ID = 13

    def __init__(self, length_scale: float = 1.0, variance: float = 1.0, nu: float = 1.5):
        super().__init__(length_scale=length_scale, variance=variance, nu=nu)

    @property
    def kernel_id(self) -> int:
        nu = self.params["nu"]
        if nu in self._IDS:
            return self._IDS[nu]
        if not (isinstance(nu, (int, float)) and 0.0 < float(nu) <= 64.0):
            raise ValueError(f"Matern nu={nu!r}: the device kernel needs 0 < nu <= 64")
        return self.GENERAL_ID

    def with_fixed(self, thetas: np.ndarray) -> np.ndarray:
        th = np.asarray(thetas, dtype=np.float64)
        if self.params["nu"] in self._IDS:
            return th
        nu = np.full(th.shape[:-1] + (1,), float(self.params["nu"]))
        return np.concatenate([th, nu], axis=-1)


def create_kernel(kernel_type: str, **kwargs) -> Kernel:
    """Kernel by short name: rbf, matern, matern12, matern32, matern52, rq, exp, tc, dc, di, ss1, ss2, sshf,
    stable_spline (reference :386-411)."""
    def fixed_nu(nu):
        return lambda **kw: MaternKernel(nu=nu, **{k: v for k, v in kw.items() if k != "nu"})

    table = {
        "rbf": RBFKernel, "matern": MaternKernel, "matern12": fixed_nu(0.5), "matern32": fixed_nu(1.5),
        "matern52": fixed_nu(2.5), "rq": RationalQuadraticKernel, "exp": ExponentialKernel, "tc": TCKernel,
        "dc": DCKernel, "di": DIKernel, "ss1": FirstOrderStableSplineKernel, "ss2": SecondOrderStableSplineKernel,
        "sshf": HighFrequencyStableSplineKernel, "stable_spline": StableSplineKernel,
    }
    if kernel_type not in table:
        raise ValueError(f"Unknown kernel type: {kernel_type}. Available: {list(table.keys())}")
    return table[kernel_type](**kwargs)
