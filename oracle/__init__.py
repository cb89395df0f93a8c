"""CPU oracle for the FRF -> GPR-selection -> FIR hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the
reference's algorithm (AtsushiYanaigsawa768/gauss-process, pure Python).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline -- never as part of the product path (the product lives in
``gauss-process_b200/`` and fails loudly without its CUDA library).

Parity pinning: the reference's own tests hold no golden vectors (SURVEY.md
section 4), so this oracle is pinned against outputs of the reference itself,
generated in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference`` behind a matplotlib stub) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle function
against those vectors (bit-exact where the operation order is restated
exactly, <= a few ulp otherwise).

Third-party arithmetic the reference calls (not vendored in its tree; its
requirements.txt pins numpy 1.26.4 / scipy 1.12.0, the container has numpy
2.3.5 / scipy 1.18.1 which is what the golden vectors were produced with):
np.trapz, np.linalg.cholesky/solve/inv/pinv, scipy.spatial.distance.cdist,
np.fft.ifft, np.convolve, sklearn StandardScaler.
"""
from . import frf, gp, fir  # noqa: F401
