"""jax.scipy.linalg -> scipy.linalg (LAPACK through OpenBLAS, the routines jaxlib's CPU backend calls)."""
import scipy.linalg as _sl
from scipy.linalg import block_diag, cho_factor, cho_solve, cholesky, solve, solve_triangular  # noqa: F401


def qr(a, overwrite_a=False, lwork=None, mode="full", pivoting=False, check_finite=True):
    """jax.scipy.linalg.qr of the JAX versions the reference was written for returns the bare R for mode="r"
    (src/pnmol/base/sqrt.py:66-70 slices it as an array), SciPy a 1-tuple."""
    if pivoting:
        raise NotImplementedError
    out = _sl.qr(a, mode=mode)
    return out[0] if mode == "r" else out
