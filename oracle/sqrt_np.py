"""Oracle (test infrastructure): square-root covariance algebra, NumPy float64.

Restates ``src/pnmol/base/sqrt.py`` of the reference.  The QR is LAPACK ``dgeqrf``
(through ``numpy.linalg.qr``), the routine jaxlib-CPU dispatches ``jnp.linalg.qr`` /
``jax.scipy.linalg.qr`` to, so the row signs of ``R`` follow the same ``dlarfg``
convention (needed for quirk Q1).
"""
import numpy as np
import scipy.linalg


def triu_factor(stack):
    """R factor of a tall stack; reference: ``sqrtm_to_cholesky`` (sqrt.py:16-23) without
    the final transpose."""
    return np.linalg.qr(np.asarray(stack, dtype=np.float64), mode="r")


def chol_of_sum(S1, S2):
    """Lower factor of S1 S1^T + S2 S2^T; reference ``propagate_cholesky_factor``
    (sqrt.py:9-12): R-factor of vstack(S1^T, S2^T), transposed."""
    return triu_factor(np.vstack((S1.T, S2.T))).T


def _split_update(R, m, D):
    # sqrt.py:67-73 / 89-95: the three blocks of the big triangular factor
    R1 = R[:m, :m]
    R2 = R[:m, m:]
    R3 = R[m:m + D, m:m + D]
    gain = scipy.linalg.solve_triangular(R1, R2, lower=False).T
    return R3.T, gain, R1.T


def measurement_update(H, C, meas_sqrtm=None):
    """Square-root Kalman update.

    ``meas_sqrtm`` given: reference ``update_sqrt`` (sqrt.py:34-73), which pads the
    measurement factor with D-m zero columns (needs D >= m, quirk Q7).
    ``meas_sqrtm`` None: reference ``update_sqrt_no_meascov`` (sqrt.py:77-95).
    Returns (posterior factor (D,D), gain (D,m), innovation factor (m,m)).
    """
    H = np.asarray(H, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    m, D = H.shape
    big = np.zeros((2 * D, m + D))
    big[:D, :m] = C.T @ H.T
    big[:D, m:] = C.T
    if meas_sqrtm is not None:
        if D < m:
            raise ValueError("update_sqrt needs input_dim >= output_dim (sqrt.py:55-57)")
        big[D:D + m, :m] = np.asarray(meas_sqrtm, dtype=np.float64).T
    R = triu_factor(big)
    return _split_update(R, m, D)
