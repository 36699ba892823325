"""NumPy restatement of src/pnmol/base/kalman.py (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Pinned to outputs of the reference's own source (tests/golden/reference_kalman.npz, produced by
tests/golden/make_reference_golden.py) in tests/test_reference_golden.py."""
import numpy as np
import scipy.linalg

from . import sqrt_np


def filter_step(m, sc, phi, sq, h, b, data):
    """kalman.py:12-31."""
    m_pred = phi @ m
    x1 = phi @ sc
    sc_pred = sqrt_np.chol_of_sum(x1, sq)
    cross = (x1 @ sc.T).T
    sgain = scipy.linalg.cho_solve((sc_pred, True), cross.T).T
    sc_new, kgain, _ = sqrt_np.measurement_update(h, sc_pred, None)
    z = h @ m_pred + b
    return m_pred - kgain @ (z - data), sc_new, sgain, m_pred, sc_pred, x1


def smoother_step_traditional(m, sc, m_fut, sc_fut, sgain, mp, scp):
    """kalman.py:35-46."""
    new_cov = sc @ sc.T + sgain @ (sc_fut @ sc_fut.T - scp @ scp.T) @ sgain.T
    return m + sgain @ (m_fut - mp), np.linalg.cholesky(new_cov)


def smoother_step_sqrt(m, sc, m_fut, sc_fut, sgain, sq, mp, x):
    """kalman.py:49-66."""
    d = m.shape[0]
    zeros = np.zeros((d, d))
    M = np.block([[x.T, sc.T], [sq.T, zeros.T], [zeros.T, sc_fut.T @ sgain.T]])
    R = scipy.linalg.qr(M, mode="r")[0]
    return m - sgain @ (mp - m_fut), R[d:2 * d, d:].T
