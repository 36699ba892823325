"""Oracle (test infrastructure): integrated-Wiener-process prior, NumPy float64.

Restates ``src/pnmol/base/iwp.py`` and ``src/pnmol/base/stacked_ssm.py`` in dense form,
exactly as the reference builds them (Kronecker products, dense diagonal
preconditioners).
"""
import numpy as np
import scipy.linalg
import scipy.special


def iwp_1d(nu):
    """(A_1d, L_Q1d) of the preconditioned IWP(nu); iwp.py:13-30: flipped lower Pascal and
    Cholesky factor of the flipped Hilbert matrix."""
    n = nu + 1
    A = np.flip(scipy.linalg.pascal(n, kind="lower", exact=False)).astype(np.float64)
    Q = np.flip(scipy.linalg.hilbert(n))
    return A, np.linalg.cholesky(Q)


def iwp_dense(nu, diffusion_sqrtm):
    """(A, L_Q) = (I_d kron A_1d, diffusion_sqrtm kron L_Q1d); iwp.py:32-53."""
    A1, LQ1 = iwp_1d(nu)
    d = diffusion_sqrtm.shape[0]
    return np.kron(np.eye(d), A1), np.kron(diffusion_sqrtm, LQ1)


def nordsieck_scales(nu, dt):
    """Scaling vector and inverse; iwp.py:55-62."""
    powers = np.arange(nu, -1, -1)
    fact = scipy.special.factorial(powers)
    powers = powers + 0.5
    return (np.abs(dt) ** powers) / fact, (np.abs(dt) ** (-powers)) * fact


def nordsieck_dense(nu, d, dt):
    """Dense (P, P^-1) = I_d kron diag(.); iwp.py:64-97."""
    p, pinv = nordsieck_scales(nu, dt)
    eye = np.eye(d)
    return np.kron(eye, np.diag(p)), np.kron(eye, np.diag(pinv))


def projection(nu, d, deriv):
    """E_deriv = I_d kron e_deriv^T; iwp.py:125-133."""
    return np.kron(np.eye(d), np.eye(1, nu + 1, deriv))


def non_preconditioned(nu, diffusion_sqrtm, dt):
    """iwp.py:99-122 (testing helper)."""
    d = diffusion_sqrtm.shape[0]
    P, Pinv = nordsieck_dense(nu, d, dt)
    A, LQ = iwp_dense(nu, diffusion_sqrtm)
    return P @ A @ Pinv, P @ LQ


def stack_blockdiag(*mats):
    """stacked_ssm.py:16-48 assemble block diagonals of the per-process matrices."""
    return scipy.linalg.block_diag(*mats)
