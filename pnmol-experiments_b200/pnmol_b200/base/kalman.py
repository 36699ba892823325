"""Gaussian filtering and smoothing routines on the GPU (API of src/pnmol/base/kalman.py).

``smoother_step_sqrt`` (kalman.py:49-66) is one batched CUDA kernel (`pnmol_b200_smoother_step`: the third QR shape of
the code base, 3d x 2d).  ``filter_step`` (kalman.py:12-31) composes the square-root primitives of ``base.sqrt`` with a
few small dense products and a Cholesky solve on the device; ``smoother_step_traditional`` (kalman.py:35-46) is the
covariance-form check the reference's own tests compare against.  None of this is on the EK1 path of the PDE solvers
(the reference only calls it from odetools/init.py); it is the follow-on SURVEY section 8f ranks fourth.
"""
import torch

from .. import _lib
from . import sqrt
from .sqrt import _prep


def filter_step(m, sc, phi, sq, h, b, data):
    """kalman.py:12-31: returns (m, sc, sgain, m_pred, sc_pred, x1)."""
    m, sc, phi, sq, h, b, data = (_prep(a) for a in (m, sc, phi, sq, h, b, data))
    m_pred = phi @ m
    x1 = phi @ sc
    sc_pred = sqrt.propagate_cholesky_factor(x1, sq)
    cross = (x1 @ sc.T).T
    sgain = torch.cholesky_solve(cross.T, sc_pred).T  # cho_solve((sc_pred, lower), cross^T)^T
    sc_new, kgain, _ = sqrt.update_sqrt_no_meascov(h, sc_pred)
    z = h @ m_pred + b
    return m_pred - kgain @ (z - data), sc_new, sgain, m_pred, sc_pred, x1


def smoother_step_traditional(m, sc, m_fut, sc_fut, sgain, mp, scp):
    """kalman.py:35-46 (covariance form)."""
    m, sc, m_fut, sc_fut, sgain, mp, scp = (_prep(a) for a in (m, sc, m_fut, sc_fut, sgain, mp, scp))
    new_mean = m + sgain @ (m_fut - mp)
    new_cov = sc @ sc.T + sgain @ (sc_fut @ sc_fut.T - scp @ scp.T) @ sgain.T
    return new_mean, torch.linalg.cholesky(new_cov)


def smoother_step_sqrt(m, sc, m_fut, sc_fut, sgain, sq, mp, x):
    """kalman.py:49-66; every argument may carry a leading batch dimension."""
    args = [_prep(a) for a in (m, sc, m_fut, sc_fut, sgain, sq, mp, x)]
    batched = args[0].dim() == 2
    if not batched:
        args = [a[None] for a in args]
    args = [a.contiguous() for a in args]
    B, d = args[0].shape
    mean_out = torch.empty((B, d), dtype=torch.float64, device=args[0].device)
    chol_out = torch.empty((B, d, d), dtype=torch.float64, device=args[0].device)
    _lib.check(_lib.load().pnmol_b200_smoother_step(*[_lib.ptr(a) for a in args], _lib.ptr(mean_out), _lib.ptr(chol_out), d, B,
                                                    args[0].device.index, _lib.current_stream(args[0].device)))
    return (mean_out, chol_out) if batched else (mean_out[0], chol_out[0])
