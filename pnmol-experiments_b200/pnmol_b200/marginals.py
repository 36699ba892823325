"""Marginal read-outs of EK1 solutions on the device.

The reference computes them in its experiment scripts from the full covariance,
``stds = sqrt(diag(cov_sqrtm @ cov_sqrtm.T) @ E0.T)`` (experiments/figure1.py:76-89, figure3.py:87-93); here the
standard deviation of state component j is the norm of row ``j (nu+1)`` of the factor, computed by a device kernel
(`pnmol_b200_marginal_std`) or, for whole trajectories, fused into the persistent step kernel
(`PDEFilter.solve_marginals`, `pnmol_b200_run_marginals`) so that the ``T x D x D`` factor trajectory is never written.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _engine

MarginalSolution = namedtuple("MarginalSolution", ["t", "mean", "std", "info", "diffusion_squared_calibrated"])


def _num_derivatives(E0, D):
    E0 = np.asarray(E0)
    return D // E0.shape[0] - 1


def read_mean_and_std(sol, E0):
    """experiments/figure1.py:76-80 for a white-noise PDESolution: (means (T, d), stds (T, d))."""
    means = sol.mean[:, 0]
    nu = _num_derivatives(E0, sol.cov_sqrtm.shape[-1])
    return means, _engine.marginal_std(sol.cov_sqrtm, nu)


def read_mean_and_std_latent(sol, E0):
    """experiments/figure1.py:83-89 for a latent-force PDESolution: the state half of mean and std."""
    d = np.asarray(E0).shape[0]
    means = sol.mean[:, 0, :d]
    nu = sol.cov_sqrtm.shape[-1] // (2 * d) - 1
    return means, _engine.marginal_std(sol.cov_sqrtm, nu)[..., :d]


def read_mean_and_std_and_cov(final_state, E0):
    """experiments/figure3.py:87-93 for a final PDEFilterState: (mean (d,), std (d,), cov (d, d))."""
    E0t = torch.as_tensor(np.asarray(E0), dtype=torch.float64, device=final_state.y.cov_sqrtm.device)
    L0 = E0t @ final_state.y.cov_sqrtm
    nu = _num_derivatives(E0, final_state.y.cov_sqrtm.shape[-1])
    return final_state.y.mean[0], _engine.marginal_std(final_state.y.cov_sqrtm, nu)[: E0t.shape[0]], L0 @ L0.T
