"""Integrated Wiener process prior (API of src/pnmol/base/iwp.py).

The CUDA step never forms the dense ``D x D`` matrices: it receives ``A_1d``, ``L_Q1d``
and the spatial factor and applies the Kronecker structure directly; the dense
assemblies below exist for API parity (tests, read-outs such as ``E0`` in
experiments/figure1.py:19) and are host-side NumPy.
"""
from collections import namedtuple
from functools import cached_property

import numpy as np
import scipy.linalg
import scipy.special


class IntegratedWienerTransition(namedtuple("_IWP", "wiener_process_dimension num_derivatives wp_diffusion_sqrtm")):
    @cached_property
    def preconditioned_discretize_1d(self):
        """iwp.py:13-30."""
        n = self.num_derivatives + 1
        A_1d = np.flip(scipy.linalg.pascal(n, kind="lower", exact=False)).astype(np.float64)
        return A_1d, np.linalg.cholesky(np.flip(scipy.linalg.hilbert(n)))

    @cached_property
    def preconditioned_discretize(self):
        """iwp.py:32-53."""
        A_1d, L_Q1d = self.preconditioned_discretize_1d
        return (np.kron(np.eye(self.wiener_process_dimension), A_1d),
                np.kron(np.asarray(self.wp_diffusion_sqrtm), L_Q1d))

    def nordsieck_preconditioner_1d_raw(self, dt):
        """iwp.py:55-62."""
        powers = np.arange(self.num_derivatives, -1, -1)
        scales = scipy.special.factorial(powers)
        powers = powers + 0.5
        return (np.abs(dt) ** powers) / scales, (np.abs(dt) ** (-powers)) * scales

    def nordsieck_preconditioner_1d(self, dt):
        p, pinv = self.nordsieck_preconditioner_1d_raw(dt)
        return np.diag(p), np.diag(pinv)

    def nordsieck_preconditioner(self, dt):
        """iwp.py:79-97."""
        p, pinv = self.nordsieck_preconditioner_1d(dt)
        eye = np.eye(self.wiener_process_dimension)
        return np.kron(eye, p), np.kron(eye, pinv)

    def non_preconditioned_discretize(self, dt):
        """iwp.py:99-122."""
        P, Pinv = self.nordsieck_preconditioner(dt)
        A, LQ = self.preconditioned_discretize
        return P @ A @ Pinv, P @ LQ

    def projection_matrix(self, derivative_to_project_onto):
        return np.kron(np.eye(self.wiener_process_dimension), self.projection_matrix_1d(derivative_to_project_onto))

    def projection_matrix_1d(self, derivative_to_project_onto):
        return np.eye(1, self.num_derivatives + 1, derivative_to_project_onto)

    @property
    def state_dimension(self):
        return self.wiener_process_dimension * (self.num_derivatives + 1)
