"""PDE problem containers (host-side set-up).

The reference composes its problem classes from mix-ins (src/pnmol/pde/problems.py:11-108,
src/pnmol/pde/mixins.py:16-295).  What the EK1 solvers consume is a discretised problem
with the attributes ``L, E_sqrtm, B, R_sqrtm, y0, t0, tmax, mesh_spatial`` and, for
semi-linear problems, ``f, df``; this module provides exactly that with the same class
names, attribute names and ``discretize`` / ``discretize_system`` methods.  The classical
method-of-lines conversion (``to_tornadox_ivp``) belongs to the third-party tornadox
baseline and is out of scope.
"""
from collections import namedtuple

import numpy as np
import scipy.linalg

from .. import discretize

#: Device-side description of a point-wise reaction term (id in include/pnmol_b200.h).
Reaction = namedtuple("Reaction", "id params ncomp")

REACTION_IDS = {"none": 0, "spruce": 1, "sir": 2, "lotka_volterra": 3}


class PDE:
    """problems.py:11-42 plus the IVP / boundary / non-linearity attributes of mixins.py."""

    bcond = None  # "dirichlet" | "neumann"
    is_system = False
    is_semilinear = False

    def __init__(self, *, diffop, diffop_scale, bbox, t0=None, tmax=None, y0_fun=None, f=None, df=None,
                 df_diagonal=None, reaction=None):
        self.diffop, self.diffop_scale = diffop, diffop_scale
        self.bbox = np.asarray(bbox, dtype=np.float64)
        self.t0, self.tmax, self.y0_fun = t0, tmax, y0_fun
        self.f, self.df, self.df_diagonal = f, df, df_diagonal
        self.reaction = reaction
        self.L = self.E_sqrtm = self.mesh_spatial = None
        self.B = self.R_sqrtm = self.y0 = None

    def __repr__(self):
        return f"{type(self).__name__}(is_discretized={self.is_discretized})"

    @property
    def is_discretized(self):
        return self.L is not None

    @property
    def dimension(self):
        return self.bbox.ndim

    @property
    def t_span(self):
        return self.t0, self.tmax

    @property
    def num_components(self):
        return len(self.diffop) if self.is_system else 1

    def _boundary_operator(self, mesh_spatial, kernel, nugget_gram_matrix):
        if self.bcond == "neumann":  # mixins.py:41-49
            if self.dimension > 1:
                raise NotImplementedError
            return discretize.fd_probabilistic_neumann_1d(mesh_spatial=mesh_spatial, kernel=kernel, stencil_size=2,
                                                          nugget_gram_matrix=nugget_gram_matrix)
        B = mesh_spatial.boundary_projection_matrix  # mixins.py:51-54
        return B, np.zeros((B.shape[0], B.shape[0]))

    def discretize(self, *, mesh_spatial, kernel, stencil_size_interior, stencil_size_boundary, nugget_gram_matrix=0.0):
        """mixins.py:19-59."""
        L, E = discretize.fd_probabilistic(self.diffop, mesh_spatial=mesh_spatial, kernel=kernel,
                                           stencil_size_interior=stencil_size_interior,
                                           stencil_size_boundary=stencil_size_boundary,
                                           nugget_gram_matrix=nugget_gram_matrix)
        self.L, self.E_sqrtm = self.diffop_scale * L, self.diffop_scale * E
        self.mesh_spatial = mesh_spatial
        self.B, self.R_sqrtm = self._boundary_operator(mesh_spatial, kernel, nugget_gram_matrix)
        if self.y0_fun is not None:
            self.y0 = np.asarray(self.y0_fun(mesh_spatial.points), dtype=np.float64)[:, 0]

    def discretize_system(self, *, mesh_spatial, kernel, stencil_size_interior, stencil_size_boundary,
                          nugget_gram_matrix=0.0):
        """mixins.py:66-122 (one operator per component, block-diagonal result)."""
        if self.bcond != "neumann":
            raise NotImplementedError("the reference's system + Dirichlet discretisation is broken (mixins.py:112-114)")
        Ls, Es = [], []
        for op, scale in zip(self.diffop, self.diffop_scale):
            L, E = discretize.fd_probabilistic(op, mesh_spatial=mesh_spatial, kernel=kernel,
                                               stencil_size_interior=stencil_size_interior,
                                               stencil_size_boundary=stencil_size_boundary,
                                               nugget_gram_matrix=nugget_gram_matrix)
            Ls.append(scale * L)
            Es.append(scale * E)
        self.L, self.E_sqrtm = scipy.linalg.block_diag(*Ls), scipy.linalg.block_diag(*Es)
        self.mesh_spatial = mesh_spatial
        B, R = self._boundary_operator(mesh_spatial, kernel, nugget_gram_matrix)
        c = len(self.diffop)
        self.B, self.R_sqrtm = scipy.linalg.block_diag(*([B] * c)), scipy.linalg.block_diag(*([R] * c))
        if self.y0_fun is not None:
            self.y0 = np.asarray(self.y0_fun(mesh_spatial.points), dtype=np.float64).squeeze()


class LinearEvolutionDirichlet(PDE):
    bcond = "dirichlet"


class LinearEvolutionNeumann(PDE):
    bcond = "neumann"


class SemiLinearEvolutionDirichlet(PDE):
    bcond, is_semilinear = "dirichlet", True


class SemiLinearEvolutionNeumann(PDE):
    bcond, is_semilinear = "neumann", True


class SystemLinearPDENeumann(PDE):
    bcond, is_system = "neumann", True


class SystemSemiLinearEvolutionNeumann(PDE):
    bcond, is_system, is_semilinear = "neumann", True, True
