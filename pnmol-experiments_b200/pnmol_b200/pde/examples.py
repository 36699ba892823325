"""Problem recipes with the signatures of src/pnmol/pde/examples.py.

Each semi-linear recipe carries, next to the NumPy callables ``f`` / ``df`` (the
reference's ``pde.f`` / ``pde.df``), a ``Reaction`` tag so that the CUDA step can evaluate
the same point-wise reaction and its Jacobian on the device.  ``num=`` may be given
instead of ``dx=`` to avoid the reference's floating-point floor in ``from_bbox_1d``.
"""
import functools

import numpy as np

from .. import diffops, kernels, mesh
from . import problems


def _mesh(bbox, dx, num):
    if num is not None:
        return mesh.RectangularMesh.from_bbox_1d(bbox, num=num)
    return mesh.RectangularMesh.from_bbox_1d(bbox, step=dx)


def _bbox(bbox):
    return np.asarray([0.0, 1.0] if bbox is None else bbox, dtype=np.float64)


def heat_1d(*, bbox=None, t0=0.0, tmax=5.0, y0_fun=None, diffusion_rate=0.05, bcond="dirichlet"):
    """examples.py:50-81."""
    bbox = _bbox(bbox)
    if y0_fun is None:
        y0_fun = lambda x: gaussian_bell_1d_centered(x, bbox) * sin_bell_1d(x)  # noqa: E731
    cls = {"dirichlet": problems.LinearEvolutionDirichlet, "neumann": problems.LinearEvolutionNeumann}.get(bcond)
    if cls is None:
        raise ValueError
    return cls(diffop=diffops.laplace(), diffop_scale=diffusion_rate, bbox=bbox, t0=t0, tmax=tmax, y0_fun=y0_fun)


def heat_1d_discretized(*, bbox=None, dx=0.05, num=None, stencil_size_interior=3, stencil_size_boundary=3, t0=0.0,
                        tmax=5.0, y0_fun=None, diffusion_rate=0.05, nugget_gram_matrix_fd=0.0, kernel=None,
                        bcond="dirichlet"):
    """examples.py:13-47."""
    heat = heat_1d(bbox=bbox, t0=t0, tmax=tmax, y0_fun=y0_fun, diffusion_rate=diffusion_rate, bcond=bcond)
    heat.discretize(mesh_spatial=_mesh(heat.bbox, dx, num), kernel=kernel or kernels.SquareExponential(),
                    stencil_size_interior=stencil_size_interior, stencil_size_boundary=stencil_size_boundary,
                    nugget_gram_matrix=nugget_gram_matrix_fd)
    return heat


def sir_1d(*, bbox=None, t0=0.0, tmax=50.0, diffusion_rate_S=0.1, diffusion_rate_I=0.1, diffusion_rate_R=0.1,
           beta=0.3, gamma=0.07, N=1000.0):
    """examples.py:127-178."""
    bbox = _bbox(bbox)

    def y0_fun(x):
        infectious = 200.0 * gaussian_bell_1d_centered(x, bbox, width=0.5) + 1.0
        return np.concatenate((N * np.ones_like(infectious) - infectious, infectious, np.zeros_like(infectious)))

    def f(t, x):
        s, i, r = np.split(x, 3)
        force = beta * s * i / (s + i + r)
        return np.concatenate((-force, force - gamma * i, gamma * i))

    def df(t, x):
        s, i, r = np.split(x, 3)
        tot = s + i + r
        force = beta * s * i / tot
        ds, di, dr = beta * i / tot - force / tot, beta * s / tot - force / tot, -force / tot
        zero, dg = np.zeros_like(s), np.diag
        return np.block([[dg(-ds), dg(-di), dg(-dr)], [dg(ds), dg(di - gamma), dg(dr)],
                         [dg(zero), dg(zero + gamma), dg(zero)]])

    lap = diffops.laplace()
    return problems.SystemSemiLinearEvolutionNeumann(
        diffop=(lap, lap, lap), diffop_scale=(diffusion_rate_S, diffusion_rate_I, diffusion_rate_R), bbox=bbox, t0=t0,
        tmax=tmax, y0_fun=y0_fun, f=f, df=df, df_diagonal=None,
        reaction=problems.Reaction(problems.REACTION_IDS["sir"], (beta, gamma), 3))


def sir_1d_discretized(*, bbox=None, dx=0.05, num=None, t0=0.0, tmax=50.0, beta=0.3, gamma=0.07, N=1000.0,
                       diffusion_rate_S=0.1, diffusion_rate_I=0.1, diffusion_rate_R=0.1, kernel=None,
                       nugget_gram_matrix_fd=0.0, stencil_size_interior=3, stencil_size_boundary=3):
    """examples.py:84-124."""
    sir = sir_1d(bbox=bbox, t0=t0, tmax=tmax, diffusion_rate_S=diffusion_rate_S, diffusion_rate_I=diffusion_rate_I,
                 diffusion_rate_R=diffusion_rate_R, beta=beta, gamma=gamma, N=N)
    sir.discretize_system(mesh_spatial=_mesh(sir.bbox, dx, num), kernel=kernel or kernels.SquareExponential(),
                          stencil_size_interior=stencil_size_interior, stencil_size_boundary=stencil_size_boundary,
                          nugget_gram_matrix=nugget_gram_matrix_fd)
    return sir


def lotka_volterra_1d(*, bbox=None, t0=0.0, tmax=10.0, a=0.5, b=0.05, c=0.05, d=0.5, diffusion_scale_u=0.1,
                      diffusion_scale_v=0.1):
    """examples.py:206-248."""
    bbox = _bbox(bbox)

    def y0_fun(x):
        return np.concatenate((5 * np.ones_like(x), 20.0 * gaussian_bell_1d(x)))

    def f(_, x):
        u, v = np.split(x, 2)
        return np.concatenate((a * u - b * u * v, c * u * v - d * v))

    def df(_, x):
        u, v = np.split(x, 2)
        dg = np.diag
        return np.block([[dg(a - b * v), dg(-b * u)], [dg(c * v), dg(c * u - d)]])

    lap = diffops.laplace()
    return problems.SystemSemiLinearEvolutionNeumann(
        diffop=(lap, lap), diffop_scale=(diffusion_scale_u, diffusion_scale_v), bbox=bbox, t0=t0, tmax=tmax,
        y0_fun=y0_fun, f=f, df=df, df_diagonal=None,
        reaction=problems.Reaction(problems.REACTION_IDS["lotka_volterra"], (a, b, c, d), 2))


def lotka_volterra_1d_discretized(*, dx=0.05, num=None, kernel=None, nugget_gram_matrix_fd=0.0, stencil_size_interior=3,
                                  stencil_size_boundary=3, **kwargs):
    """examples.py:181-203."""
    pde = lotka_volterra_1d(**kwargs)
    pde.discretize_system(mesh_spatial=_mesh(pde.bbox, dx, num), kernel=kernel or kernels.SquareExponential(),
                          stencil_size_interior=stencil_size_interior, stencil_size_boundary=stencil_size_boundary,
                          nugget_gram_matrix=nugget_gram_matrix_fd)
    return pde


def spruce_budworm_1d(*, bbox=None, t0=0.0, tmax=10.0, diffusion_rate=0.1, y0_fun=None, bcond="dirichlet",
                      growth_rate=1.0):
    """examples.py:290-341 (Fisher's equation)."""
    bbox = _bbox(bbox)
    g = growth_rate
    cls = {"dirichlet": problems.SemiLinearEvolutionDirichlet, "neumann": problems.SemiLinearEvolutionNeumann}.get(bcond)
    if cls is None:
        raise ValueError
    return cls(t0=t0, tmax=tmax, y0_fun=y0_fun or sin_bell_1d, bbox=bbox, diffop=diffops.laplace(),
               diffop_scale=diffusion_rate, f=lambda _, x: g * x * (1.0 - x),
               df=lambda _, x: np.diag(g * (1.0 - 2.0 * x)), df_diagonal=None,
               reaction=problems.Reaction(problems.REACTION_IDS["spruce"], (g,), 1))


def spruce_budworm_1d_discretized(*, bbox=None, t0=0.0, tmax=10.0, diffusion_rate=1.0, y0_fun=None, dx=0.1, num=None,
                                  kernel=None, nugget_gram_matrix_fd=0.0, stencil_size_interior=3,
                                  stencil_size_boundary=3, bcond="dirichlet", growth_rate=1.0):
    """examples.py:251-287."""
    spruce = spruce_budworm_1d(bbox=bbox, t0=t0, tmax=tmax, diffusion_rate=diffusion_rate, y0_fun=y0_fun, bcond=bcond,
                               growth_rate=growth_rate)
    spruce.discretize(mesh_spatial=_mesh(spruce.bbox, dx, num), kernel=kernel or kernels.SquareExponential(),
                      stencil_size_interior=stencil_size_interior, stencil_size_boundary=stencil_size_boundary,
                      nugget_gram_matrix=nugget_gram_matrix_fd)
    return spruce


# initial-condition defaults, examples.py:347-357
def gaussian_bell_1d_centered(x, bbox, width=1.0):
    return np.exp(-((x - 0.5 * (bbox[1] + bbox[0])) ** 2) / width ** 2)


def gaussian_bell_1d(x):
    return np.exp(-(x ** 2))


def sin_bell_1d(x):
    return 0.1 * np.sin(np.pi * x)
