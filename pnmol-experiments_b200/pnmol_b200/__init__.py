"""pnmol_b200: the EK1 filter loop of the probabilistic numerical method of lines on B200.

Keeps the Python solver / discretisation / problem API of
schmidtjonathan/pnmol-experiments (``src/pnmol``) for the white-noise and latent-force
EK1 solvers and runs the per-step square-root filter arithmetic in hand-written sm_100a
CUDA kernels behind the C ABI of ``include/pnmol_b200.h``.  Problem set-up
(mesh, kernels, finite-difference discretisation) is host-side NumPy, as it is host-side
JAX in the reference; everything from ``initialize`` on runs on the GPU.
"""
from . import base, diffops, discretize, ensemble, kernels, latent, mesh, odetools, pde, pdefilter, white  # noqa: F401

__all__ = ["base", "diffops", "discretize", "ensemble", "kernels", "latent", "mesh", "odetools", "pde", "pdefilter", "white"]
