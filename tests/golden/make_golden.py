"""Generate tests/golden/*.npz from the NumPy oracle (NOT from the reference, which cannot
be imported here: no jax / tornadox in the image -- see oracle/__init__.py).

    python tests/golden/make_golden.py

Each file freezes the inputs of one small configuration (L, E_sqrtm, B, R_sqrtm, y0,
chol of the spatial Gram matrix) and the oracle's outputs (initial state, every step's mean
and factor, local diffusions) on an exactly representable time grid.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pnmol-experiments_b200"), os.path.dirname(HERE)]

from oracle import ek1_np  # noqa: E402

import cases  # noqa: E402

CONFIGS = {
    "heat_dirichlet_white_linear": ("heat", "white_linear", dict(num=6, bcond="dirichlet")),
    "heat_neumann_white_linear": ("heat", "white_linear", dict(num=6, bcond="neumann")),
    "heat_dirichlet_latent_linear": ("heat", "latent_linear", dict(num=6, bcond="dirichlet")),
    "spruce_dirichlet_white_semilinear": ("spruce", "white_semilinear", dict(num=6, bcond="dirichlet")),
    "spruce_dirichlet_latent_semilinear": ("spruce", "latent_semilinear", dict(num=6, bcond="dirichlet")),
    "sir_neumann_white_semilinear": ("sir", "white_semilinear", dict(num=5)),
}

if __name__ == "__main__":
    for name, (prob, kind, kw) in CONFIGS.items():
        case = cases.make_case(prob, tmax=0.5, **kw)
        o = case["opde"]
        sol = ek1_np.solve(kind, o, case["dt"], case["nu"], case["gram_sqrtm"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), L=o.L, E_sqrtm=o.E_sqrtm, B=o.B, R_sqrtm=o.R_sqrtm,
                            y0=o.y0, gram_sqrtm=case["gram_sqrtm"], dt=case["dt"], nu=case["nu"], t=sol.t, mean=sol.mean,
                            cov_sqrtm=sol.cov_sqrtm, diffusion_squared_calibrated=sol.diffusion_squared_calibrated,
                            kind=kind, problem=prob)
        print(name, sol.t.shape, sol.mean.shape, float(sol.diffusion_squared_calibrated))
