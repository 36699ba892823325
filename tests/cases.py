"""Shared problem/solver cases for the tests (product objects + matching oracle objects)."""
import numpy as np

from oracle import ek1_np, setup_np
from pnmol_b200 import kernels, latent, white
from pnmol_b200.odetools import step
from pnmol_b200.pde import examples

SOLVERS = {
    "white_linear": white.LinearWhiteNoiseEK1,
    "white_semilinear": white.SemiLinearWhiteNoiseEK1,
    "latent_linear": latent.LinearLatentForceEK1,
    "latent_semilinear": latent.SemiLinearLatentForceEK1,
}


def make_case(name, *, num=6, bcond="dirichlet", tmax=1.0, nu=2, dt=2.0 ** -4, prior="se"):
    """name: heat | spruce | sir | lv.  Returns dict(pde, opde, kernel, okernel, copies, nu, dt)."""
    se_prod = kernels.SquareExponential() + kernels.WhiteNoise()
    se_orac = setup_np.Sum(setup_np.SE(), setup_np.White())
    mat_prod = kernels.Matern52() + kernels.WhiteNoise()
    mat_orac = setup_np.Sum(setup_np.Matern52(), setup_np.White())
    kp, ko = (se_prod, se_orac) if prior == "se" else (mat_prod, mat_orac)
    copies = 1
    if name == "heat":
        pde = examples.heat_1d_discretized(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
        opde = setup_np.heat_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
    elif name == "spruce":
        pde = examples.spruce_budworm_1d_discretized(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
        opde = setup_np.spruce_budworm_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
    elif name == "sir":
        pde = examples.sir_1d_discretized(num=num, tmax=tmax, diffusion_rate_S=0.035, diffusion_rate_I=0.035,
                                          diffusion_rate_R=0.035, stencil_size_boundary=min(5, num))
        opde = setup_np.sir_1d(num=num, tmax=tmax, diffusion_rates=(0.035,) * 3, n_bnd=min(5, num))
        copies = 3
    elif name == "lv":
        pde = examples.lotka_volterra_1d_discretized(num=num, tmax=tmax)
        opde = setup_np.lotka_volterra_1d(num=num, tmax=tmax)
        copies = 2
    else:
        raise KeyError(name)
    kernel = kernels.duplicate(kp, copies) if copies > 1 else kp
    gram_sqrtm = np.linalg.cholesky(setup_np.gram(ko, opde.points, copies))
    return dict(pde=pde, opde=opde, kernel=kernel, gram_sqrtm=gram_sqrtm, nu=nu, dt=dt, copies=copies)


def make_solver(kind, case, **kw):
    return SOLVERS[kind](num_derivatives=case["nu"], steprule=step.Constant(case["dt"]), spatial_kernel=case["kernel"], **kw)


def block_rel(a, b, n):
    """max |a - b| / max |b| per derivative block; a, b are (D,) means in
    (n, dd) layout or (D, D) covariances with flat index j n + i."""
    a, b = np.asarray(a), np.asarray(b)
    worst = 0.0
    if a.ndim == 2 and a.shape[0] == a.shape[1]:
        for i in range(n):
            for j in range(n):
                ref = np.max(np.abs(b[i::n, j::n]))
                worst = max(worst, np.max(np.abs(a[i::n, j::n] - b[i::n, j::n])) / max(ref, 1e-300))
        return worst
    for i in range(a.shape[0]):
        worst = max(worst, np.max(np.abs(a[i] - b[i])) / max(np.max(np.abs(b[i])), 1e-300))
    return worst


def cov(L):
    L = np.asarray(L)
    return L @ L.T


def cov_excess(La, Lb, n, rtol=None):
    """Covariance parity as L L^T: max over entries of |dP_ij| / tol_ij (parity holds iff < 1) with

        tol_ij = rtol * max|P_block(i,j)| + c * eps * s * (sigma_i + sigma_j),

    sigma_i = sqrt(P_ii), s = max sigma, c = 10 D.  The first term is the north-star tolerance in the
    per-derivative-block norm.  The second is the backward-error floor of ANY
    orthogonal factorisation (|dL| <~ eps * s  =>  |dP_ij| <~ eps * s * (sigma_i + sigma_j)): it only matters for
    blocks whose magnitude is below eps * cond, e.g. the derivative-0 block right after conditioning on y0 with
    the reference's 1e-10 nugget (variances ~1e-20 next to O(1) ones), which LAPACK itself reproduces to ~1e-6."""
    rtol = COV_RTOL if rtol is None else rtol
    Pa, Pb = cov(La), cov(Lb)
    D = Pb.shape[0]
    sig = np.sqrt(np.abs(np.diag(Pb)))
    floor = 10.0 * D * np.finfo(np.float64).eps * sig.max() * (sig[:, None] + sig[None, :])
    tol = np.empty_like(Pb)
    for i in range(n):
        for j in range(n):
            tol[i::n, j::n] = rtol * np.max(np.abs(Pb[i::n, j::n]))
    return float(np.max(np.abs(Pa - Pb) / np.maximum(tol + floor, 1e-300)))


import contextlib


@contextlib.contextmanager
def perturbed_oracle(seed=0):
    """Run the oracle with eps * max|A| * N(0,1) added to every QR input A (the normwise backward error of any
    Householder QR, LAPACK's included): the spread between this and the exact oracle measures how reproducible
    the reference's own result is (its conditioning floor)."""
    from oracle import sqrt_np

    orig = sqrt_np.triu_factor
    rng = np.random.default_rng(seed)

    def noisy(stack):
        stack = np.asarray(stack, dtype=np.float64)
        return orig(stack + np.finfo(np.float64).eps * np.max(np.abs(stack)) * rng.standard_normal(stack.shape))

    sqrt_np.triu_factor = noisy
    try:
        yield
    finally:
        sqrt_np.triu_factor = orig


def oracle_pair(fn):
    """(exact oracle result, result under eps-level normwise perturbations of every QR input)."""
    ref = fn()
    with perturbed_oracle():
        ref_eps = fn()
    return ref, ref_eps


def mean_excess(a, b, rtol=None, spread=None):
    """Mean parity: max over derivative rows of |dm| / (rtol * max|row| + 10 D eps max|m|) (holds iff < 1).  The
    second term only matters for rows that are pure rounding noise (e.g. the second-derivative row of the initial
    mean, ~1e-18 next to O(0.1) rows)."""
    rtol = MEAN_RTOL if rtol is None else rtol
    a, b = np.asarray(a), np.asarray(b)
    floor = 10.0 * b.size * np.finfo(np.float64).eps * np.max(np.abs(b))
    # spread: a second oracle result from eps-perturbed inputs; 10 x its distance is the reproducibility floor
    extra = [0.0] * b.shape[0] if spread is None else [10.0 * np.max(np.abs(np.asarray(spread)[i] - b[i])) for i in range(b.shape[0])]
    return float(max(np.max(np.abs(a[i] - b[i])) / max(rtol * np.max(np.abs(b[i])) + floor + extra[i], 1e-300)
                     for i in range(b.shape[0])))


MEAN_RTOL = 1e-9  # BASELINE.json north_star: filter means to rtol 1e-9
COV_RTOL = 1e-8   # covariances compared as L L^T to rtol 1e-8
