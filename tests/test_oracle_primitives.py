"""Pin the oracle to the reference's own known-answer tests for the hot path
(tests/test_base/test_sqrt.py:37-109, tests/test_base/test_iwp.py:20-114,
tests/test_discretize.py:64-71, tests/test_pdefilter.py:141-146,
tests/test_odetools/test_step.py:28-45 of the reference)."""
import numpy as np
import pytest

from oracle import ek1_np, prior_np, setup_np, sqrt_np

import cases


def _iwp1():
    return prior_np.iwp_1d(1)  # (H, SQ) of the reference's sqrt fixtures (test_sqrt.py:10-34)


@pytest.mark.parametrize("style", ["full", "partial"])
def test_propagate_cholesky_factor(style):
    H, SQ = _iwp1()
    SC = SQ
    if style == "partial":
        H, SQ = H[:1], SQ[:1, :1]
    chol = sqrt_np.chol_of_sum(H @ SC, SQ)
    assert np.allclose(chol @ chol.T, H @ SC @ SC.T @ H.T + SQ @ SQ.T)
    assert np.allclose(np.tril(chol), chol)


@pytest.mark.parametrize("style", ["full", "partial"])
@pytest.mark.parametrize("noise", [True, False])
def test_update_sqrt(style, noise):
    H, SQ = _iwp1()
    SC = SQ
    if style == "partial":
        H, SQ = H[:1], SQ[:1, :1]
    C_new, K, S_chol = sqrt_np.measurement_update(H, SC, SQ if noise else None)
    assert C_new.shape == SC.shape and K.shape == (H.shape[1], H.shape[0]) and S_chol.shape == (H.shape[0],) * 2
    S = H @ SC @ SC.T @ H.T + (SQ @ SQ.T if noise else 0.0)
    Kx = SC @ SC.T @ H.T @ np.linalg.inv(S)
    assert np.allclose(C_new @ C_new.T, SC @ SC.T - Kx @ S @ Kx.T)
    assert np.allclose(C_new, np.tril(C_new))
    assert np.allclose(K, Kx)
    assert np.allclose(S_chol @ S_chol.T, S) and np.allclose(S_chol, np.tril(S_chol))


def test_iwp_closed_forms():
    dt = 0.1
    A, LQ = prior_np.non_preconditioned(2, np.eye(1), dt)
    assert np.allclose(A, [[1.0, dt, dt ** 2 / 2], [0, 1.0, dt], [0, 0, 1.0]])
    assert np.allclose(LQ @ LQ.T, [[dt ** 5 / 20, dt ** 4 / 8, dt ** 3 / 6], [dt ** 4 / 8, dt ** 3 / 3, dt ** 2 / 2],
                                   [dt ** 3 / 6, dt ** 2 / 2, dt]])
    p, pinv = prior_np.nordsieck_scales(2, dt)
    assert np.allclose(p * pinv, 1.0)
    P, Pinv = prior_np.nordsieck_dense(2, 3, dt)
    assert np.allclose(P @ Pinv, np.eye(9))
    E0 = prior_np.projection(2, 1, 0)
    assert E0.shape == (1, 3) and (E0 == 1).sum() == 1


def test_fd_weights_polynomial_kernel():
    """tests/test_discretize.py:64-71: polynomial kernel recovers [-2, 1, 1] / dx^2, zero uncertainty."""
    pts = setup_np.mesh_1d([0.0, 1.0], step=0.1)
    dx = 0.1
    w, unc = setup_np.stencil_weights(setup_np.Poly(order=2, const=1.0), pts[1, 0], pts[[1, 0, 2], 0], "laplace")
    assert np.allclose(w * dx ** 2, [-2.0, 1.0, 1.0])
    assert np.allclose(unc, 0.0, atol=1e-8)


def test_mesh_floor_quirk():
    assert setup_np.mesh_1d([0.0, 1.0], step=1 / 99).shape[0] == 99  # quirk Q6 (mesh.py:94)
    assert setup_np.mesh_1d([0.0, 1.0], num=100).shape[0] == 100


def test_kernel_derivatives_match_autograd():
    torch = pytest.importorskip("torch")
    x = torch.tensor(0.31, dtype=torch.float64, requires_grad=True)
    y = torch.tensor(0.47, dtype=torch.float64, requires_grad=True)

    def derivs(kfun):
        k = kfun(x, y)
        (kx,) = torch.autograd.grad(k, x, create_graph=True)
        (kxx,) = torch.autograd.grad(kx, x, create_graph=True)
        (kxy,) = torch.autograd.grad(kx, y, create_graph=True)
        (kxxy,) = torch.autograd.grad(kxx, y, create_graph=True)
        (kxxyy,) = torch.autograd.grad(kxxy, y, create_graph=True)
        return [float(v) for v in (kx, kxx, kxy, kxxyy)]

    r, s = 1.3, 0.8
    se = setup_np.SE(r, s)
    ref = derivs(lambda a, b: s ** 2 * torch.exp(-r ** 2 * (a - b) ** 2 / 2))
    got = [se.dx(0.31, 0.47), se.dxx(0.31, 0.47), se.dxy(0.31, 0.47), se.dxxyy(0.31, 0.47)]
    assert np.allclose(got, ref, rtol=1e-12)
    mat = setup_np.Matern52(r, s)

    def matern(a, b):
        dist = torch.sqrt(5.0 * (a - b) ** 2 * r ** 2)
        return s ** 2 * (1 + dist + dist ** 2 / 3.0) * torch.exp(-dist)

    ref = derivs(matern)
    got = [float(mat.dx(0.31, 0.47)), float(mat.dxx(0.31, 0.47)), float(mat.dxy(0.31, 0.47)), float(mat.dxxyy(0.31, 0.47))]
    assert np.allclose(got, ref, rtol=1e-10)
    pol = setup_np.Poly(order=3, const=0.7)
    ref = derivs(lambda a, b: (a * b + 0.7) ** 3)
    got = [pol.dx(0.31, 0.47), pol.dxx(0.31, 0.47), pol.dxy(0.31, 0.47), pol.dxxyy(0.31, 0.47)]
    assert np.allclose(got, ref, rtol=1e-12)
    # the Taylor constants the reference substitutes for Matern NaNs (discretize.py:184-197)
    eps = 1e-4
    assert np.isclose(mat.nan_first, float(mat.dxx(0.0, eps)), rtol=1e-3)
    assert np.isclose(mat.nan_second, float(mat.dxxyy(0.0, eps)), rtol=1e-3)


@pytest.mark.parametrize("name", ["spruce", "sir", "lv"])
def test_reaction_jacobians_by_finite_differences(name):
    case = cases.make_case(name, num=5, bcond="dirichlet" if name == "spruce" else "neumann")
    o = case["opde"]
    x = o.y0 + 0.1 * np.arange(o.y0.size) / o.y0.size + 0.2
    J = o.df(0.0, x)
    h = 1e-6
    for k in range(x.size):
        e = np.zeros_like(x); e[k] = h
        fd = (o.f(0.0, x + e) - o.f(0.0, x - e)) / (2 * h)
        assert np.allclose(J[:, k], fd, rtol=1e-6, atol=1e-7)
    # product-side callables agree with the oracle's
    p = case["pde"]
    assert np.allclose(p.f(0.0, x), o.f(0.0, x)) and np.allclose(p.df(0.0, x), J)


def test_constant_step_schedule_has_sliver():
    ts, dts = ek1_np.constant_step_schedule(0.0, 1.0, 0.1)  # tests/test_pdefilter.py config: 11 steps
    assert len(dts) == 11 and dts[-1] < 1e-15 and ts[-1] == 1.0
    ts, dts = ek1_np.constant_step_schedule(0.0, 3.0, 2.0 ** -4)
    assert len(dts) == 48 and np.all(dts == 2.0 ** -4)


@pytest.mark.parametrize("bcond", ["dirichlet", "neumann"])
@pytest.mark.parametrize("kind", ["white_linear", "latent_linear", "white_semilinear", "latent_semilinear"])
def test_solve_no_nan(kind, bcond):
    """tests/test_pdefilter.py:141-146 of the reference (dx=0.2, dt=0.1, tmax=1, nu=2)."""
    case = cases.make_case("spruce" if "semi" in kind else "heat", num=6, bcond=bcond, dt=0.1)
    sol = ek1_np.solve(kind, case["opde"], 0.1, 2, case["gram_sqrtm"])
    assert sol.t.shape == (12,)
    assert not np.isnan(sol.mean).any() and not np.isnan(sol.cov_sqrtm).any()


def test_initialisation_reproduces_initial_condition():
    case = cases.make_case("heat", num=10)
    st = ek1_np.white_initialize(case["opde"], 2, case["gram_sqrtm"])
    assert np.allclose(st.mean[0], case["opde"].y0, atol=1e-9)
