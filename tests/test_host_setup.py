"""Host-side (NumPy) set-up of the product package: same results as the oracle's literal
restatement, same API semantics as the reference modules."""
import numpy as np
import pytest

from oracle import setup_np
from pnmol_b200 import _engine, discretize, diffops, kernels, mesh
from pnmol_b200.base import iwp, stacked_ssm
from pnmol_b200.odetools import step

import cases


@pytest.mark.parametrize("name,bcond", [("heat", "dirichlet"), ("heat", "neumann"), ("spruce", "dirichlet"),
                                        ("sir", "neumann"), ("lv", "neumann")])
def test_discretisation_matches_oracle(name, bcond):
    case = cases.make_case(name, num=9, bcond=bcond)
    for attr in ("L", "E_sqrtm", "B", "R_sqrtm", "y0"):
        a, b = getattr(case["pde"], attr), getattr(case["opde"], attr)
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13), attr
    X = case["pde"].mesh_spatial.points
    G = case["kernel"](X, X.T)
    assert np.allclose(np.linalg.cholesky(G), case["gram_sqrtm"], rtol=1e-13, atol=1e-15)


def test_matern_fd_discretisation_matches_oracle():
    pts = setup_np.mesh_1d([0.0, 1.0], num=7)
    Lo, Eo = setup_np.fd_laplace(pts, setup_np.Matern52(1.0, 1.0), 3, 4)
    msh = mesh.RectangularMesh.from_bbox_1d([0.0, 1.0], num=7)
    L, E = discretize.fd_probabilistic(diffops.laplace(), msh, kernels.Matern52(), 3, 4)
    assert np.allclose(L, Lo, rtol=1e-12) and np.allclose(E, Eo, rtol=1e-10, atol=1e-14)


def test_fd_coefficients_polynomial():
    msh = mesh.RectangularMesh.from_bbox_1d([0.0, 1.0], step=0.1)
    w, unc = discretize.fd_coefficients(msh[1], msh[((1, 0, 2),)], kernels.Polynomial(const=1.0))
    assert np.allclose(w * 0.1 ** 2, [-2.0, 1.0, 1.0]) and np.isclose(unc, 0.0, atol=1e-8)


def test_kernel_call_conventions():
    k = kernels.SquareExponential(input_scale=2.0, output_scale=0.5)
    X = np.linspace(0, 1, 4).reshape(-1, 1)
    G = k(X, X.T)
    assert G.shape == (4, 4) and np.allclose(np.diag(G), 0.25)
    assert k(X, X).shape == (4,)
    assert np.isclose(k(X[0], X[1]), G[0, 1])
    Gw = (k + kernels.WhiteNoise(output_scale=0.1))(X, X.T)
    assert np.allclose(Gw - G, 0.01 * np.eye(4))
    Gd = kernels.duplicate(k, 3)(X, X.T)
    assert Gd.shape == (12, 12) and np.allclose(Gd[4:8, 4:8], G) and np.all(Gd[:4, 4:] == 0)


def test_mesh_api():
    msh = mesh.RectangularMesh.from_bbox_1d([0.0, 1.0], step=1 / 99)
    assert len(msh) == 99  # quirk Q6
    msh = mesh.RectangularMesh.from_bbox_1d([0.0, 1.0], num=11)
    assert msh.shape == (11, 1) and msh.boundary[2].tolist() == [0, 10] and len(msh.interior[2]) == 9
    assert msh.boundary_projection_matrix.shape == (2, 11)
    nb, idx = msh.neighbours(msh.points[[5]], 3)
    assert sorted(idx[0].tolist()) == [4, 5, 6]
    with pytest.raises(ValueError):
        mesh.RectangularMesh.from_bbox_1d([0.0, 1.0])


def test_iwp_api_reference_tests():
    """tests/test_base/test_iwp.py of the reference, on the product's IWP."""
    dt = 0.1
    p = iwp.IntegratedWienerTransition(wiener_process_dimension=1, num_derivatives=2, wp_diffusion_sqrtm=np.eye(1))
    A, LQ = p.non_preconditioned_discretize(dt)
    assert np.allclose(A, [[1.0, dt, dt ** 2 / 2], [0, 1.0, dt], [0, 0, 1.0]])
    assert np.allclose((LQ @ LQ.T)[0], [dt ** 5 / 20, dt ** 4 / 8, dt ** 3 / 6])
    P, Pinv = p.nordsieck_preconditioner(dt)
    Ap, LQp = p.preconditioned_discretize
    assert np.allclose(P @ Ap @ Pinv, A) and np.allclose(P @ LQp, LQ)
    assert p.projection_matrix(0).shape == (1, 3) and p.state_dimension == 3
    raw, raw_inv = p.nordsieck_preconditioner_1d_raw(dt)
    assert np.allclose(raw * raw_inv, 1.0)
    assert np.array_equal(raw, _engine.nordsieck_raw(2, dt)[0])
    ssm = stacked_ssm.StackedSSM([p, p])
    assert ssm.state_dimension == 6 and ssm.projection_matrix(0).shape == (2, 6)
    assert ssm.projection_matrix(1, 1).shape == (1, 6) and ssm.projection_matrix(1, 1)[0, 4] == 1


def test_step_rules():
    """tests/test_odetools/test_step.py:15-122 semantics."""
    c = step.Constant(0.1)
    assert c.suggest(0.3, 12.0) == 0.1 and c.is_accepted(1e9) and c.scale_error_estimate(None, None) is None
    assert c.first_dt(None) == 0.1
    a = step.Adaptive(abstol=0.1, reltol=0.01)
    assert a.is_accepted(0.5) and not a.is_accepted(1.5)
    assert a.suggest(1.0, 1.0, local_convergence_rate=3) == pytest.approx(0.95)
    assert a.suggest(1.0, 1e-12, local_convergence_rate=1) == pytest.approx(10.0)
    assert a.suggest(1.0, 1e12, local_convergence_rate=1) == pytest.approx(0.2)
    with pytest.raises(ValueError):
        a.suggest(1.0, 1.0)
    err, ref = np.array([0.2, 0.2]), np.array([1.0, 3.0])
    expect = np.linalg.norm(err / (0.1 + 0.01 * ref)) / np.sqrt(2)
    assert a.scale_error_estimate(err, ref) == pytest.approx(expect)
    case = cases.make_case("heat", num=8)
    assert a.first_dt(case["pde"]) > 0


def test_ell_roundtrip_and_schedule():
    M = np.array([[0.0, 2.0, 0.0], [1.0, 0.0, 3.0]])
    col, val = _engine.to_ell(M)
    assert col.tolist() == [[1, -1], [0, 2]] and val.tolist() == [[2.0, 0.0], [1.0, 3.0]]
    dts = _engine.constant_step_schedule(0.0, 1.0, 0.1)
    assert len(dts) == 11 and dts[-1] < 1e-15  # sliver step of pdefilter.py:140,220-223
    assert len(_engine.constant_step_schedule(0.0, 3.0, 2.0 ** -4)) == 48


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pnmol_b200 import _lib

    case = cases.make_case("heat", num=6)
    solver = cases.make_solver("white_linear", case)
    with pytest.raises(_lib.PnmolB200Error):
        solver.solve(case["pde"])
    from pnmol_b200.base import sqrt

    with pytest.raises(_lib.PnmolB200Error):
        sqrt.propagate_cholesky_factor(np.eye(2), np.eye(2))
