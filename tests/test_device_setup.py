"""Batched problem set-up on the device: probabilistic FD stencils and spatial-Gram Cholesky
factors, checked against the oracle's literal restatement (oracle/setup_np.py) and the reference's known answers
(/root/reference/tests/test_discretize.py:30-71)."""
import numpy as np
import pytest
import torch

from oracle import setup_np

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps


def _oracle_kernel(k):
    from pnmol_b200 import kernels

    if isinstance(k, kernels.SquareExponential):
        return setup_np.SE(k.input_scale, k.output_scale)
    if isinstance(k, kernels.Matern52):
        return setup_np.Matern52(k.input_scale, k.output_scale)
    return setup_np.Poly(order=k.order, const=k.const)


def _kernel_lists():
    from pnmol_b200 import kernels

    return {
        "se": [kernels.SquareExponential(input_scale=r, output_scale=s) for r, s in ((1.0, 1.0), (2.5, 0.7), (7.0, 1.3), (0.4, 2.0))],
        "matern52": [kernels.Matern52(input_scale=r, output_scale=s) for r, s in ((1.0, 1.0), (3.0, 0.5), (0.3, 2.0))],
        "polynomial": [kernels.Polynomial(order=o, const=c) for o, c in ((2, 1.0), (3, 1.0), (4, 0.5))],
    }


@pytest.mark.parametrize("family", ["se", "matern52", "polynomial"])
@pytest.mark.parametrize("mode,stencil", [("laplace", 3), ("laplace", 5), ("gradient", 2)])
def test_fd_coefficients_batched_match_the_oracle(family, mode, stencil):
    """Every (kernel, point) stencil system solved by one device thread equals the oracle's numpy.linalg.solve to the
    accuracy the Gram matrix's condition number allows (both are LU with partial pivoting)."""
    from pnmol_b200 import diffops, discretize

    klist = _kernel_lists()[family]
    pts = setup_np.mesh_1d((0.0, 1.0), num=21)
    x = pts[:, 0]
    order = np.argsort(np.abs(x[:, None] - x[None, :]), axis=1, kind="stable")[:, :stencil]
    nbrs = x[order]
    diffop = diffops.laplace() if mode == "laplace" else diffops.gradient()
    w, u = discretize.fd_coefficients_batched(x, nbrs, klist, diffop)
    w, u = w.cpu().numpy(), u.cpu().numpy()
    assert w.shape == (len(klist), len(x), stencil) and u.shape == (len(klist), len(x))
    for b, k in enumerate(klist):
        ok = _oracle_kernel(k)
        for p in range(len(x)):
            X = nbrs[p]
            cond = np.linalg.cond(ok.k(X[:, None], X[None, :]))
            if not np.isfinite(cond) or cond > 1e13:
                continue  # numerically singular stencil (polynomial kernel with more points than monomials)
            wr, ur = setup_np.stencil_weights(ok, x[p], nbrs[p], mode)
            tol = max(1e-10, 200 * EPS * cond)
            np.testing.assert_allclose(w[b, p], wr, rtol=tol, atol=tol * np.abs(wr).max())
            # the variance top - w.rhs is a difference of O(|w|.|rhs|) terms
            rhs = ok.k(X[:, None], X[None, :]) @ wr
            assert abs(u[b, p] - ur) <= tol * (abs(ur) + np.abs(wr) @ np.abs(rhs)) + 1e-300, (b, p, u[b, p], ur)


def test_polynomial_kernel_known_answer():
    """/root/reference/tests/test_discretize.py:30-71: with a polynomial kernel the stencil (x1; x1, x0, x2) reproduces the
    classical second-difference coefficients (-2, 1, 1) / dx^2 with zero uncertainty."""
    from pnmol_b200 import diffops, discretize, kernels

    dx = 0.1
    x = np.arange(0.0, 1.0 + 1e-12, dx)
    w, u = discretize.fd_coefficients_batched(x[[1]], x[[1, 0, 2]][None, :], [kernels.Polynomial(const=1.0)], diffops.laplace())
    np.testing.assert_allclose(w.cpu().numpy()[0, 0] * dx ** 2, [-2.0, 1.0, 1.0], atol=1e-8)
    np.testing.assert_allclose(u.cpu().numpy()[0, 0], 0.0, atol=1e-8)


def test_fd_probabilistic_batched_equals_the_host_discretisation():
    """discretize.py:12-113 over a sweep of input scales: L and E_sqrtm of every member equal the host (NumPy) set-up
    that the filter tests use."""
    from pnmol_b200 import diffops, discretize, kernels, mesh

    m = mesh.RectangularMesh.from_bbox_1d([0.0, 1.0], num=30)
    klist = [kernels.SquareExponential(input_scale=r) for r in (0.5, 1.0, 2.0, 4.0, 8.0)]
    L, E = discretize.fd_probabilistic_batched(diffops.laplace(), m, klist, stencil_size_interior=3, stencil_size_boundary=3)
    L, E = L.cpu().numpy(), E.cpu().numpy()
    for b, k in enumerate(klist):
        Lh, Eh = discretize.fd_probabilistic(diffops.laplace(), m, k, 3, 3)
        assert np.array_equal(L[b] != 0.0, Lh != 0.0)
        np.testing.assert_allclose(L[b], Lh, rtol=1e-6, atol=1e-6 * np.abs(Lh).max())   # cond(Gram) ~ 1e9 at dx = 1/29
        np.testing.assert_allclose(np.diag(E[b]), np.diag(Eh), rtol=1e-5, atol=1e-5 * np.abs(np.diag(Eh)).max() + 1e-9)


@pytest.mark.parametrize("d", [12, 50, 200])
def test_gram_cholesky_batched(d):
    """white.py:82-94: cholesky(spatial_kernel(X, X.T)) for a sweep of kernels; shared-memory (d <= 168) and in-place
    (larger d) factorisations."""
    from pnmol_b200 import kernels

    x = np.linspace(0.0, 1.0, d)
    for klist in ([kernels.SquareExponential(input_scale=r) + kernels.WhiteNoise(output_scale=0.1) for r in (1.0, 3.0, 10.0)],
                  [kernels.Matern52(input_scale=r, output_scale=s) + kernels.WhiteNoise(output_scale=0.05)
                   for r, s in ((1.0, 1.0), (4.0, 0.5))]):
        chol, status, _ = kernels.gram_cholesky_batched(klist, x[:, None])
        chol = chol.cpu().numpy()
        assert int(status.max()) == 0
        for b, k in enumerate(klist):
            K = k(x[:, None], x[None, :])
            ref = np.linalg.cholesky(K)
            assert np.all(np.triu(chol[b], 1) == 0.0)
            np.testing.assert_allclose(chol[b] @ chol[b].T, K, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(chol[b], ref, rtol=1e-8, atol=1e-10)


def test_not_positive_definite_is_reported():
    from pnmol_b200 import kernels

    x = np.linspace(0.0, 1.0, 40)
    _, status, _ = kernels.gram_cholesky_batched([kernels.SquareExponential(input_scale=0.1)], x[:, None])
    assert int(status[0]) == 1   # numerically singular without a nugget: NaN factor in the reference too


def test_mle_input_scale_matches_the_reference_formula():
    """kernels.py:186-211: log-likelihood of every trial input scale in one launch; argmax equals the host evaluation of
    the reference's formula (solve + log det)."""
    from pnmol_b200 import kernels

    rng = np.random.default_rng(5)
    x = np.linspace(0.0, 1.0, 25)
    trials = np.array([4.0, 6.0, 9.0, 14.0, 20.0])
    K_true = kernels.Matern52(input_scale=9.0)(x[:, None], x[None, :])
    data = np.linalg.cholesky(K_true + 1e-10 * np.eye(25)) @ rng.standard_normal(25)
    _, _, ll = kernels.gram_cholesky_batched([kernels.Matern52(input_scale=float(s)) for s in trials], x[:, None], data=data)
    ll = ll.cpu().numpy()
    ref = []
    for s in trials:
        K = kernels.Matern52(input_scale=float(s))(x[:, None], x[None, :])
        a = data @ np.linalg.solve(K, data)
        b = np.log(np.linalg.det(K))
        ref.append(-0.5 * (a + b + 25 * np.log(2 * np.pi)))
    np.testing.assert_allclose(ll, ref, rtol=1e-7)
    best = kernels.mle_input_scale(mesh_points=x[:, None], data=data, kernel_type=kernels.Matern52, input_scale_trials=trials)
    assert best == trials[int(np.argmax(ref))]
