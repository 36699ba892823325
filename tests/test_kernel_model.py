"""The algorithm implemented by the CUDA kernels (NumPy emulation in kernel_model.py, with
the row-support envelopes computed by libpnmol_b200's host code) equals the oracle."""
import numpy as np
import pytest

from oracle import ek1_np

import cases
import kernel_model


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__

    __graft_entry__.ensure_built()


CASES = [("heat", "white_linear", "dirichlet", 8), ("heat", "white_linear", "neumann", 8),
         ("heat", "latent_linear", "dirichlet", 7), ("spruce", "white_semilinear", "dirichlet", 8),
         ("spruce", "latent_semilinear", "dirichlet", 6), ("sir", "white_semilinear", "neumann", 5),
         ("sir", "latent_semilinear", "neumann", 4), ("lv", "white_semilinear", "neumann", 6)]


@pytest.mark.parametrize("blocked", [True, False], ids=["blocked", "unblocked"])
@pytest.mark.parametrize("name,kind,bcond,num", CASES + [("heat", "white_linear", "dirichlet", 23), ("heat", "latent_linear", "neumann", 12)])
def test_model_matches_oracle(name, kind, bcond, num, blocked):
    case = cases.make_case(name, num=num, bcond=bcond)
    family = kind.split("_")[0]
    mdl = kernel_model.Model(case["pde"], family, case["nu"], case["gram_sqrtm"], blocked=blocked)
    init, stepf, semil = ek1_np.KINDS[kind]
    st = init(case["opde"], case["nu"], case["gram_sqrtm"], 1.0, semil)
    m0, C0 = mdl.initialize(case["pde"].y0)
    n = case["nu"] + 1
    assert not np.isnan(m0).any() and not np.isnan(C0).any()  # no read outside the envelopes
    assert cases.cov_excess(C0, st.cov_sqrtm, n) < 1
    with cases.perturbed_oracle():  # reproducibility floor of the reference's own initial mean, see DESIGN.md
        st_eps = init(case["opde"], case["nu"], case["gram_sqrtm"], 1.0, semil)
    assert cases.mean_excess(m0, st.mean, spread=st_eps.mean) < 1
    mean, chol = st.mean, st.cov_sqrtm
    for _ in range(3):
        st = stepf(case["opde"], st, case["dt"], case["nu"], case["gram_sqrtm"], semil)
        mean, chol, err, diff = mdl.step(mean, chol, case["dt"])
        assert not np.isnan(mean).any() and not np.isnan(chol).any()
        assert np.allclose(np.tril(chol), chol)
        assert cases.mean_excess(mean, st.mean) < 1
        assert cases.cov_excess(chol, st.cov_sqrtm, n) < 1
        if err is not None:
            assert np.allclose(err, st.error_estimate, rtol=1e-8)


def test_model_step_from_dense_factor_and_first_step_diffusion():
    """A dense (non-triangular) input factor goes through the dense predict envelope; from identical
    inputs the QR sign convention (LAPACK dlarfg) reproduces the reference's quirk-Q1 diffusion."""
    case = cases.make_case("heat", num=7, bcond="neumann")
    mdl = kernel_model.Model(case["pde"], "white", 2, case["gram_sqrtm"])
    st = ek1_np.white_initialize(case["opde"], 2, case["gram_sqrtm"])
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.standard_normal(st.cov_sqrtm.shape))
    dense = st.cov_sqrtm @ Q
    ref = ek1_np.white_step(case["opde"], st._replace(cov_sqrtm=dense), case["dt"], 2, case["gram_sqrtm"])
    mean, chol, err, diff = mdl.step(st.mean, dense, case["dt"], dense=True)
    assert cases.mean_excess(mean, ref.mean) < 1
    assert cases.cov_excess(chol, ref.cov_sqrtm, 3) < 1
    assert diff == pytest.approx(ref.diffusion_squared_local, rel=1e-9)


def test_structure_envelopes_are_monotone_and_cover_diagonal():
    case = cases.make_case("sir", num=6)
    mdl = kernel_model.Model(case["pde"], "latent", 2, case["gram_sqrtm"])
    D, m = mdl.D, mdl.m
    for te in (mdl.te_p, mdl.te_u):
        assert np.all(np.diff(te) >= 0)
    assert np.all(mdl.te_p >= np.minimum(np.arange(D), D - 1))
    assert np.all(mdl.te_u >= np.minimum(np.arange(m + D), D - 1))
    assert np.all(np.diff(mdl.be_p) >= 0) and np.all(np.diff(mdl.be_u) >= 0)
