"""Pins against golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/make_reference_golden.py: the
unmodified src/pnmol/{white,latent,pdefilter}.py, base/*.py, odetools/step.py executed on a NumPy stand-in for the used
slice of the JAX API, because jax/jaxlib <= 0.3.1 cannot be installed here).

* CPU: the NumPy oracle equals the reference's trajectory for all four solver classes (means, factors, calibrated and
  local diffusions, error estimates, reference states, info counters, the adaptive step rule) -- this is what makes
  the oracle a trustworthy checker for everything else in tests/.
* GPU: the CUDA path (through the Python API -> ctypes -> C ABI) against the same files at the north-star
  tolerances (means rtol 1e-9, covariances as L L^T rtol 1e-8, per derivative block).
"""
import glob
import os

import numpy as np
import pytest

from oracle import ek1_np, setup_np

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "reference_*.npz")) if "adaptive" not in p and "kalman" not in p)
KALMAN = os.path.join(HERE, "golden", "reference_kalman.npz")
ADAPTIVE = os.path.join(HERE, "golden", "reference_adaptive_heat_neumann_white_linear.npz")
IDS = [os.path.basename(p)[len("reference_"):-4] for p in FILES]


def _oracle_problem(g):
    prob, num, bcond = str(g["problem"]), int(g["num"]), str(g["bcond"])
    tmax = float(g["tmax"])
    if prob == "heat":
        o = setup_np.heat_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
    elif prob == "spruce":
        o = setup_np.spruce_budworm_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond)
    elif prob == "sir":
        o = setup_np.sir_1d(num=num, tmax=tmax, diffusion_rates=(0.035,) * 3, n_bnd=min(5, num))
    else:
        o = setup_np.lotka_volterra_1d(num=num, tmax=tmax)
    for key in ("L", "E_sqrtm", "B", "R_sqrtm", "y0"):  # the frozen inputs are the ones the reference ran on
        assert np.array_equal(getattr(o, key), g[key]), key
    return o


def _info(g):
    return {str(k): int(v) for k, v in zip(g["info_keys"], g["info_vals"])}


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_equals_reference_source(path):
    g = np.load(path, allow_pickle=False)
    kind, nu, dt = str(g["kind"]), int(g["nu"]), float(g["dt"])
    n = nu + 1
    o = _oracle_problem(g)
    Lk = np.linalg.cholesky(g["gram"])  # white.py:85 / latent.py:139
    sol = ek1_np.solve(kind, o, dt, nu, Lk)
    assert np.array_equal(sol.t, g["t"]) and sol.info == _info(g)
    # same NumPy/LAPACK calls in the same order: equal to rounding (bitwise on the generating machine)
    for k in range(len(g["t"])):
        assert cases.block_rel(sol.mean[k], g["mean"][k], n) <= 1e-12, k
        assert cases.block_rel(cases.cov(sol.cov_sqrtm[k]), cases.cov(g["cov_sqrtm"][k]), n) <= 1e-11, k
    assert float(sol.diffusion_squared_calibrated) == pytest.approx(float(g["diffusion_squared_calibrated"]), rel=1e-10)
    states = list(ek1_np.generate(kind, o, dt, nu, Lk))[1:]
    assert np.allclose([float(s.diffusion_squared_local) for s in states], g["diffusion_squared_local"], rtol=1e-10)
    if kind.startswith("white"):
        assert np.allclose(np.stack([s.error_estimate for s in states]), g["error_estimate"], rtol=1e-10, atol=0)
        assert np.allclose(np.stack([s.reference_state for s in states]), g["reference_state"], rtol=1e-12, atol=1e-300)
    final, _ = ek1_np.simulate_final_state(kind, o, dt, nu, Lk)
    assert cases.block_rel(cases.cov(final.cov_sqrtm), cases.cov(g["final_cov_sqrtm"]), n) <= 1e-10


def test_oracle_adaptive_equals_reference_source():
    g = np.load(ADAPTIVE, allow_pickle=False)
    o = setup_np.heat_1d(num=int(g["num"]), tmax=float(g["tmax"]), diffusion_rate=0.05, bcond="neumann")
    assert np.array_equal(o.L, g["L"]) and np.array_equal(o.y0, g["y0"])
    Lk = np.linalg.cholesky(g["gram"])
    final, cal, info = ek1_np.simulate_final_state_adaptive("white_linear", o, int(g["nu"]), Lk, abstol=float(g["abstol"]),
                                                           reltol=float(g["reltol"]))
    assert info == _info(g) and final.t == float(g["t"])
    assert cases.block_rel(final.mean, g["mean"], 3) <= 1e-12
    assert cases.block_rel(cases.cov(final.cov_sqrtm), cases.cov(g["cov_sqrtm"]), 3) <= 1e-10


def test_oracle_kalman_equals_reference_source():
    from oracle import kalman_np

    g = np.load(KALMAN, allow_pickle=False)
    for d in g["sizes"]:
        v = {k[len(f"d{d}_"):]: g[k] for k in g.files if k.startswith(f"d{d}_")}
        m1, sc1, sgain, mp, scp, x = kalman_np.filter_step(v["m"], v["sc"], v["phi"], v["sq"], v["h"], v["b"], v["data"])
        assert np.allclose(m1, v["m1"], rtol=1e-12, atol=1e-13) and np.allclose(sgain, v["sgain"], rtol=1e-11, atol=1e-13)
        assert np.allclose(sc1 @ sc1.T, v["sc1"] @ v["sc1"].T, rtol=1e-11, atol=1e-13)
        ms, scs = kalman_np.smoother_step_sqrt(v["m"], v["sc"], v["m_fut"], v["sc_fut"], v["sgain"], v["sq"], v["mp"], v["x"])
        assert np.allclose(ms, v["m_smooth"], rtol=1e-12, atol=1e-13) and np.allclose(scs, v["sc_smooth"], rtol=1e-9, atol=1e-11)
        mt, sct = kalman_np.smoother_step_traditional(v["m"], v["sc"], v["m_fut"], v["sc_fut"], v["sgain"], v["mp"], v["scp"])
        assert np.allclose(sct, v["sc_smooth_traditional"], rtol=1e-10, atol=1e-12)


# ----------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
def test_cuda_smoother_step_reproduces_reference_source():
    """the square-root RTS smoother step (kalman.py:49-66) and the filter step it follows."""
    import torch

    import __graft_entry__

    __graft_entry__.ensure_built()
    from pnmol_b200.base import kalman

    g = np.load(KALMAN, allow_pickle=False)
    for d in g["sizes"]:
        v = {k[len(f"d{d}_"):]: g[k] for k in g.files if k.startswith(f"d{d}_")}
        ms, scs = kalman.smoother_step_sqrt(v["m"], v["sc"], v["m_fut"], v["sc_fut"], v["sgain"], v["sq"], v["mp"], v["x"])
        ms, scs = ms.cpu().numpy(), scs.cpu().numpy()
        assert np.allclose(ms, v["m_smooth"], rtol=1e-11, atol=1e-13)
        assert np.array_equal(np.triu(scs, 1), np.zeros_like(scs))
        assert np.allclose(scs @ scs.T, v["sc_smooth"] @ v["sc_smooth"].T, rtol=1e-9, atol=1e-12)
        assert np.allclose(scs, v["sc_smooth"], rtol=1e-8, atol=1e-11)  # same signs as LAPACK (dlarfg convention)
        mt, sct = kalman.smoother_step_traditional(v["m"], v["sc"], v["m_fut"], v["sc_fut"], v["sgain"], v["mp"], v["scp"])
        assert np.allclose(scs @ scs.T, (sct @ sct.T).cpu().numpy(), rtol=1e-8, atol=1e-11)   # tests/test_base/test_kalman.py:131-135
        out = kalman.filter_step(v["m"], v["sc"], v["phi"], v["sq"], v["h"], v["b"], v["data"])
        assert np.allclose(out[0].cpu().numpy(), v["m1"], rtol=1e-9, atol=1e-12)
        assert np.allclose(out[2].cpu().numpy(), v["sgain"], rtol=1e-8, atol=1e-11)
        L1 = out[1].cpu().numpy()
        assert np.allclose(L1 @ L1.T, v["sc1"] @ v["sc1"].T, rtol=1e-8, atol=1e-11)
        # batched call = stacked single calls
        stack = lambda key: np.stack([v[key], v[key]])
        mb, sb = kalman.smoother_step_sqrt(*[stack(k) for k in ("m", "sc", "m_fut", "sc_fut", "sgain", "sq", "mp", "x")])
        assert torch.equal(mb[0], mb[1]) and np.allclose(sb[1].cpu().numpy(), scs)



def _product_case(g):
    prob, num, bcond = str(g["problem"]), int(g["num"]), str(g["bcond"])
    kw = dict(bcond=bcond) if prob in ("heat", "spruce") else {}
    return cases.make_case(prob, num=num, tmax=float(g["tmax"]), nu=int(g["nu"]), **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["cta", "large"])
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_path_reproduces_reference_source(path, family, monkeypatch):
    import __graft_entry__

    __graft_entry__.ensure_built()
    monkeypatch.setenv("PNMOL_B200_PATH", family)
    g = np.load(path, allow_pickle=False)
    kind, nu, dt = str(g["kind"]), int(g["nu"]), float(g["dt"])
    n = nu + 1
    case = _product_case(g)
    assert np.allclose(case["gram_sqrtm"], np.linalg.cholesky(g["gram"]), rtol=1e-13, atol=0)
    sol = cases.make_solver(kind, case).solve(case["pde"])
    assert np.array_equal(np.asarray(sol.t), g["t"]) and sol.info == _info(g)
    with cases.perturbed_oracle():  # the reference's own reproducibility (conditioning floor), see cases.mean_excess
        eps = ek1_np.solve(kind, case["opde"], dt, nu, case["gram_sqrtm"])
    mean, chol = sol.mean.cpu().numpy(), sol.cov_sqrtm.cpu().numpy()
    for k in range(len(g["t"])):
        assert cases.mean_excess(mean[k], g["mean"][k], spread=eps.mean[k]) < 1, k
        assert cases.cov_excess(chol[k], g["cov_sqrtm"][k], n) < 1, k


@pytest.mark.gpu
def test_cuda_adaptive_reproduces_reference_source():
    import __graft_entry__

    __graft_entry__.ensure_built()
    from pnmol_b200 import white
    from pnmol_b200.odetools import step

    g = np.load(ADAPTIVE, allow_pickle=False)
    case = cases.make_case("heat", num=int(g["num"]), bcond="neumann", tmax=float(g["tmax"]))
    solver = white.LinearWhiteNoiseEK1(num_derivatives=int(g["nu"]), spatial_kernel=case["kernel"],
                                       steprule=step.Adaptive(abstol=float(g["abstol"]), reltol=float(g["reltol"])))
    state, info = solver.simulate_final_state(case["pde"])
    assert info == _info(g) and state.t == float(g["t"])
    assert cases.mean_excess(state.y.mean.cpu().numpy(), g["mean"]) < 1
