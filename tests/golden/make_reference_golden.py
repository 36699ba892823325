"""Golden vectors produced by the REFERENCE'S OWN SOURCE for the EK1 path.

    python tests/golden/make_reference_golden.py          # needs /root/reference (this container only)

jax/jaxlib (<= 0.3.1) and tornadox cannot be installed here, so `import pnmol` fails on the stock interpreter.  This
script puts a NumPy-backed stand-in for the used slice of the JAX API (tests/golden/jax_numpy_shim: jax.numpy ->
numpy, jax.scipy.linalg -> scipy.linalg, jit -> identity, vmap -> loop) in front of the path and then executes the
reference's unmodified files

    src/pnmol/white.py, latent.py, pdefilter.py, base/sqrt.py, base/iwp.py, base/stacked_ssm.py, base/rv.py,
    odetools/step.py

on float64 NumPy: `Solver(...).solve(pde)` with `step.Constant`.  The shim has no autodiff, so the discretised
problem (L, E_sqrtm, B, R_sqrtm, y0, f, df) and the spatial Gram matrix are supplied as arrays/callables (built by
oracle/setup_np.py, which restates src/pnmol/{discretize,kernels,mesh}.py and pde/examples.py and is pinned to the
reference's own known-answer tests in tests/test_oracle_primitives.py); everything from `initialize` on -- the hot
path of SURVEY section 8 -- is the reference's code.  The linear-algebra kernels are the same LAPACK routines jaxlib's
CPU backend dispatches to (geqrf, trtrs/trsm, potrf, getrf) through NumPy/SciPy.

Output: tests/golden/reference_<config>.npz with the inputs and the reference's trajectory (t, mean, cov_sqrtm,
calibrated diffusion, info counters).  tests/test_reference_golden.py pins the oracle (CPU) and the CUDA path (GPU)
to them.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE_SRC = os.environ.get("PNMOL_REFERENCE_SRC", "/root/reference/src")
sys.path[:0] = [os.path.join(HERE, "jax_numpy_shim"), REFERENCE_SRC, ROOT]

import pnmol  # noqa: E402  (the reference, running on the shim)
from oracle import setup_np  # noqa: E402

CONFIGS = {
    "heat_dirichlet_white_linear": ("heat", "white_linear", dict(num=6, bcond="dirichlet")),
    "heat_neumann_white_linear": ("heat", "white_linear", dict(num=6, bcond="neumann")),
    "heat_dirichlet_latent_linear": ("heat", "latent_linear", dict(num=6, bcond="dirichlet")),
    "heat_neumann_latent_linear": ("heat", "latent_linear", dict(num=6, bcond="neumann")),
    "spruce_dirichlet_white_semilinear": ("spruce", "white_semilinear", dict(num=6, bcond="dirichlet")),
    "spruce_neumann_white_semilinear": ("spruce", "white_semilinear", dict(num=6, bcond="neumann")),
    "spruce_dirichlet_latent_semilinear": ("spruce", "latent_semilinear", dict(num=6, bcond="dirichlet")),
    "sir_neumann_white_semilinear": ("sir", "white_semilinear", dict(num=5)),
    "lv_neumann_white_semilinear": ("lv", "white_semilinear", dict(num=6)),
    "heat_dirichlet_white_linear_N20": ("heat", "white_linear", dict(num=20, bcond="dirichlet")),
    # BASELINE-sized members (C1/C5: heat N=50, D=150; SIR N=17, D=153), two steps; first-order prior (figure-3 style)
    "heat_dirichlet_white_linear_N50": ("heat", "white_linear", dict(num=50, bcond="dirichlet", tmax=0.125)),
    "sir_neumann_white_semilinear_N17": ("sir", "white_semilinear", dict(num=17, tmax=0.125)),
    "heat_neumann_white_linear_nu1": ("heat", "white_linear", dict(num=8, bcond="neumann", nu=1)),
    "spruce_dirichlet_latent_semilinear_nu1": ("spruce", "latent_semilinear", dict(num=6, bcond="dirichlet", nu=1)),
}
SOLVERS = {
    "white_linear": pnmol.white.LinearWhiteNoiseEK1,
    "white_semilinear": pnmol.white.SemiLinearWhiteNoiseEK1,
    "latent_linear": pnmol.latent.LinearLatentForceEK1,
    "latent_semilinear": pnmol.latent.SemiLinearLatentForceEK1,
}
DT, NU, TMAX = 2.0 ** -4, 2, 0.5


def oracle_problem(prob, num, bcond="dirichlet", tmax=TMAX, nu=NU):
    """Same recipes as tests/cases.py:make_case (oracle side)."""
    if prob == "heat":
        return setup_np.heat_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond), 1
    if prob == "spruce":
        return setup_np.spruce_budworm_1d(num=num, tmax=tmax, diffusion_rate=0.05, bcond=bcond), 1
    if prob == "sir":
        return setup_np.sir_1d(num=num, tmax=tmax, diffusion_rates=(0.035,) * 3, n_bnd=min(5, num)), 3
    if prob == "lv":
        return setup_np.lotka_volterra_1d(num=num, tmax=tmax), 2
    raise KeyError(prob)


def main():
    for name, (prob, kind, kw) in CONFIGS.items():
        o, copies = oracle_problem(prob, **kw)
        gram = setup_np.gram(setup_np.Sum(setup_np.SE(), setup_np.White()), o.points, copies)
        pde = SimpleNamespace(L=o.L, E_sqrtm=o.E_sqrtm, B=o.B, R_sqrtm=o.R_sqrtm, y0=o.y0, t0=o.t0, tmax=o.tmax, f=o.f,
                              df=o.df, mesh_spatial=SimpleNamespace(points=o.points))
        nu = kw.get("nu", NU)
        solver = SOLVERS[kind](num_derivatives=nu, steprule=pnmol.odetools.step.Constant(DT),
                               spatial_kernel=lambda X, Y, g=gram: g)   # white.py:85 / latent.py:139: k(X, X.T)
        sol = solver.solve(pde)
        mean, chol = np.asarray(sol.mean), np.asarray(sol.cov_sqrtm)
        assert np.isfinite(mean).all() and np.isfinite(chol).all()
        info = {k: int(v) for k, v in sol.info.items()}
        # per-step by-products through the reference's own generator (white: error estimate and reference state)
        states = [st for st, _ in solver.solution_generator(pde)][1:]
        diffs = np.array([float(st.diffusion_squared_local) for st in states])
        extra = {}
        if kind.startswith("white"):
            extra = dict(error_estimate=np.stack([np.asarray(st.error_estimate) for st in states]),
                         reference_state=np.stack([np.asarray(st.reference_state) for st in states]))
        final, _ = solver.simulate_final_state(pde)  # pdefilter.py:105-116 (factor rescaled by the calibration)
        extra["final_cov_sqrtm"] = np.asarray(final.y.cov_sqrtm)
        np.savez_compressed(os.path.join(HERE, "reference_" + name + ".npz"), L=o.L, E_sqrtm=o.E_sqrtm, B=o.B,
                            R_sqrtm=o.R_sqrtm, y0=o.y0, gram=gram, dt=DT, nu=nu, tmax=o.tmax, t=np.asarray(sol.t), mean=mean,
                            cov_sqrtm=chol, diffusion_squared_calibrated=float(sol.diffusion_squared_calibrated),
                            kind=kind, problem=prob, num=kw["num"], bcond=kw.get("bcond", "neumann"),
                            diffusion_squared_local=diffs, **extra,
                            info_keys=np.array(sorted(info)), info_vals=np.array([info[k] for k in sorted(info)]))
        print(f"{name:40s} t {np.asarray(sol.t).shape} mean {mean.shape} chol {chol.shape} "
              f"sigma^2 {float(sol.diffusion_squared_calibrated):.6e} info {info}")


def adaptive():
    """simulate_final_state with step.Adaptive (pdefilter.py:105-227, odetools/step.py:58-133) on the heat problem."""
    o, _ = oracle_problem("heat", 9, "neumann")
    gram = setup_np.gram(setup_np.Sum(setup_np.SE(), setup_np.White()), o.points, 1)
    pde = SimpleNamespace(L=o.L, E_sqrtm=o.E_sqrtm, B=o.B, R_sqrtm=o.R_sqrtm, y0=o.y0, t0=o.t0, tmax=o.tmax, f=o.f, df=o.df,
                          mesh_spatial=SimpleNamespace(points=o.points))
    rule = dict(abstol=1e-3, reltol=1e-2)
    solver = pnmol.white.LinearWhiteNoiseEK1(num_derivatives=NU, steprule=pnmol.odetools.step.Adaptive(**rule),
                                             spatial_kernel=lambda X, Y: gram)
    final, info = solver.simulate_final_state(pde)
    info = {k: int(v) for k, v in info.items()}
    np.savez_compressed(os.path.join(HERE, "reference_adaptive_heat_neumann_white_linear.npz"), L=o.L, E_sqrtm=o.E_sqrtm,
                        B=o.B, R_sqrtm=o.R_sqrtm, y0=o.y0, gram=gram, nu=NU, tmax=TMAX, num=9, abstol=rule["abstol"],
                        reltol=rule["reltol"], t=float(final.t), mean=np.asarray(final.y.mean),
                        cov_sqrtm=np.asarray(final.y.cov_sqrtm), info_keys=np.array(sorted(info)),
                        info_vals=np.array([info[k] for k in sorted(info)]))
    print("adaptive heat neumann N=9:", info, "t =", float(final.t))


def kalman():
    """Filter step + square-root / covariance-form smoother step of src/pnmol/base/kalman.py on a random system
    (the scenario of the reference's tests/test_base/test_kalman.py)."""
    rng = np.random.default_rng(20261018)
    out = {}
    for d, k in ((4, 2), (12, 5), (33, 7)):
        phi = rng.standard_normal((d, d)) / np.sqrt(d) + np.eye(d)
        sq = np.tril(rng.standard_normal((d, d))) * 0.3 + 0.5 * np.eye(d)
        h, b, data = rng.standard_normal((k, d)), rng.standard_normal(k), rng.standard_normal(k)
        m, sc = rng.standard_normal(d), np.tril(rng.standard_normal((d, d))) + 2 * np.eye(d)
        m1, sc1, sgain, mp, scp, x = pnmol.base.kalman.filter_step(m, sc, phi, sq, h, b, data)
        m2, sc2, *_ = pnmol.base.kalman.filter_step(m1, sc1, phi, sq, h, b, data + 0.1)   # a "future" state
        ms, scs = pnmol.base.kalman.smoother_step_sqrt(m, sc, m2, sc2, sgain, sq, mp, x)
        mt, sct = pnmol.base.kalman.smoother_step_traditional(m, sc, m2, sc2, sgain, mp, scp)
        assert np.allclose(ms, mt) and np.allclose(scs @ scs.T, sct @ sct.T)
        for key, val in dict(phi=phi, sq=sq, h=h, b=b, data=data, m=m, sc=sc, m1=m1, sc1=sc1, sgain=sgain, mp=mp, scp=scp, x=x,
                             m_fut=m2, sc_fut=sc2, m_smooth=ms, sc_smooth=scs, sc_smooth_traditional=sct).items():
            out[f"d{d}_{key}"] = np.asarray(val)
    np.savez_compressed(os.path.join(HERE, "reference_kalman.npz"), sizes=np.array([4, 12, 33]), **out)
    print("kalman: filter_step + smoother_step_sqrt / traditional for d = 4, 12, 33")


def _reference_pde(o):
    return SimpleNamespace(L=o.L, E_sqrtm=o.E_sqrtm, B=o.B, R_sqrtm=o.R_sqrtm, y0=o.y0, t0=o.t0, tmax=o.tmax, f=o.f, df=o.df,
                           mesh_spatial=SimpleNamespace(points=o.points))


def baseline_c1_48_steps():
    """BASELINE config 1 / the C5 member at full length: heat N=50 (D=150, m=52), dt=2^-4, tmax=3 => 48 free-running
    steps through the reference's own solve().  All means, the factors at steps 0, 1, 12, 24, 36, 48, every local
    diffusion, the calibrated diffusion and the rescaled final factor of simulate_final_state."""
    o, _ = oracle_problem("heat", 50, "dirichlet", tmax=3.0)
    gram = setup_np.gram(setup_np.Sum(setup_np.SE(), setup_np.White()), o.points, 1)
    pde = _reference_pde(o)
    solver = pnmol.white.LinearWhiteNoiseEK1(num_derivatives=NU, steprule=pnmol.odetools.step.Constant(DT),
                                             spatial_kernel=lambda X, Y: gram)
    sol = solver.solve(pde)
    states = [st for st, _ in solver.solution_generator(pde)][1:]
    final, _ = solver.simulate_final_state(pde)
    keep = np.array([0, 1, 12, 24, 36, 48])
    info = {k: int(v) for k, v in sol.info.items()}
    np.savez_compressed(os.path.join(HERE, "baseline_c1_heat_N50_48steps.npz"), gram=gram, dt=DT, nu=NU, tmax=3.0, num=50,
                        t=np.asarray(sol.t), mean=np.asarray(sol.mean), keep=keep, cov_sqrtm_keep=np.asarray(sol.cov_sqrtm)[keep],
                        diffusion_squared_local=np.array([float(s.diffusion_squared_local) for s in states]),
                        diffusion_squared_calibrated=float(sol.diffusion_squared_calibrated),
                        error_estimate_last=np.asarray(states[-1].error_estimate),
                        final_cov_sqrtm=np.asarray(final.y.cov_sqrtm), y0=o.y0, L=o.L,
                        info_keys=np.array(sorted(info)), info_vals=np.array([info[k] for k in sorted(info)]))
    print("baseline C1 48 steps:", np.asarray(sol.mean).shape, "sigma^2", float(sol.diffusion_squared_calibrated))


def _digest(L, n, idx):
    """Compact, sign-invariant digest of a factor: diag(P), P on an index subset, Frobenius norms of the n x n
    derivative blocks of P = L L^T."""
    P = L @ L.T
    blocks = np.array([[np.linalg.norm(P[i::n, j::n]) for j in range(n)] for i in range(n)])
    return np.diag(P).copy(), P[np.ix_(idx, idx)].copy(), blocks


def baseline_c2_c3_two_steps():
    """BASELINE configs 2 and 3 at full size (SIR N=100: D=900, m=306, Matern-5/2 prior; spruce N=200 latent: D=1200,
    m=202), initialise + two free-running steps through the reference's own code.  The factors are 6.5-11.5 MB each, so
    the fixture keeps the means and a digest of every covariance (diagonal, a 120 x 120 principal sub-matrix that
    covers every derivative, block norms)."""
    rng = np.random.default_rng(20261018)
    for name, prob, kind, num, bcond, dt, prior in (("c2_sir_N100", "sir", "white_semilinear", 100, "neumann", 2.0 ** -3, "matern"),
                                                    ("c3_spruce_N200_latent", "spruce", "latent_semilinear", 200, "dirichlet", 2.0 ** -4, "se")):
        kw = dict(bcond=bcond) if prob != "sir" else {}
        o, copies = oracle_problem(prob, num, tmax=2 * dt, **kw)
        kern = setup_np.Sum(setup_np.Matern52() if prior == "matern" else setup_np.SE(), setup_np.White())
        gram = setup_np.gram(kern, o.points, copies)
        pde = _reference_pde(o)
        solver = SOLVERS[kind](num_derivatives=NU, steprule=pnmol.odetools.step.Constant(dt), spatial_kernel=lambda X, Y, g=gram: g)
        states = [st for st, _ in solver.solution_generator(pde)]
        assert len(states) == 3
        D = np.asarray(states[0].y.cov_sqrtm).shape[0]
        n = NU + 1
        idx = np.sort(rng.choice(D, size=120, replace=False))
        digs = [_digest(np.asarray(st.y.cov_sqrtm), n, idx) for st in states]
        extra = {}
        if kind.startswith("white"):
            extra = dict(error_estimate=np.stack([np.asarray(st.error_estimate) for st in states[1:]]))
        np.savez_compressed(os.path.join(HERE, f"baseline_{name}_2steps.npz"), dt=dt, nu=NU, num=num, kind=kind, problem=prob,
                            bcond=bcond, prior=prior, t=np.array([float(st.t) for st in states]),
                            mean=np.stack([np.asarray(st.y.mean) for st in states]), idx=idx,
                            cov_diag=np.stack([d[0] for d in digs]), cov_sub=np.stack([d[1] for d in digs]),
                            cov_block_norms=np.stack([d[2] for d in digs]),
                            diffusion_squared_local=np.array([float(st.diffusion_squared_local) for st in states[1:]]),
                            y0=o.y0, **extra)
        print(f"baseline {name}: D={D}, sigma^2_loc", [float(st.diffusion_squared_local) for st in states[1:]])


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    if only in ("all", "small"):
        main()
        adaptive()
        kalman()
    if only in ("all", "baseline"):
        baseline_c1_48_steps()
        baseline_c2_c3_two_steps()
