"""Oracle (test infrastructure): the EK1 filter loop of the reference in NumPy float64.

Follows, line by line and in the reference's own dense formulation,
``src/pnmol/white.py:12-208`` (white-noise EK1), ``src/pnmol/latent.py:20-292``
(latent-force EK1), ``src/pnmol/pdefilter.py:75-227`` (driver loop) and
``src/pnmol/odetools/step.py:30-133`` (constant and adaptive steps).  Pinned to golden vectors produced
by the reference's own source (``tests/golden/make_reference_golden.py``, see ``oracle/__init__.py``).

A problem is any object with attributes ``L (d,d)``, ``E_sqrtm (d,d)``, ``B (nb,d)``,
``R_sqrtm (nb,nb)``, ``y0 (d,)``, ``t0``, ``tmax`` and, for semi-linear solvers,
``f(t, x)`` / ``df(t, x)`` returning ``(d,)`` / dense ``(d,d)`` NumPy arrays.
"""
from collections import namedtuple

import numpy as np
import scipy.linalg

from . import prior_np, sqrt_np

State = namedtuple("State", "t mean cov_sqrtm error_estimate reference_state diffusion_squared_local")
Solution = namedtuple("Solution", "t mean cov_sqrtm info diffusion_squared_calibrated")

WHITE_NUGGET = 1e-10  # white.py:33,51
LATENT_NUGGET = 1e-6  # latent.py:71,98


# --------------------------------------------------------------------------- white
def white_linearise(prob, p0, p1, m_pred, t, semilinear):
    """white.py:169-186 (linear) / 189-208 (semi-linear)."""
    L, B = prob.L, prob.B
    m_at = p0 @ m_pred
    if semilinear:
        fx = prob.f(t, m_at)
        Jx = prob.df(t, m_at)
        H_ode = p1 - Jx @ p0 - L @ p0
    else:
        fx = L @ m_at
        Jx = L
        H_ode = p1 - Jx @ p0
    b = Jx @ m_at - fx
    H = np.vstack((H_ode, B @ p0))
    z = H @ m_pred + np.concatenate((b, np.zeros(B.shape[0])))
    E = scipy.linalg.block_diag(prob.E_sqrtm, prob.R_sqrtm)
    return z, H, E


def white_error_estimate(Ql, z, H, E):
    """white.py:153-162."""
    S = H @ (Ql @ Ql.T) @ H.T + E @ E.T
    sigma_sq = z @ np.linalg.solve(S, z) / z.shape[0]
    sigma = np.sqrt(sigma_sq)
    return sigma, np.sqrt(np.diag(S)) * sigma


def white_initialize(prob, nu, gram_sqrtm, prior_scale=1.0, semilinear=False):
    """white.py:12-80.  ``gram_sqrtm`` = chol(k(X,X)) (white.py:85)."""
    n, d = nu + 1, prob.L.shape[0]
    E0 = prior_np.projection(nu, d, 0)
    E1 = prior_np.projection(nu, d, 1)
    C0_raw = np.kron(gram_sqrtm, prior_scale * np.eye(n))
    C0_y0, K_y0, _ = sqrt_np.measurement_update(E0, C0_raw, WHITE_NUGGET * np.eye(d))
    m_y0 = K_y0 @ prob.y0
    z, H, E = white_linearise(prob, E0, E1, m_y0, prob.t0, semilinear)
    nug = WHITE_NUGGET * np.eye(d + prob.B.shape[0])
    C0, K, _ = sqrt_np.measurement_update(H, C0_y0, E + nug)
    m0 = m_y0 - K @ z
    return State(prob.t0, m0.reshape((n, d), order="F"), C0, None, None, [])


def white_step(prob, state, dt, nu, gram_sqrtm, semilinear=False):
    """white.py:96-146."""
    n, d = nu + 1, prob.y0.shape[0]
    nb = prob.B.shape[0]
    P, Pinv = prior_np.nordsieck_dense(nu, d, dt)
    A, Ql = prior_np.iwp_dense(nu, gram_sqrtm)
    E0 = prior_np.projection(nu, d, 0)
    E1 = prior_np.projection(nu, d, 1)

    m = Pinv @ state.mean.reshape((-1,), order="F")
    Cl = Pinv @ state.cov_sqrtm
    mp = A @ m
    z, H, E = white_linearise(prob, E0 @ P, E1 @ P, mp, state.t + dt, semilinear)
    _, err = white_error_estimate(Ql, z, H, E)
    Clp = sqrt_np.chol_of_sum(A @ Cl, Ql)
    err = err[:-nb]
    Cl_new, K, Sl = sqrt_np.measurement_update(H, Clp, E)
    m_new = mp - K @ z
    # quirk Q1 (white.py:125): solves R1 x = z, not R1^T x = z
    white_res = scipy.linalg.solve_triangular(Sl.T, z, lower=False)
    diff_sq = white_res @ white_res / white_res.shape[0]
    err = dt * err
    Cl_new = P @ Cl_new
    m_new = (P @ m_new).reshape((n, d), order="F")
    return State(state.t + dt, m_new, Cl_new, err, np.abs(m_new[0]), diff_sq)


# -------------------------------------------------------------------------- latent
def latent_linearise(prob, E0, E1, m_pred, t, P_state, P_eps, semilinear):
    """latent.py:237-263 (linear) / 266-292 (semi-linear)."""
    L = prob.L
    E0_state = E0 @ P_state
    E0_eps = E0 @ P_eps
    E1_state = E1 @ P_state
    m_at = scipy.linalg.block_diag(E0_state, E0_eps) @ m_pred
    state_at, _ = np.split(m_at, 2)
    if semilinear:
        fx = prob.f(t, state_at)
        Jx = prob.df(t, state_at)
        H_state = E1_state - Jx @ E0_state - L @ E0_state
    else:
        fx = L @ state_at
        Jx = L
        H_state = E1_state - Jx @ E0_state
    H_bc = prob.B @ E0_state
    H = np.block([[H_state, -E0_eps], [H_bc, np.zeros_like(H_bc)]])
    b = np.concatenate([Jx @ state_at - fx, np.zeros(prob.B.shape[0])])
    return H @ m_pred + b, H


def latent_initialize(prob, nu, gram_sqrtm, prior_scale=1.0, semilinear=False):
    """latent.py:20-134."""
    n, d = nu + 1, prob.L.shape[0]
    E0 = prior_np.projection(nu, d, 0)
    E1 = prior_np.projection(nu, d, 1)
    c0 = prior_scale * np.eye(n)
    C0_state_raw = np.kron(gram_sqrtm, c0)
    C0_lat_raw = np.kron(prob.E_sqrtm, c0)
    C0_state_y0, K_y0, _ = sqrt_np.measurement_update(E0, C0_state_raw, LATENT_NUGGET * np.eye(d))
    m_state = K_y0 @ prob.y0
    m_stack = np.concatenate((m_state, np.zeros(n * d)))
    C_block = scipy.linalg.block_diag(C0_state_y0, C0_lat_raw)
    eye = np.eye(n * d)
    z, H = latent_linearise(prob, E0, E1, m_stack, prob.t0, eye, eye, semilinear)
    nug = LATENT_NUGGET * np.eye(d + prob.B.shape[0])
    C0, K, _ = sqrt_np.measurement_update(H, C_block, nug)
    m0 = m_stack - K @ z
    ms, ml = np.split(m0, 2)
    mean = np.concatenate((ms.reshape((n, d), order="F"), ml.reshape((n, d), order="F")), axis=1)
    return State(prob.t0, mean, C0, None, None, [])


def latent_step(prob, state, dt, nu, gram_sqrtm, semilinear=False):
    """latent.py:155-225."""
    n, d = nu + 1, prob.L.shape[0]
    P1, P1inv = prior_np.nordsieck_dense(nu, d, dt)
    P = prior_np.stack_blockdiag(P1, P1)
    Pinv = prior_np.stack_blockdiag(P1inv, P1inv)
    A_s, Q_s = prior_np.iwp_dense(nu, gram_sqrtm)
    A_l, Q_l = prior_np.iwp_dense(nu, prob.E_sqrtm)
    A = prior_np.stack_blockdiag(A_s, A_l)
    Ql = prior_np.stack_blockdiag(Q_s, Q_l)
    E0 = prior_np.projection(nu, d, 0)
    E1 = prior_np.projection(nu, d, 1)

    ms, me = np.split(state.mean, 2, axis=-1)
    flat = np.concatenate((ms.reshape((-1,), order="F"), me.reshape((-1,), order="F")))
    flat = Pinv @ flat
    Cl = Pinv @ state.cov_sqrtm
    mp = A @ flat
    z, H = latent_linearise(prob, E0, E1, mp, state.t + dt, P1, P1, semilinear)
    Clp = sqrt_np.chol_of_sum(A @ Cl, Ql)
    Cl_new, K, Sl = sqrt_np.measurement_update(H, Clp, None)
    m_new = P @ (mp - K @ z)
    Cl_new = P @ Cl_new
    white_res = scipy.linalg.solve_triangular(Sl.T, z, lower=False)  # quirk Q1, latent.py:204
    diff_sq = white_res @ white_res / white_res.shape[0]
    ns, ne = np.split(m_new, 2)
    mean = np.concatenate((ns.reshape((n, d), order="F"), ne.reshape((n, d), order="F")), axis=-1)
    return State(state.t + dt, mean, Cl_new, None, None, diff_sq)


# -------------------------------------------------------------------------- driver
KINDS = {
    "white_linear": (white_initialize, white_step, False),
    "white_semilinear": (white_initialize, white_step, True),
    "latent_linear": (latent_initialize, latent_step, False),
    "latent_semilinear": (latent_initialize, latent_step, True),
}


def constant_step_schedule(t0, tmax, dt):
    """Time grid of ``solution_generator`` + ``perform_full_step`` with ``step.Constant``
    (pdefilter.py:140,220-223; step.py:30-55): ``t`` is accumulated in floating point and
    ``dt`` clipped to ``tmax - t``, which produces the rounding "sliver" step (quirk Q5)."""
    ts, dts = [t0], []
    t, h = t0, dt
    while t < tmax:
        dts.append(h)
        t = t + h
        ts.append(t)
        h = min(dt, tmax - t)
        assert h >= 0
    return np.array(ts), np.array(dts)


def generate(kind, prob, dt, nu, gram_sqrtm, prior_scale=1.0):
    """Yield states like ``PDEFilter.solution_generator`` (pdefilter.py:118-165) with
    constant steps."""
    init, step, semil = KINDS[kind]
    state = init(prob, nu, gram_sqrtm, prior_scale, semil)
    yield state
    h = dt
    while state.t < prob.tmax:
        state = step(prob, state, h, nu, gram_sqrtm, semil)
        h = min(dt, prob.tmax - state.t)
        yield state


def solve(kind, prob, dt, nu, gram_sqrtm, prior_scale=1.0):
    """``PDEFilter.solve`` (pdefilter.py:75-103)."""
    ts, means, covs, diffs = [], [], [], []
    for st in generate(kind, prob, dt, nu, gram_sqrtm, prior_scale):
        ts.append(st.t)
        means.append(st.mean)
        covs.append(st.cov_sqrtm)
        if isinstance(st.diffusion_squared_local, list):
            diffs.extend(st.diffusion_squared_local)
        else:
            diffs.append(st.diffusion_squared_local)
    nsteps = len(ts) - 1
    info = dict(num_f_evaluations=nsteps, num_df_evaluations=nsteps, num_df_diagonal_evaluations=0,
                num_steps=nsteps, num_attempted_steps=nsteps)
    return Solution(np.stack(ts), np.stack(means), np.stack(covs), info, np.mean(np.array(diffs)))


def simulate_final_state(kind, prob, dt, nu, gram_sqrtm, prior_scale=1.0):
    """``PDEFilter.simulate_final_state`` (pdefilter.py:105-116): last state with the
    factor rescaled by the calibrated diffusion."""
    diffs, st = [], None
    for st in generate(kind, prob, dt, nu, gram_sqrtm, prior_scale):
        if not isinstance(st.diffusion_squared_local, list):
            diffs.append(st.diffusion_squared_local)
    cal = np.mean(np.array(diffs))
    return st._replace(cov_sqrtm=st.cov_sqrtm * np.sqrt(cal)), cal


def adaptive_first_dt(prob, semilinear=False):
    """``Adaptive.first_dt`` (odetools/step.py:103-133): 0.01 ||y0|| / ||f(t0, y0)||."""
    dy0 = prob.f(prob.t0, prob.y0) if semilinear else prob.L @ prob.y0
    return 0.01 * np.linalg.norm(prob.y0) / np.linalg.norm(dy0)


def simulate_final_state_adaptive(kind, prob, nu, gram_sqrtm, *, abstol=1e-4, reltol=1e-2, max_changes=(0.2, 10.0),
                                  safety_scale=0.95, prior_scale=1.0, first_dt=None):
    """``PDEFilter.simulate_final_state`` with ``step.Adaptive`` (pdefilter.py:105-227, odetools/step.py:58-119):
    returns (final state with the factor rescaled, calibration, info)."""
    init, step, semil = KINDS[kind]
    state = init(prob, nu, gram_sqrtm, prior_scale, semil)
    dt = adaptive_first_dt(prob, semil) if first_dt is None else first_dt
    diffs, nsteps, nattempts = [], 0, 0
    small, large = max_changes
    while state.t < prob.tmax:
        accepted = False
        while not accepted:
            proposed = step(prob, state, dt, nu, gram_sqrtm, semil)
            nattempts += 1
            ratio = (dt * proposed.error_estimate) / (abstol + reltol * proposed.reference_state)
            norm = np.linalg.norm(ratio) / np.sqrt(ratio.size)
            accepted = bool(norm < 1)
            change = safety_scale * (1.0 / norm) ** (1.0 / (nu + 1))
            suggested = max(small, min(change, large)) * dt
            dt = min(suggested, prob.tmax - (proposed.t if accepted else state.t))
            assert dt >= 0, f"Invalid step size: dt={dt}"
        state = proposed
        nsteps += 1
        diffs.append(state.diffusion_squared_local)
    cal = np.mean(np.array(diffs))
    info = dict(num_f_evaluations=nattempts, num_df_evaluations=nattempts, num_df_diagonal_evaluations=0,
                num_steps=nsteps, num_attempted_steps=nattempts)
    return state._replace(cov_sqrtm=state.cov_sqrtm * np.sqrt(cal)), cal, info
