"""PDE filter driver (API of src/pnmol/pdefilter.py).

Same public surface as the reference: ``PDEFilterState``, ``PDESolution`` and the
``PDEFilter`` base class with ``solve`` / ``simulate_final_state`` / ``solution_generator`` /
``perform_full_step``.  ``attempt_step`` of the subclasses is one CUDA launch; with constant
steps and no stop locations ``solve`` and ``simulate_final_state`` take the persistent
route (one launch for the whole time loop, ``pnmol_b200_run``), which executes the very
same per-step device code.
"""
import dataclasses
from abc import ABC, abstractmethod
from collections import namedtuple
from typing import Dict, Iterable

import os

import numpy as np
import torch
from tqdm import tqdm

from . import _engine, kernels
from .base import rv
from .odetools import step


class PDEFilterState(namedtuple("_", "t y error_estimate reference_state diffusion_squared_local")):
    """PDE filter state (pdefilter.py:17-22)."""


@dataclasses.dataclass(frozen=False)
class PDESolution:
    t: object
    mean: object
    cov_sqrtm: object
    info: Dict
    diffusion_squared_calibrated: float


def _new_info():
    return dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0, num_steps=0,
                num_attempted_steps=0)


class PDEFilter(ABC):
    def __init__(self, *, steprule=None, num_derivatives=2, spatial_kernel=None, diffuse_prior_scale=1e0):
        self.steprule = steprule or step.Adaptive()
        self.num_derivatives = num_derivatives
        self.iwp = None
        self.spatial_kernel = spatial_kernel or kernels.Matern52() + kernels.WhiteNoise()
        self.E0 = None
        self.E1 = None
        self.diffuse_prior_scale = diffuse_prior_scale
        self._engine = None
        self._engine_pde = None

    def __repr__(self):
        return (f"{type(self).__name__}(num_derivatives={self.num_derivatives}, steprule={self.steprule}, "
                f"spatial_kernel={self.spatial_kernel})")

    # ------------------------------------------------------------------ drivers
    def _persistent_ok(self, stop_at, progressbar):
        return isinstance(self.steprule, step.Constant) and stop_at is None and not progressbar

    def solve(self, pde, /, *, stop_at=None, progressbar=False):
        """pdefilter.py:75-103."""
        if self._persistent_ok(stop_at, progressbar):
            return self._solve_persistent(pde)
        if self._device_adaptive_ok(stop_at, progressbar):
            state0 = self.initialize(pde)
            return self._solve_adaptive_device(pde, state0)   # (all three kernel families serve the adaptive loop)
        means, covs, times, diffs, info = [], [], [], [], dict()
        for state, info in self.solution_generator(pde, stop_at=stop_at, progressbar=progressbar):
            times.append(state.t)
            means.append(state.y.mean)
            covs.append(state.y.cov_sqrtm)
            if isinstance(state.diffusion_squared_local, list):
                diffs.extend(state.diffusion_squared_local)
            else:
                diffs.append(state.diffusion_squared_local)
        cal = torch.stack([torch.as_tensor(x) for x in diffs]).mean() if diffs else torch.tensor(float("nan"))
        return PDESolution(t=np.asarray(times), mean=torch.stack(means), cov_sqrtm=torch.stack(covs), info=info,
                           diffusion_squared_calibrated=cal)

    def simulate_final_state(self, pde, /, *, stop_at=None, progressbar=False):
        """pdefilter.py:105-116: final state, factor rescaled by the calibrated diffusion."""
        if self._persistent_ok(stop_at, progressbar):
            state0 = self.initialize(pde)
            dts = _engine.constant_step_schedule(pde.t0, pde.tmax, self.steprule.first_dt(pde))
            eng = self._engine
            mean = state0.y.mean[None].contiguous()
            chol = state0.y.cov_sqrtm[None].contiguous()
            out = eng.run(pde.t0, dts, mean, chol)
            cal = eng.rescale(chol, out["diff_sum"], len(dts))
            t = pde.t0
            for h in dts:
                t = t + h
            info = _new_info()
            info.update(num_f_evaluations=len(dts), num_df_evaluations=len(dts), num_steps=len(dts),
                        num_attempted_steps=len(dts))
            white = out["err"] is not None
            y = rv.MultivariateNormal(mean[0], _mark_tril(chol[0]))
            return PDEFilterState(t=t, y=y, error_estimate=out["err"][0] if white else None,
                                  reference_state=out["ref"][0] if white else None,
                                  diffusion_squared_local=out["diff_last"][0]), info
        if self._device_adaptive_ok(stop_at, progressbar):
            state0 = self.initialize(pde)
            eng = self._engine
            if True:  # accept/reject and the step-size proposal run inside one kernel launch (all kernel families)
                mean = state0.y.mean[None].contiguous()
                chol = state0.y.cov_sqrtm[None].contiguous()
                out = eng.run_adaptive(pde.t0, pde.tmax, self.steprule.first_dt(pde), self.steprule, mean, chol)
                status = int(out["status"][0])
                if status & 2:
                    raise RuntimeError("adaptive time loop: attempt limit reached before tmax")
                nsteps, natt = int(out["num_steps"][0]), int(out["num_attempts"][0])
                eng.rescale(chol, out["diff_sum"] / nsteps, 1)
                info = _new_info()
                info.update(num_f_evaluations=natt, num_df_evaluations=natt, num_steps=nsteps, num_attempted_steps=natt)
                y = rv.MultivariateNormal(mean[0], _mark_tril(chol[0]))
                return PDEFilterState(t=float(out["t"][0]), y=y, error_estimate=out["err"][0], reference_state=out["ref"][0],
                                      diffusion_squared_local=out["diff_last"][0]), info
        state, info, diffs = None, None, []
        for state, info in self.solution_generator(pde, stop_at=stop_at, progressbar=progressbar):
            if isinstance(state.diffusion_squared_local, list):
                diffs.extend(state.diffusion_squared_local)
            else:
                diffs.append(state.diffusion_squared_local)
        cal = torch.stack([torch.as_tensor(x) for x in diffs]).mean()
        cov_new = state.y.cov_sqrtm * torch.sqrt(cal)
        return state._replace(y=state.y._replace(cov_sqrtm=cov_new)), info

    def _device_adaptive_ok(self, stop_at, progressbar):
        """Accept/reject, step-size control and (for solve) the trajectory of accepted states run inside ONE kernel launch
        for the white-noise solvers (the latent-force solvers have no error estimate, latent.py:217-223)."""
        return (isinstance(self.steprule, step.Adaptive) and self.family == "white" and stop_at is None and not progressbar
                and os.environ.get("PNMOL_B200_HOST_ADAPTIVE") != "1")

    def _solve_adaptive_device(self, pde, state0, capacity=64):
        """solve() with step.Adaptive (pdefilter.py:75-103, 192-227) in one launch: the kernel appends every accepted state
        to a trajectory buffer; if a solve accepts more steps than the buffer holds, it is repeated once with the exact
        capacity (the step count is known then)."""
        eng = self._engine
        while True:
            mean = state0.y.mean[None].clone()
            chol = state0.y.cov_sqrtm[None].clone()
            out = eng.run_adaptive(pde.t0, pde.tmax, self.steprule.first_dt(pde), self.steprule, mean, chol, trajectory=capacity)
            status = int(out["status"][0])
            if status & 2:
                raise RuntimeError("adaptive time loop: attempt limit reached before tmax")
            nsteps, natt = int(out["num_steps"][0]), int(out["num_attempts"][0])
            if not (status & 4):
                break
            capacity = nsteps
        info = _new_info()
        info.update(num_f_evaluations=natt, num_df_evaluations=natt, num_steps=nsteps, num_attempted_steps=natt)
        ts = np.concatenate([[pde.t0], out["t_traj"][0, :nsteps].cpu().numpy()])
        means = torch.cat([state0.y.mean[None], out["mean_traj"][:nsteps, 0]])
        covs = torch.cat([state0.y.cov_sqrtm[None], out["chol_traj"][:nsteps, 0]])
        return PDESolution(t=ts, mean=means, cov_sqrtm=covs, info=info, diffusion_squared_calibrated=out["diff_sum"][0] / nsteps)

    def _solve_persistent(self, pde):
        state0 = self.initialize(pde)
        dts = _engine.constant_step_schedule(pde.t0, pde.tmax, self.steprule.first_dt(pde))
        eng = self._engine
        mean = state0.y.mean[None].clone()
        chol = state0.y.cov_sqrtm[None].clone()
        out = eng.run(pde.t0, dts, mean, chol, trajectory=True)
        ts = [pde.t0]
        for h in dts:
            ts.append(ts[-1] + h)
        info = _new_info()
        info.update(num_f_evaluations=len(dts), num_df_evaluations=len(dts), num_steps=len(dts),
                    num_attempted_steps=len(dts))
        means = torch.cat([state0.y.mean[None], out["mean_traj"][:, 0]])
        covs = torch.cat([state0.y.cov_sqrtm[None], out["chol_traj"][:, 0]])
        cal = out["diff_sum"][0] / len(dts)
        return PDESolution(t=np.asarray(ts), mean=means, cov_sqrtm=covs, info=info, diffusion_squared_calibrated=cal)

    def solve_marginals(self, pde):
        """Like solve(), but returns MarginalSolution(t, mean (T+1, n, dd), std (T+1, dd), info, diffusion) -- the
        marginal standard deviations of the 0th derivative instead of the factor trajectory (constant steps only;
        the read-out is fused into the persistent step kernel)."""
        from . import marginals

        if not self._persistent_ok(None, False):
            raise NotImplementedError("solve_marginals needs a Constant step rule")
        state0 = self.initialize(pde)
        dts = _engine.constant_step_schedule(pde.t0, pde.tmax, self.steprule.first_dt(pde))
        eng = self._engine
        mean = state0.y.mean[None].clone()
        chol = state0.y.cov_sqrtm[None].clone()
        std0 = eng.marginal_std(chol)
        out = eng.run_marginals(pde.t0, dts, mean, chol)
        ts = [pde.t0]
        for h in dts:
            ts.append(ts[-1] + h)
        info = _new_info()
        info.update(num_f_evaluations=len(dts), num_df_evaluations=len(dts), num_steps=len(dts),
                    num_attempted_steps=len(dts))
        means = torch.cat([state0.y.mean[None], out["mean_traj"][:, 0]])
        stds = torch.cat([std0, out["std_traj"][:, 0]])
        return marginals.MarginalSolution(t=np.asarray(ts), mean=means, std=stds, info=info,
                                          diffusion_squared_calibrated=out["diff_sum"][0] / len(dts))

    def solution_generator(self, pde, /, *, stop_at=None, progressbar=False):
        """Generate solver steps (pdefilter.py:118-165)."""
        time_stopper = self._process_event_inputs(stop_at_locations=stop_at)
        state = self.initialize(pde)
        info = _new_info()
        yield state, info
        dt = self.steprule.first_dt(pde)
        progressbar_steps = 100
        threshold = increment = pde.tmax / progressbar_steps
        pbar = tqdm(total=progressbar_steps) if progressbar else None
        while state.t < pde.tmax:
            if pbar is not None:
                while state.t + dt >= threshold:
                    pbar.update()
                    threshold += increment
                pbar.set_description(f"t={state.t:.4f}, dt={dt:.2E}")
            if time_stopper is not None:
                dt = time_stopper.adjust_dt_to_time_stops(state.t, dt)
            state, dt, step_info = self.perform_full_step(state, dt, pde)
            info["num_steps"] += 1
            for key in ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations", "num_attempted_steps"):
                info[key] += step_info[key]
            yield state, info
        if pbar is not None:
            pbar.update()
            pbar.set_description(f"t={state.t:.4f}, dt={dt:.2E}")
            pbar.close()

    @staticmethod
    def _process_event_inputs(stop_at_locations):
        return _TimeStopper(stop_at_locations) if stop_at_locations is not None else None

    def perform_full_step(self, state, initial_dt, pde):
        """One accepted step, including accept/reject by the step rule (pdefilter.py:177-227)."""
        dt = initial_dt
        accepted = False
        proposed = None
        step_info = dict(num_f_evaluations=0, num_df_evaluations=0, num_df_diagonal_evaluations=0, num_attempted_steps=0)
        while not accepted:
            proposed, attempt_info = self.attempt_step(state, dt, pde)
            step_info["num_attempted_steps"] += 1
            for key in ("num_f_evaluations", "num_df_evaluations", "num_df_diagonal_evaluations"):
                step_info[key] += attempt_info.get(key, 0)
            internal_norm = self.steprule.scale_error_estimate(
                unscaled_error_estimate=dt * proposed.error_estimate if proposed.error_estimate is not None else None,
                reference_state=proposed.reference_state)
            accepted = self.steprule.is_accepted(internal_norm)
            suggested = self.steprule.suggest(dt, internal_norm, local_convergence_rate=self.num_derivatives + 1)
            dt = min(suggested, pde.tmax - (proposed.t if accepted else state.t))
            assert dt >= 0, f"Invalid step size: dt={dt}"
        return proposed, dt, step_info

    # ------------------------------------------------------------------ engine plumbing
    family = None

    def _gram_sqrtm(self, pde):
        X = pde.mesh_spatial.points
        return np.linalg.cholesky(self.spatial_kernel(X, X.T))  # white.py:85, latent.py:139

    def _make_engine(self, pde, gram_sqrtm, **members):
        return _engine.Engine(pde, family=self.family, num_derivatives=self.num_derivatives, gram_sqrtm=gram_sqrtm,
                              **members)

    def _engine_for(self, pde):
        if self._engine is None or self._engine_pde is not pde:
            self.initialize(pde)
        return self._engine

    @abstractmethod
    def initialize(self, pde):
        raise NotImplementedError

    @abstractmethod
    def attempt_step(self, state, dt, pde):
        raise NotImplementedError


def _mark_tril(t):
    t._pnmol_b200_tril = True
    return t


def _factor_flags(chol):
    if getattr(chol, "_pnmol_b200_tril", False):
        return 0
    dense = bool(torch.triu(chol, diagonal=1).ne(0).any().item())
    return _engine.FLAG_DENSE_FACTOR if dense else 0


class _TimeStopper:
    """Make the solver stop at specified time points (pdefilter.py:238-256)."""

    def __init__(self, locations: Iterable):
        self._locations = iter(locations)
        self._next_location = next(self._locations)

    def adjust_dt_to_time_stops(self, t, dt):
        if t >= self._next_location:
            try:
                self._next_location = next(self._locations)
            except StopIteration:
                self._next_location = np.inf
        if t + dt > self._next_location:
            dt = self._next_location - t
        return dt
