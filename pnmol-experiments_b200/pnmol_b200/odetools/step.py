"""Step-size selection (src/pnmol/odetools/step.py).

``Constant`` (step.py:30-55) is what every configuration of the hot path uses.
``Adaptive`` (step.py:58-119): ``simulate_final_state`` of the white-noise solvers and adaptive
ensembles evaluate it inside the persistent kernel (``pnmol_b200_run_adaptive``); ``solve`` and
``solution_generator`` (variable-length trajectories) use the host-side scalar logic below, one
``attempt_step`` launch per trial step.
"""
import numpy as np


class StepRule:
    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        raise NotImplementedError

    def is_accepted(self, scaled_error_estimate):
        raise NotImplementedError

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        raise NotImplementedError

    def first_dt(self, discretized_pde):
        raise NotImplementedError


class Constant(StepRule):
    def __init__(self, dt):
        self.dt = dt
        self.min_step, self.max_step = 1e-15, 1e15

    def __repr__(self):
        return f"{type(self).__name__}(dt={self.dt})"

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        return self.dt

    def is_accepted(self, scaled_error_estimate):
        return True

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        return None  # step.py:50-52: never used further

    def first_dt(self, discretized_pde):
        return self.dt


class Adaptive(StepRule):
    def __init__(self, abstol=1e-4, reltol=1e-2, max_changes=(0.2, 10.0), safety_scale=0.95, min_step=1e-15,
                 max_step=1e15):
        self.abstol, self.reltol = abstol, reltol
        self.max_changes, self.safety_scale = max_changes, safety_scale
        self.min_step, self.max_step = min_step, max_step

    def __repr__(self):
        return f"{type(self).__name__}(abstol={self.abstol}, reltol={self.reltol})"

    def suggest(self, previous_dt, scaled_error_estimate, local_convergence_rate=None):
        if local_convergence_rate is None:
            raise ValueError("Please provide a local convergence rate.")
        small, large = self.max_changes
        change = self.safety_scale * (1.0 / scaled_error_estimate) ** (1.0 / local_convergence_rate)
        return float(max(small, min(change, large))) * previous_dt

    def is_accepted(self, scaled_error_estimate):
        return bool(scaled_error_estimate < 1)

    def scale_error_estimate(self, unscaled_error_estimate, reference_state):
        err = _host(unscaled_error_estimate)
        ref = _host(reference_state)
        if err.shape != ref.shape:
            raise ValueError("Unscaled error estimate needs same shape as reference state.")
        ratio = err / (self.abstol + self.reltol * ref)
        return float(np.linalg.norm(ratio) / np.sqrt(ratio.size))

    def first_dt(self, discretized_pde):
        y0 = np.asarray(discretized_pde.y0)
        if getattr(discretized_pde, "is_semilinear", False):
            dy0 = discretized_pde.f(discretized_pde.t0, y0)  # step.py:122-126
        else:
            dy0 = discretized_pde.L @ y0  # step.py:129-133
        return 0.01 * np.linalg.norm(y0) / np.linalg.norm(dy0)


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
