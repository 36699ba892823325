"""White-noise EK1 solvers (API of src/pnmol/white.py).

``initialize`` (white.py:12-80) and ``attempt_step`` (white.py:96-146) are single CUDA
launches; the linear / semi-linear difference (white.py:169-208) is the reaction functor
evaluated inside the kernel.
"""
from . import pdefilter
from .base import iwp, rv


class _WhiteNoiseEK1Base(pdefilter.PDEFilter):
    family = "white"

    def initialize(self, pde):
        self.iwp, self.E0, self.E1, gram_sqrtm = self.initialize_iwp(pde=pde)
        self._engine = self._make_engine(pde, gram_sqrtm)
        self._engine_pde = pde
        mean, chol, _ = self._engine.initialize(pde.y0, pde.t0, self.diffuse_prior_scale)
        y = rv.MultivariateNormal(mean=mean[0], cov_sqrtm=pdefilter._mark_tril(chol[0]))
        return pdefilter.PDEFilterState(t=pde.t0, y=y, error_estimate=None, reference_state=None,
                                        diffusion_squared_local=[])

    def initialize_iwp(self, pde):
        """white.py:82-94."""
        gram_sqrtm = self._gram_sqrtm(pde)
        prior = iwp.IntegratedWienerTransition(num_derivatives=self.num_derivatives,
                                               wiener_process_dimension=pde.y0.shape[0], wp_diffusion_sqrtm=gram_sqrtm)
        return prior, prior.projection_matrix(0), prior.projection_matrix(1), gram_sqrtm

    def attempt_step(self, state, dt, pde):
        eng = self._engine_for(pde)
        flags = pdefilter._factor_flags(state.y.cov_sqrtm)
        mean, chol, err, ref, diff, _ = eng.step(state.t, dt, state.y.mean, state.y.cov_sqrtm, flags)
        new_state = pdefilter.PDEFilterState(
            t=state.t + dt, error_estimate=err[0], reference_state=ref[0],
            y=rv.MultivariateNormal(mean[0], pdefilter._mark_tril(chol[0])), diffusion_squared_local=diff[0])
        return new_state, dict(num_f_evaluations=1, num_df_evaluations=1)


class LinearWhiteNoiseEK1(_WhiteNoiseEK1Base):
    """white.py:169-186."""


class SemiLinearWhiteNoiseEK1(_WhiteNoiseEK1Base):
    """white.py:189-208."""
