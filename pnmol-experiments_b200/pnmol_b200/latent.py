"""Latent-force EK1 solvers (API of src/pnmol/latent.py).

The filter state stacks the PDE state IWP and an error-process IWP whose diffusion is
``pde.E_sqrtm`` (latent.py:136-153); the measurement update is noise-free
(latent.py:197) and there is no error estimate (latent.py:217-223), so these solvers run
with constant steps only, like the reference.
"""
from . import pdefilter
from .base import iwp, rv, stacked_ssm


class _LatentForceEK1Base(pdefilter.PDEFilter):
    family = "latent"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.ssm = None
        self.state_iwp = None
        self.lf_iwp = None

    def initialize(self, pde):
        self.state_iwp, self.lf_iwp, self.E0, self.E1, gram_sqrtm = self.initialize_iwp_latent(pde=pde)
        self.ssm = stacked_ssm.StackedSSM(processes=[self.state_iwp, self.lf_iwp])
        self._engine = self._make_engine(pde, gram_sqrtm)
        self._engine_pde = pde
        mean, chol, _ = self._engine.initialize(pde.y0, pde.t0, self.diffuse_prior_scale)
        y = rv.MultivariateNormal(mean=mean[0], cov_sqrtm=pdefilter._mark_tril(chol[0]))
        return pdefilter.PDEFilterState(t=pde.t0, y=y, error_estimate=None, reference_state=None,
                                        diffusion_squared_local=[])

    def initialize_iwp_latent(self, pde):
        """latent.py:136-153."""
        gram_sqrtm = self._gram_sqrtm(pde)
        d = pde.y0.shape[0]
        prior_state = iwp.IntegratedWienerTransition(num_derivatives=self.num_derivatives, wiener_process_dimension=d,
                                                     wp_diffusion_sqrtm=gram_sqrtm)
        prior_latent = iwp.IntegratedWienerTransition(num_derivatives=self.num_derivatives, wiener_process_dimension=d,
                                                      wp_diffusion_sqrtm=pde.E_sqrtm)
        return prior_state, prior_latent, prior_latent.projection_matrix(0), prior_latent.projection_matrix(1), gram_sqrtm

    def attempt_step(self, state, dt, pde):
        eng = self._engine_for(pde)
        flags = pdefilter._factor_flags(state.y.cov_sqrtm)
        mean, chol, _, _, diff, _ = eng.step(state.t, dt, state.y.mean, state.y.cov_sqrtm, flags)
        new_state = pdefilter.PDEFilterState(
            t=state.t + dt, error_estimate=None, reference_state=None,
            y=rv.MultivariateNormal(mean[0], pdefilter._mark_tril(chol[0])), diffusion_squared_local=diff[0])
        return new_state, dict(num_f_evaluations=1, num_df_evaluations=1)


class LinearLatentForceEK1(_LatentForceEK1Base):
    """latent.py:237-263."""


class SemiLinearLatentForceEK1(_LatentForceEK1Base):
    """latent.py:266-292."""
