"""Ensembles of independent EK1 solves (the batch axis the GPU path is built for).

The reference solves one problem per Python call; its only batching primitives are the
unused ``jax.vmap`` wrappers of src/pnmol/base/sqrt.py:27-30.  An ensemble here is ``B``
members of one discretised problem that differ in their initial condition, their
diffusivity (which scales ``L`` and ``E_sqrtm``, src/pnmol/pde/mixins.py:37-38), the output
scale of the spatial prior kernel (which scales ``chol(k(X,X))``, src/pnmol/white.py:85) and
the reaction parameters.  Every member is an independent ``simulate_final_state``
(src/pnmol/pdefilter.py:105-116); one CTA owns one member for the whole time loop.

Multi-GPU: members are partitioned into contiguous slices, one per rank
(``torch.distributed``, one process per GPU); the time loop needs no communication and
the final states are gathered once at the end.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _engine
from .odetools import step as _step

EnsembleResult = namedtuple("EnsembleResult", "t mean cov_sqrtm diffusion_squared_calibrated status num_steps")


def member_slice(num_members, world_size, rank):
    """Contiguous, balanced partition of ``num_members`` over ``world_size`` ranks."""
    base, extra = divmod(num_members, world_size)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def _take(x, sl, batch):
    if x is None:
        return None
    x = np.asarray(x, dtype=np.float64)
    return x.reshape(batch, -1)[sl]


class EnsembleSolver:
    """Binds a solver (its kind, order, prior kernel, constant step) to an ensemble of one problem."""

    def __init__(self, solver, pde, *, y0=None, diff_scale=None, prior_scale=None, reaction_params=None, device=None):
        self.adaptive = isinstance(solver.steprule, _step.Adaptive)
        if self.adaptive and (solver.family != "white" or getattr(pde, "is_semilinear", False)):
            raise NotImplementedError("adaptive ensembles: linear white-noise solvers (per-member first_dt needs L y0)")
        y0 = np.asarray(pde.y0 if y0 is None else y0, dtype=np.float64)
        self.y0 = np.ascontiguousarray(y0.reshape(-1, pde.y0.shape[0]))
        self.batch = self.y0.shape[0]
        self.solver, self.pde = solver, pde
        if self.adaptive:
            # Adaptive.first_dt per member (odetools/step.py:103-133): 0.01 ||y0|| / ||diff_scale L y0||
            ds = np.ones((self.batch, 1)) if diff_scale is None else np.asarray(diff_scale, dtype=np.float64).reshape(self.batch, -1)
            ncomp = getattr(pde, "num_components", 1)
            rows = np.repeat(np.broadcast_to(ds, (self.batch, ncomp)), pde.L.shape[0] // ncomp, axis=1)
            dy0 = rows * (self.y0 @ np.asarray(pde.L, dtype=np.float64).T)
            self.dt0 = 0.01 * np.linalg.norm(self.y0, axis=1) / np.linalg.norm(dy0, axis=1)
            self.dts = np.zeros(0)
        else:
            self.dts = _engine.constant_step_schedule(pde.t0, pde.tmax, solver.steprule.first_dt(pde))
        self.engine = _engine.Engine(pde, family=solver.family, num_derivatives=solver.num_derivatives,
                                     gram_sqrtm=solver._gram_sqrtm(pde), batch=self.batch, device=device,
                                     diff_scale=diff_scale, prior_scale=prior_scale, reaction_params=reaction_params)

    @property
    def t_final(self):
        t = self.pde.t0
        for h in self.dts:
            t = t + h
        return t

    def initialize(self, y0_device=None):
        y0 = self.y0 if y0_device is None else y0_device
        return self.engine.initialize(y0, self.pde.t0, self.solver.diffuse_prior_scale)

    def simulate_final_state(self, *, rescale=True, flags=0):
        """Device-resident route: returns CUDA tensors."""
        mean, chol, status0 = self.initialize()
        if self.adaptive:  # every member with its own step sizes, accept/reject on the device
            out = self.engine.run_adaptive(self.pde.t0, self.pde.tmax, self.dt0, self.solver.steprule, mean, chol, flags=flags)
            cal = out["diff_sum"] / out["num_steps"]
            if rescale:
                self.engine.rescale(chol, cal, 1)
            res = EnsembleResult(out["t"], mean, chol, cal, torch.maximum(status0, out["status"]), out["num_steps"])
            self.last_adaptive = out
            return res
        out = self.engine.run(self.pde.t0, self.dts, mean, chol, flags=flags)
        if rescale:
            cal = self.engine.rescale(chol, out["diff_sum"], len(self.dts))
        else:
            cal = out["diff_sum"] / len(self.dts)
        return EnsembleResult(self.t_final, mean, chol, cal, torch.maximum(status0, out["status"]), len(self.dts))

    def simulate_final_state_host(self, *, mean_host=None, chol_host=None, flags=0):
        """Host-buffer route through ``pnmol_b200_simulate_final_state_host``: host y0 in, host results out."""
        mean, chol, cal, status = self.engine.simulate_final_state_host(
            self.y0, self.pde.t0, self.solver.diffuse_prior_scale, self.dts, mean_host=mean_host, chol_host=chol_host,
            flags=flags)
        return EnsembleResult(self.t_final, mean, chol, cal, status, len(self.dts))


def simulate_final_state(solver, pde, **members):
    """One-call ensemble version of ``solver.simulate_final_state(pde)``."""
    return EnsembleSolver(solver, pde, **members).simulate_final_state()


def simulate_final_state_distributed(solver, pde, *, y0, diff_scale=None, prior_scale=None, reaction_params=None,
                                     gather=True, compute=None):
    """Shard the members over the ranks of the default ``torch.distributed`` group and gather the
    final means (and factors) on every rank.

    ``compute(solver, pde, **members) -> EnsembleResult`` defaults to the CUDA ensemble; the
    hook exists so that the sharding / gathering logic can be exercised with the ``gloo``
    backend on CPU (tests/test_ensemble_sharding.py)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    y0 = np.asarray(y0, dtype=np.float64)
    B = y0.shape[0]
    sl = member_slice(B, world, rank)
    compute = compute or simulate_final_state
    local = compute(solver, pde, y0=y0[sl], diff_scale=_take(diff_scale, sl, B), prior_scale=_take(prior_scale, sl, B),
                    reaction_params=_take(reaction_params, sl, B))
    if not gather:
        return local
    sizes = [member_slice(B, world, r) for r in range(world)]
    counts = [s.stop - s.start for s in sizes]

    def allgather(x):
        # ranks may own one member more or less: pad to the largest slice, gather, trim
        cmax = max(counts)
        pad = x.new_zeros((cmax,) + tuple(x.shape[1:]))
        pad[: x.shape[0]] = x
        out = x.new_empty((world * cmax,) + tuple(x.shape[1:]))
        if x.is_cuda:
            dist.all_gather_into_tensor(out, pad)
        else:
            dist.all_gather(list(out.chunk(world)), pad)
        if len(set(counts)) == 1:
            return out
        return torch.cat([out[r * cmax: r * cmax + c] for r, c in enumerate(counts)])

    return EnsembleResult(local.t, allgather(local.mean), allgather(local.cov_sqrtm),
                          allgather(local.diffusion_squared_calibrated), allgather(local.status), local.num_steps)
