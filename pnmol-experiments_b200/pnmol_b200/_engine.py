"""Host-side owner of one libpnmol_b200 handle: marshals a discretised problem into the
C ABI (ELL stencils, IWP constants, per-member ensemble axes) and exposes
initialize / step / run on torch CUDA float64 tensors."""
import ctypes

import numpy as np
import torch

from . import _lib

KINDS = {("white", False): 0, ("white", True): 1, ("latent", False): 2, ("latent", True): 3}
FLAG_DENSE_FACTOR = 1
FLAG_NO_ERROR_ESTIMATE = 2


def to_ell(M):
    """Dense (rows, cols) -> (col int32 [rows, w], val float64 [rows, w]), col = -1 padding."""
    M = np.asarray(M, dtype=np.float64)
    nz = M != 0.0
    w = max(1, int(nz.sum(axis=1).max())) if M.size else 1
    col = -np.ones((M.shape[0], w), dtype=np.int32)
    val = np.zeros((M.shape[0], w), dtype=np.float64)
    for r in range(M.shape[0]):
        idx = np.nonzero(nz[r])[0]
        col[r, :len(idx)] = idx
        val[r, :len(idx)] = M[r, idx]
    return np.ascontiguousarray(col), np.ascontiguousarray(val)


def nordsieck_raw(num_derivatives, dt):
    """src/pnmol/base/iwp.py:55-62, evaluated on the host with NumPy's pow."""
    import scipy.special

    powers = np.arange(num_derivatives, -1, -1)
    scales = scipy.special.factorial(powers)
    powers = powers + 0.5
    return (np.abs(dt) ** powers) / scales, (np.abs(dt) ** (-powers)) * scales


def constant_step_schedule(t0, tmax, dt):
    """Step sizes of solution_generator + perform_full_step under step.Constant
    (src/pnmol/pdefilter.py:140,220-223), including the floating-point sliver step."""
    dts, t, h = [], t0, dt
    while t < tmax:
        dts.append(h)
        t = t + h
        h = min(dt, tmax - t)
        assert h >= 0, f"Invalid step size: dt={h}"
    return np.asarray(dts, dtype=np.float64)


class Engine:
    def __init__(self, pde, *, family, num_derivatives, gram_sqrtm, batch=1, device=None, diff_scale=None,
                 prior_scale=None, reaction_params=None):
        if not torch.cuda.is_available():
            raise _lib.PnmolB200Error("pnmol_b200 needs a CUDA device: the EK1 path has no CPU fallback")
        self.lib = _lib.load()
        dev = torch.device("cuda") if device is None else torch.device(device)
        # normalise once: torch.device("cuda") has index None and means the CURRENT device, not device 0
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        semil = bool(getattr(pde, "is_semilinear", False)) and pde.reaction is not None
        if getattr(pde, "is_semilinear", False) and pde.reaction is None:
            raise NotImplementedError("semi-linear problems need a device Reaction tag (pde.reaction); arbitrary "
                                      "host callables f/df cannot be called from the CUDA step")
        self.semilinear = semil
        self.family = family
        self.kind = KINDS[(family, semil)]
        L = np.asarray(pde.L, dtype=np.float64)
        E = np.asarray(pde.E_sqrtm, dtype=np.float64)
        if np.any(E - np.diag(np.diag(E)) != 0.0):
            raise NotImplementedError("only diagonal E_sqrtm (probabilistic FD discretisations) is supported")
        B = np.asarray(pde.B, dtype=np.float64)
        R = np.ascontiguousarray(np.asarray(pde.R_sqrtm, dtype=np.float64))
        self.d, self.nb = L.shape[0], B.shape[0]
        self.nu, self.n = num_derivatives, num_derivatives + 1
        self.ncomp = getattr(pde, "num_components", 1)
        self.dd = 2 * self.d if family == "latent" else self.d
        self.D, self.m = self.n * self.dd, self.d + self.nb
        self.batch = batch
        reaction_id = pde.reaction.id if semil else 0
        h = ctypes.c_void_p()
        _lib.check(self.lib.pnmol_b200_create(ctypes.byref(h), self.kind, self.d, self.nu, self.nb, self.ncomp, batch,
                                              reaction_id, self.device.index))
        self.h = h
        lcol, lval = to_ell(L)
        bcol, bval = to_ell(B)
        ediag = np.ascontiguousarray(np.diag(E))
        _lib.check(self.lib.pnmol_b200_set_operator(h, _lib.ptr(lcol), _lib.ptr(lval), lcol.shape[1], _lib.ptr(ediag),
                                                    _lib.ptr(bcol), _lib.ptr(bval), bcol.shape[1], _lib.ptr(R)))
        from .base import iwp as _iwp

        A1d, LQ1d = _iwp.IntegratedWienerTransition(1, num_derivatives, np.eye(1)).preconditioned_discretize_1d
        Lk = np.ascontiguousarray(np.asarray(gram_sqrtm, dtype=np.float64))
        _lib.check(self.lib.pnmol_b200_set_prior(h, _lib.ptr(np.ascontiguousarray(A1d)),
                                                 _lib.ptr(np.ascontiguousarray(LQ1d)), _lib.ptr(Lk)))
        rp = None
        nparams = 0
        if semil:
            base = np.asarray(pde.reaction.params, dtype=np.float64)
            rp = np.tile(base, (batch, 1)) if reaction_params is None else np.asarray(reaction_params, dtype=np.float64)
            rp = np.ascontiguousarray(rp.reshape(batch, -1))
            nparams = rp.shape[1]
        ds = None if diff_scale is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(diff_scale, dtype=np.float64).reshape(batch, -1), (batch, self.ncomp)))
        ps = None if prior_scale is None else np.ascontiguousarray(np.asarray(prior_scale, dtype=np.float64).reshape(batch))
        if ds is not None or ps is not None or rp is not None:
            _lib.check(self.lib.pnmol_b200_set_members(h, _lib.ptr(ds), _lib.ptr(ps), _lib.ptr(rp), nparams))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.pnmol_b200_destroy(h)
            except Exception:
                pass

    @property
    def path(self):
        """"small" (one warp per member, workspace in shared memory), "single_cta" (one CTA per member) or "multi_cta"
        (whole grid per member)."""
        rc = self.lib.pnmol_b200_path(self.h)
        if rc < 0:
            _lib.check(rc)
        return {0: "single_cta", 1: "multi_cta", 2: "small"}[rc]

    def cluster_size(self):
        """(requested, observed) thread-block cluster size of the multi-CTA kernels (diagnostics; (0, 0) elsewhere)."""
        import ctypes

        req, obs = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(self.lib.pnmol_b200_cluster_size(self.h, ctypes.byref(req), ctypes.byref(obs)))
        return req.value, obs.value

    # ------------------------------------------------------------------ helpers
    def _empty(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _stream(self):
        return _lib.current_stream(self.device)

    def _precond(self, dts):
        pv = np.stack([nordsieck_raw(self.nu, dt)[0] for dt in dts])
        pinv = np.stack([nordsieck_raw(self.nu, dt)[1] for dt in dts])
        return np.ascontiguousarray(pv), np.ascontiguousarray(pinv)

    # ------------------------------------------------------------------ compute
    def initialize(self, y0, t0, diffuse_prior_scale):
        y0 = torch.as_tensor(y0, dtype=torch.float64, device=self.device).reshape(self.batch, self.d).contiguous()
        mean = self._empty(self.batch, self.n, self.dd)
        chol = self._empty(self.batch, self.D, self.D)
        status = self._empty(self.batch, dtype=torch.int32)
        _lib.check(self.lib.pnmol_b200_initialize(self.h, _lib.ptr(y0), float(t0), float(diffuse_prior_scale),
                                                  _lib.ptr(mean), _lib.ptr(chol), _lib.ptr(status), self._stream()))
        return mean, chol, status

    def step(self, t, dt, mean, chol, flags=0):
        pv, pinv = nordsieck_raw(self.nu, dt)
        pv, pinv = np.ascontiguousarray(pv), np.ascontiguousarray(pinv)
        mean = mean.reshape(self.batch, self.n, self.dd).contiguous()
        chol = chol.reshape(self.batch, self.D, self.D).contiguous()
        mean_out, chol_out = torch.empty_like(mean), torch.empty_like(chol)
        white = self.family == "white"
        err = self._empty(self.batch, self.d) if white else None
        ref = self._empty(self.batch, self.d) if white else None
        diff = self._empty(self.batch)
        status = self._empty(self.batch, dtype=torch.int32)
        _lib.check(self.lib.pnmol_b200_step(self.h, float(t + dt), float(dt), _lib.ptr(pv), _lib.ptr(pinv), _lib.ptr(mean),
                                            _lib.ptr(chol), _lib.ptr(mean_out), _lib.ptr(chol_out), _lib.ptr(err),
                                            _lib.ptr(ref), _lib.ptr(diff), _lib.ptr(status), int(flags), self._stream()))
        return mean_out, chol_out, err, ref, diff, status

    def run(self, t0, dts, mean, chol, *, trajectory=False, flags=0):
        """In-place multi-step run; returns dict with diff_sum, diff_last, err, ref, status (+ trajectories)."""
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float64))
        T = len(dts)
        pv, pinv = self._precond(dts)
        assert mean.is_contiguous() and chol.is_contiguous()
        mean_tmp, chol_tmp = torch.empty_like(mean), torch.empty_like(chol)
        white = self.family == "white"
        out = dict(err=self._empty(self.batch, self.d) if white else None,
                   ref=self._empty(self.batch, self.d) if white else None,
                   diff_last=self._empty(self.batch), diff_sum=self._empty(self.batch),
                   status=self._empty(self.batch, dtype=torch.int32),
                   mean_traj=self._empty(T, self.batch, self.n, self.dd) if trajectory else None,
                   chol_traj=self._empty(T, self.batch, self.D, self.D) if trajectory else None)
        _lib.check(self.lib.pnmol_b200_run(self.h, float(t0), _lib.ptr(dts), _lib.ptr(pv), _lib.ptr(pinv), T, _lib.ptr(mean),
                                           _lib.ptr(chol), _lib.ptr(mean_tmp), _lib.ptr(chol_tmp), _lib.ptr(out["err"]),
                                           _lib.ptr(out["ref"]), _lib.ptr(out["diff_last"]), _lib.ptr(out["diff_sum"]),
                                           _lib.ptr(out["mean_traj"]), _lib.ptr(out["chol_traj"]), _lib.ptr(out["status"]),
                                           int(flags), self._stream()))
        out["_keepalive"] = (mean_tmp, chol_tmp)
        return out

    def run_marginals(self, t0, dts, mean, chol, *, flags=0):
        """Multi-step run that returns the mean trajectory and the marginal standard deviations of the 0th derivative
        (fused read-out: the factor trajectory is never written).  mean/chol are updated in place to the final state."""
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float64))
        T = len(dts)
        pv, pinv = self._precond(dts)
        assert mean.is_contiguous() and chol.is_contiguous()
        mean_tmp, chol_tmp = torch.empty_like(mean), torch.empty_like(chol)
        out = dict(diff_last=self._empty(self.batch), diff_sum=self._empty(self.batch),
                   status=self._empty(self.batch, dtype=torch.int32),
                   mean_traj=self._empty(T, self.batch, self.n, self.dd), std_traj=self._empty(T, self.batch, self.dd))
        _lib.check(self.lib.pnmol_b200_run_marginals(
            self.h, float(t0), _lib.ptr(dts), _lib.ptr(pv), _lib.ptr(pinv), T, _lib.ptr(mean), _lib.ptr(chol),
            _lib.ptr(mean_tmp), _lib.ptr(chol_tmp), _lib.ptr(out["diff_last"]), _lib.ptr(out["diff_sum"]),
            _lib.ptr(out["mean_traj"]), _lib.ptr(out["std_traj"]), _lib.ptr(out["status"]), int(flags), self._stream()))
        out["_keepalive"] = (mean_tmp, chol_tmp)
        return out

    def run_adaptive(self, t0, tmax, dt0, rule, mean, chol, *, max_attempts=100000, flags=0, trajectory=0):
        """Adaptive time loop on the device (white-noise solvers, CTA-per-member and small-state paths): mean/chol are
        advanced in place from t0 to tmax, every member with its own step sizes.  `rule` is a step.Adaptive; dt0 is
        [batch].  trajectory = capacity (accepted steps per member) of the optional trajectory output
        (t_traj [batch, cap], mean_traj [cap, batch, n, dd], chol_traj [cap, batch, D, D]; status bit 4 = truncated)."""
        assert mean.is_contiguous() and chol.is_contiguous()
        dt0 = torch.as_tensor(np.broadcast_to(np.asarray(dt0, dtype=np.float64), (self.batch,)).copy(), device=self.device)
        mean_tmp, chol_tmp = torch.empty_like(mean), torch.empty_like(chol)
        out = dict(t=self._empty(self.batch), dt=self._empty(self.batch), diff_sum=self._empty(self.batch),
                   diff_last=self._empty(self.batch), num_steps=self._empty(self.batch, dtype=torch.int32),
                   num_attempts=self._empty(self.batch, dtype=torch.int32), status=self._empty(self.batch, dtype=torch.int32),
                   err=self._empty(self.batch, self.d), ref=self._empty(self.batch, self.d))
        cap = int(trajectory)
        if cap > 0:
            out.update(t_traj=self._empty(self.batch, cap), mean_traj=self._empty(cap, self.batch, self.n, self.dd),
                       chol_traj=self._empty(cap, self.batch, self.D, self.D))
        small, large = rule.max_changes
        _lib.check(self.lib.pnmol_b200_run_adaptive_trajectory(
            self.h, float(t0), float(tmax), _lib.ptr(dt0), float(rule.abstol), float(rule.reltol), float(small), float(large),
            float(rule.safety_scale), int(max_attempts), _lib.ptr(mean), _lib.ptr(chol), _lib.ptr(mean_tmp), _lib.ptr(chol_tmp),
            _lib.ptr(out["t"]), _lib.ptr(out["dt"]), _lib.ptr(out["diff_sum"]), _lib.ptr(out["diff_last"]),
            _lib.ptr(out["num_steps"]), _lib.ptr(out["num_attempts"]), _lib.ptr(out["status"]), _lib.ptr(out["err"]),
            _lib.ptr(out["ref"]), _lib.ptr(out.get("t_traj")), _lib.ptr(out.get("mean_traj")), _lib.ptr(out.get("chol_traj")),
            cap, int(flags), self._stream()))
        out["_keepalive"] = (mean_tmp, chol_tmp, dt0)
        return out

    def marginal_std(self, chol):
        """sqrt(diag(E0 L L^T E0^T)) for factors of shape (..., D, D) -> (..., dd)."""
        return marginal_std(chol, self.nu)

    def rescale(self, chol, diff_sum, nsteps):
        cal = self._empty(self.batch)
        _lib.check(self.lib.pnmol_b200_rescale(self.h, _lib.ptr(chol), _lib.ptr(diff_sum), int(nsteps), _lib.ptr(cal),
                                               self._stream()))
        return cal

    def simulate_final_state_host(self, y0_host, t0, diffuse_prior_scale, dts, *, mean_host=None, chol_host=None, flags=0):
        """HOST buffers in, HOST buffers out (pinned if the caller pinned them)."""
        dts = np.ascontiguousarray(np.asarray(dts, dtype=np.float64))
        pv, pinv = self._precond(dts)
        y0_host = _as_host(y0_host, (self.batch, self.d))
        if mean_host is None:
            mean_host = torch.empty((self.batch, self.n, self.dd), dtype=torch.float64).pin_memory()
        if chol_host is None:
            chol_host = torch.empty((self.batch, self.D, self.D), dtype=torch.float64).pin_memory()
        cal = torch.empty(self.batch, dtype=torch.float64)
        status = torch.empty(self.batch, dtype=torch.int32)
        _lib.check(self.lib.pnmol_b200_simulate_final_state_host(
            self.h, _lib.ptr(y0_host), float(t0), float(diffuse_prior_scale), _lib.ptr(dts), _lib.ptr(pv), _lib.ptr(pinv),
            len(dts), _lib.ptr(mean_host), _lib.ptr(chol_host), _lib.ptr(cal), _lib.ptr(status), int(flags), self._stream()))
        return mean_host, chol_host, cal, status


def marginal_std(chol, num_derivatives):
    """Marginal standard deviations of the 0th derivative: norm of every (nu+1)-th row of the factor (device kernel)."""
    chol = torch.as_tensor(chol, dtype=torch.float64)
    if not chol.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.PnmolB200Error("pnmol_b200 needs a CUDA device: the read-out has no CPU fallback")
        chol = chol.cuda()
    chol = chol.contiguous()
    D = chol.shape[-1]
    n = num_derivatives + 1
    lead = chol.shape[:-2]
    count = int(np.prod(lead)) if len(lead) else 1
    out = torch.empty((count, D // n), dtype=torch.float64, device=chol.device)
    _lib.check(_lib.load().pnmol_b200_marginal_std(_lib.ptr(chol), _lib.ptr(out), D, int(num_derivatives), count,
                                                   chol.device.index if chol.device.index is not None else torch.cuda.current_device(),
                                                   _lib.current_stream(chol.device)))
    return out.reshape(*lead, D // n)


def _as_host(x, shape):
    t = torch.as_tensor(x, dtype=torch.float64)
    if t.is_cuda:
        raise ValueError("host buffer expected")
    return t.reshape(shape).contiguous()
