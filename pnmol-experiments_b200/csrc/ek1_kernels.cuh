// __global__ entry points built from the device phases in ek1_device.cuh.
#pragma once
#include "ek1_device.cuh"

namespace pnmol {

#ifndef PNMOL_PACE_PER_QR
#define PNMOL_PACE_PER_QR 0
#endif
struct RunArgs {
    int nsteps, flags, final_in_b;
    double pv0[kMaxN], pinv0[kMaxN], dt0, tnew0;              // single-step parameters (nsteps == 1, by value)
    const double *dts, *tnew, *pv, *pinv;                     // multi-step parameters (device arrays)
    double *mean_a, *chol_a, *mean_b, *chol_b;                // ping-pong state; step 0 reads a
    double *err_out, *ref_out, *diff_last, *diff_sum, *mean_traj, *chol_traj;
    int32_t* status;
    double* std_traj;  // [nsteps][batch][dd] marginal standard deviations of the 0th derivative, or nullptr
};

// Fused marginal read-out (experiments/figure1.py:76-89, figure3.py:87-93: sqrt(diag(E0 L L^T E0^T))): the standard
// deviation of state component j is the norm of row j n of the factor.  Warps [w0, w0 + nw) share the rows.
__device__ __forceinline__ void marginal_std_rows(const double* __restrict__ chol, int D, int n, int dd, double* __restrict__ out,
                                                  int w0, int nw) {
    const int lane = threadIdx.x & 31;
    for (int j = w0; j < dd; j += nw) {
        const double* row = chol + (size_t)(j * n) * D;
        double a0 = 0.0, a1 = 0.0;
        int c = lane;
        for (; c + 32 < D; c += 64) { a0 = fma(row[c], row[c], a0); a1 = fma(row[c + 32], row[c + 32], a1); }
        if (c < D) a0 = fma(row[c], row[c], a0);
        const double ssq = warp_sum(a0 + a1);
        if (lane == 0) out[j] = sqrt(ssq);
    }
}

struct InitArgs {
    const double* y0;
    double t0, prior_scale0, nugget;
    double *mean_out, *chol_out;
    int32_t* status;
};

// Pace keeping (see k_run): thread 0 announces the CTA at the counter and waits until `target` CTAs have done so.
// Followed by a block barrier at the caller.
__device__ __forceinline__ void pace_wait(unsigned* gsync, unsigned target) {
    if (threadIdx.x == 0) {
        atomicAdd(gsync, 1u);
        while (*reinterpret_cast<volatile unsigned*>(gsync) < target) __nanosleep(100);
    }
}

// One EK1 step for member b: state (mean_in, chol_in) -> (mean_out, chol_out).
static __device__ void ek1_step(const Problem& P, int b, int slot, const Smem& sm, double dt, double tnew,
                         const double* mean_in, const double* chol_in, double* mean_out, double* chol_out,
                         double* err_out, double* ref_out, double* diff_out, int flags, int* nonfinite,
                         const double* pv_prev = nullptr, bool write_factor = true, unsigned* pace_target = nullptr,
                         unsigned pace_active = 0, bool err_inverse = false) {
    const int tid = threadIdx.x;
    PhaseClock pc;
    pc.start(P.prof);
    const int n = P.n, D = P.D;
    double* W = P.W + (size_t)slot * P.ld * (P.m + P.D);
    int32_t* Hcol = sm.Hcol ? sm.Hcol : P.Hcol + (size_t)slot * P.m * P.wh;
    double* Hval = sm.Hval ? sm.Hval : P.Hval + (size_t)slot * P.m * P.wh;
    // [setup + predict]  m = P^-1 mean (flattened column-major, index j n + i), mp = A m   white.py:104-107
    for (int k = tid; k < D; k += kThreads) {
        const int j = k / n, i = k - j * n;
        double acc = 0.0;
        for (int s = 0; s < n; ++s) acc = fma(P.A1d[i * n + s], sm.pinv[s] * mean_in[(size_t)s * P.dd + j], acc);
        sm.mp[k] = acc;
    }
    __syncthreads();
    evaluate_ode(P, b, sm, sm.pv[0], sm.pv[1], Hcol, Hval);
    pc.mark(0);
    const bool dense = flags & 1;
    // (pv_prev: the previous step of this launch left its factor in W, see build_predict)
    build_predict(P, b, sm, chol_in, dense ? sm.te_pd : sm.te_p, W + (size_t)P.m * P.ld, threadIdx.x >> 5, kWarps, pv_prev,
                  P.D + (P.latent ? 0 : P.m));
    pc.mark(1);
    Shape sp;
    sp.nt = D; sp.nbot = D; sp.ncols = D; sp.te = dense ? sm.te_pd : sm.te_p; sp.be = sm.be_p; sp.ldr = P.ld;
    householder_qr_fast(W + (size_t)P.m * P.ld, P.ld, sp, sm.fq, pc);
#if PNMOL_PACE_PER_QR
    if (pace_target) {   // second pace point of the step (the first is at the end of the step, in k_run)
        *pace_target += pace_active;
        pace_wait(P.gsync, *pace_target);
        __syncthreads();
    }
#endif
    pc.mark(2);
    // small m: the error estimate's forward solve is deferred into update_stage (one warp, next to the build of the
    // update matrix); its assembly / cached factor is set up here by all threads
    const bool est = !P.latent && !(flags & 2);
    const bool defer = est && P.ldm > 0 && P.m <= 96;
    if (est) {
        error_estimate(P, b, sm, sm.pv[1], dt, E_STEP_WHITE, 0.0, Hcol, Hval, P.F + (size_t)slot * P.m * P.d,
                       P.S + (size_t)slot * P.m * P.m, err_out, defer, err_inverse);
    }
    pc.mark(3);
    UpdateOut out;
    out.err_solve = defer; out.err_dt = dt; out.err_out = err_out;
    out.err_inverse = defer && err_inverse && !P.semilinear && err_inverse_fits(P);
    out.mean_out = mean_out; out.chol_out = write_factor ? chol_out : nullptr; out.diff_out = diff_out;
    out.ref_out = P.latent ? nullptr : ref_out; out.scale_by_p = true;
    update_stage(P, b, sm, P.m, P.latent ? E_NONE : E_STEP_WHITE, 0.0, nullptr, sm.te_u, sm.be_u, Hcol, Hval, W, out,
                 nonfinite, pc);
}

#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_RUN)
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_run(const Problem P, const RunArgs a) {
    extern __shared__ __align__(16) double smem_raw[];
    Smem sm = carve(smem_raw, P.D, P.m, P.dd, P.vld, P.ldm, P.whs);
    sm.fq.slot = sm_slot(P);
    __shared__ int nonfinite;
    __shared__ double diff_s;
    for (double* q = sm.fqbase + threadIdx.x; q < sm.fqend; q += kThreads) *q = 0.0;  // reflector buffers start finite
    if (threadIdx.x == 0) sm.ekey[0] = -1.0;  // no cached error-estimate factor yet
    load_envelopes(P, sm);
    const int tid = threadIdx.x;
    const size_t msz = (size_t)P.D, csz = (size_t)P.D * P.D;
    // Pace keeping (P.gsync != nullptr): after every step the CTAs of the grid wait for each other.  Members are
    // independent, so this is not needed for correctness; it keeps the CTAs in lock-step.  The step kernel's instruction
    // working set (hundreds of KB over a step) lives in instruction caches that are shared between SMs: while all
    // CTAs execute the same phase at the same time a line fetched by one SM serves the others; once they drift apart
    // every SM streams the whole kernel through the caches on its own and the step slows down by up to 40 %
    // (DESIGN.md section 3).  All CTAs of the grid are co-resident (grid <= CTAs per SM x SMs), so waiting cannot
    // deadlock; in round r only the min(grid, batch - r grid) CTAs that still have a member take part.
    unsigned pace_target = 0;
    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        if (tid == 0) nonfinite = 0;
        double diffsum = 0.0;
        const int first_of_round = b - (int)blockIdx.x;
        const unsigned pace_active = (unsigned)(P.batch - first_of_round < (int)gridDim.x ? P.batch - first_of_round : (int)gridDim.x);
        __syncthreads();
        for (int s = 0; s < a.nsteps; ++s) {
            double dt, tnew;
            if (a.nsteps == 1 && a.pv == nullptr) {
                if (tid < P.n) { sm.pv[tid] = a.pv0[tid]; sm.pinv[tid] = a.pinv0[tid]; }
                dt = a.dt0; tnew = a.tnew0;
            } else {
                if (tid < P.n) { sm.pv[tid] = a.pv[(size_t)s * P.n + tid]; sm.pinv[tid] = a.pinv[(size_t)s * P.n + tid]; }
                dt = a.dts[s]; tnew = a.tnew[s];
            }
            __syncthreads();
            const bool even = (s & 1) == 0;
            const double* min_ = (even ? a.mean_a : a.mean_b) + b * msz;
            const double* cin_ = (even ? a.chol_a : a.chol_b) + b * csz;
            double* mout = (even ? a.mean_b : a.mean_a) + b * msz;
            double* cout = (even ? a.chol_b : a.chol_a) + b * csz;
            // only the first step may see a user-supplied (possibly dense) factor
            const int flags = s == 0 ? a.flags : (a.flags & ~1);
            // Fused time loop: unless a factor trajectory / marginal read-out needs it, only the last step writes the
            // D x D factor to the state; the others hand it to their successor inside the workspace.
            const bool fuse = !a.chol_traj && !a.std_traj && a.pv != nullptr && !(a.flags & 4);
            const bool from_w = fuse && s > 0, to_w = fuse && s + 1 < a.nsteps;
            ek1_step(P, b, blockIdx.x, sm, dt, tnew, min_, cin_, mout, cout,
                     a.err_out ? a.err_out + (size_t)b * P.d : nullptr,
                     a.ref_out ? a.ref_out + (size_t)b * P.d : nullptr, &diff_s, flags, &nonfinite,
                     from_w ? a.pv + (size_t)(s - 1) * P.n : nullptr, !to_w,
                     (P.gsync && pace_active > 1) ? &pace_target : nullptr, pace_active,
                     /* cached inverse factor of the error estimate: */ a.nsteps > 2 && a.pv != nullptr);
            __syncthreads();
            diffsum += diff_s;
            if (a.mean_traj) {
                double* mt = a.mean_traj + ((size_t)s * P.batch + b) * msz;
                for (size_t k = tid; k < msz; k += kThreads) mt[k] = mout[k];
            }
            if (a.chol_traj) {
                double* ct = a.chol_traj + ((size_t)s * P.batch + b) * csz;
                for (size_t k = tid; k < csz; k += kThreads) ct[k] = cout[k];
            }
            if (a.std_traj) marginal_std_rows(cout, P.D, P.n, P.dd, a.std_traj + ((size_t)s * P.batch + b) * P.dd, tid >> 5, kWarps);
            if (P.gsync && pace_active > 1) {
                pace_target += pace_active;
                pace_wait(P.gsync, pace_target);
            }
            __syncthreads();
        }
        if ((a.nsteps & 1) && !a.final_in_b) {  // result sits in b: bring it home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = tid; k < msz; k += kThreads) md[k] = ms[k];
            for (size_t k = tid; k < csz; k += kThreads) cd[k] = cs[k];
        }
        if (tid == 0) {
            if (a.diff_last) a.diff_last[b] = diff_s;
            if (a.diff_sum) a.diff_sum[b] = diffsum;
            if (a.status) a.status[b] = nonfinite;
        }
        if (nonfinite) {  // rows outside the envelopes are read (times zero) by the tile-aligned trailing updates:
            double* Wz = P.W + (size_t)blockIdx.x * P.ld * (P.m + P.D);  // do not leave NaNs behind for the next member
            for (size_t k = tid; k < (size_t)P.ld * (P.m + P.D); k += kThreads) Wz[k] = 0.0;
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_run(const Problem P, const RunArgs a);
#endif

// Adaptive time loop on the device (src/pnmol/pdefilter.py:118-227 with src/pnmol/odetools/step.py:58-119): every
// member advances with its own step size, accept/reject and step-size proposal happen in the kernel.  White-noise
// solvers only (the latent-force solvers return no error estimate, src/pnmol/latent.py:217-223).
struct AdaptiveArgs {
    double t0, tmax, abstol, reltol, change_min, change_max, safety, inv_rate;
    const double* dt0;            // [batch] first step size (step.py:103-133, computed on the host)
    double *mean_a, *chol_a, *mean_b, *chol_b;   // current state in a; b is the proposal buffer
    double *err, *ref;            // [batch][d] scratch: error estimate and reference state of the proposal
    double *t_out, *dt_out, *diff_sum, *diff_last;   // [batch]
    int32_t *nsteps, *nattempts, *status;            // [batch]; status: 1 non-finite, 2 attempt limit reached, 4 trajectory truncated
    int max_attempts, flags;
    // optional trajectory of the accepted states (solve() with step.Adaptive): slot s = state after accepted step s + 1
    double *t_traj;                // [batch][max_traj]
    double *mean_traj, *chol_traj; // [max_traj][batch][D], [max_traj][batch][D * D]
    int max_traj;
};

#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_ADAPTIVE)
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_run_adaptive(const Problem P, const AdaptiveArgs a) {
    extern __shared__ __align__(16) double smem_raw[];
    Smem sm = carve(smem_raw, P.D, P.m, P.dd, P.vld, P.ldm, P.whs);
    sm.fq.slot = sm_slot(P);
    __shared__ int nonfinite;
    __shared__ double diff_s;
    for (double* q = sm.fqbase + threadIdx.x; q < sm.fqend; q += kThreads) *q = 0.0;  // reflector buffers start finite
    if (threadIdx.x == 0) sm.ekey[0] = -1.0;  // no cached error-estimate factor yet
    load_envelopes(P, sm);
    const int tid = threadIdx.x;
    const int nu = P.n - 1;
    const size_t msz = (size_t)P.D, csz = (size_t)P.D * P.D;
    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        if (tid == 0) nonfinite = 0;
        double t = a.t0, dt = a.dt0[b], diffsum = 0.0, difflast = 0.0;
        int nsteps = 0, natt = 0, cur = 0, stat = 0;
        double* err = a.err + (size_t)b * P.d;
        double* ref = a.ref + (size_t)b * P.d;
        __syncthreads();
        while (t < a.tmax) {
            if (natt >= a.max_attempts) { stat |= 2; break; }
            if (!(dt >= 0.0)) { stat |= 1; break; }  // pdefilter.py:225 asserts dt >= 0 (a NaN proposal ends here too)
            // Nordsieck preconditioner p_i = |dt|^(nu - i + 1/2) / (nu - i)!   (iwp.py:55-62)
            if (tid < P.n) {
                const int k = nu - tid;
                double fact = 1.0;
                for (int q = 2; q <= k; ++q) fact *= q;
                const double pw = pow(fabs(dt), k + 0.5);
                sm.pv[tid] = pw / fact;
                sm.pinv[tid] = fact / pw;
            }
            __syncthreads();
            const double* min_ = (cur ? a.mean_b : a.mean_a) + b * msz;
            const double* cin_ = (cur ? a.chol_b : a.chol_a) + b * csz;
            double* mout = (cur ? a.mean_a : a.mean_b) + b * msz;
            double* cout = (cur ? a.chol_a : a.chol_b) + b * csz;
            const int flags = (natt == 0 || nsteps == 0) ? a.flags : (a.flags & ~1);  // only a user-supplied factor may be dense
            ek1_step(P, b, blockIdx.x, sm, dt, t + dt, min_, cin_, mout, cout, err, ref, &diff_s, flags, &nonfinite);
            __syncthreads();
            // scaled error norm (step.py:97-108 on dt * error_estimate, pdefilter.py:208-213)
            double part = 0.0;
            for (int i = tid; i < P.d; i += kThreads) {
                const double r = dt * err[i] / (a.abstol + a.reltol * ref[i]);
                part = fma(r, r, part);
            }
            const double norm = sqrt(block_sum(part, sm.red)) / sqrt((double)P.d);
            double change = a.safety * pow(1.0 / norm, a.inv_rate);
            change = fmax(a.change_min, fmin(change, a.change_max));
            if (!(norm == norm)) change = norm;  // NaN propagates like jnp.minimum / jnp.maximum
            const double suggested = change * dt;
            ++natt;
            if (norm < 1.0) {  // accepted: the proposal becomes the state
                t = t + dt;
                cur ^= 1;
                ++nsteps;
                difflast = diff_s;
                diffsum += diff_s;
                if (a.mean_traj) {  // the accepted state joins the trajectory
                    if (nsteps <= a.max_traj) {
                        double* mt = a.mean_traj + ((size_t)(nsteps - 1) * P.batch + b) * msz;
                        double* ct = a.chol_traj + ((size_t)(nsteps - 1) * P.batch + b) * csz;
                        for (size_t k = tid; k < msz; k += kThreads) mt[k] = mout[k];
                        for (size_t k = tid; k < csz; k += kThreads) ct[k] = cout[k];
                        if (tid == 0) a.t_traj[(size_t)b * a.max_traj + nsteps - 1] = t;
                    } else {
                        stat |= 4;
                    }
                }
                dt = fmin(suggested, a.tmax - t);
                if (!(suggested == suggested)) dt = suggested;
            } else {
                dt = fmin(suggested, a.tmax - t);
                if (!(suggested == suggested)) dt = suggested;
                if (tid == 0) nonfinite = 0;  // a rejected proposal may be non-finite without harm
            }
            __syncthreads();
        }
        if (cur) {  // bring the final state home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = tid; k < msz; k += kThreads) md[k] = ms[k];
            for (size_t k = tid; k < csz; k += kThreads) cd[k] = cs[k];
        }
        __syncthreads();
        if (tid == 0) {
            a.t_out[b] = t; a.dt_out[b] = dt; a.diff_sum[b] = diffsum; a.diff_last[b] = difflast;
            a.nsteps[b] = nsteps; a.nattempts[b] = natt; a.status[b] = stat | (nonfinite ? 1 : 0);
        }
        if (nonfinite || stat) {
            double* Wz = P.W + (size_t)blockIdx.x * P.ld * (P.m + P.D);
            for (size_t k = tid; k < (size_t)P.ld * (P.m + P.D); k += kThreads) Wz[k] = 0.0;
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_run_adaptive(const Problem P, const AdaptiveArgs a);
#endif

// initialize(): two square-root updates on a Kronecker-structured prior factor.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_INIT)
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_init(const Problem P, const InitArgs a) {
    extern __shared__ __align__(16) double smem_raw[];
    Smem sm = carve(smem_raw, P.D, P.m, P.dd, P.vld, P.ldm, P.whs);
    sm.fq.slot = sm_slot(P);
    __shared__ int nonfinite;
    for (double* q = sm.fqbase + threadIdx.x; q < sm.fqend; q += kThreads) *q = 0.0;  // reflector buffers start finite
    if (threadIdx.x == 0) sm.ekey[0] = -1.0;  // no cached error-estimate factor yet
    load_envelopes(P, sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.n, d = P.d, D = P.D, nd = P.n * P.d;
    const int slot = blockIdx.x;
    double* W = P.W + (size_t)slot * P.ld * (P.m + P.D);
    int32_t* Hcol = sm.Hcol ? sm.Hcol : P.Hcol + (size_t)slot * P.m * P.wh;
    double* Hval = sm.Hval ? sm.Hval : P.Hval + (size_t)slot * P.m * P.wh;
    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        if (tid == 0) nonfinite = 0;
        double* chol = a.chol_out + (size_t)b * D * D;
        double* mean = a.mean_out + (size_t)b * D;
        const double ps = P.priorscale ? P.priorscale[b] : 1.0;
        // C0 = kron(Lk, c0 I_n) (white.py:23-24); latent: blockdiag(., kron(E_sqrtm, c0 I_n)) (latent.py:60-62,81-83)
        for (int r = warp; r < D; r += kWarps) {
            const int rb = r / n, ri = r - rb * n;
            for (int c = lane; c < D; c += 32) {
                const int cb = c / n, ci = c - cb * n;
                double v = 0.0;
                if (ri == ci && c <= r) {
                    if (r < nd) {
                        v = a.prior_scale0 * (ps * P.Lk[(size_t)rb * d + cb]);
                    } else if (rb == cb) {
                        const int comp = (rb - d) / P.npts;
                        const double ds = P.diffscale ? P.diffscale[(size_t)b * P.ncomp + comp] : 1.0;
                        v = a.prior_scale0 * (ds * P.Ediag[rb - d]);
                    }
                }
                chol[(size_t)r * D + c] = v;
            }
        }
        // update on the initial condition: H = E0, z = 0 - y0 so that m = K y0 (white.py:32-39)
        for (int k = tid; k < D; k += kThreads) sm.mp[k] = 0.0;
        for (int i = tid; i < d; i += kThreads) {
            sm.z[i] = -a.y0[(size_t)b * d + i];
            for (int w = 0; w < P.wh; ++w) { Hcol[(size_t)i * P.wh + w] = w == 0 ? i * n : -1; Hval[(size_t)i * P.wh + w] = w == 0 ? 1.0 : 0.0; }
        }
        if (tid < n) { sm.pv[tid] = 1.0; sm.pinv[tid] = 1.0; }
        __syncthreads();
        UpdateOut o1;
        o1.mean_out = nullptr; o1.chol_out = chol; o1.diff_out = nullptr; o1.ref_out = nullptr; o1.scale_by_p = false;
        PhaseClock pc;
        pc.start(nullptr);
        update_stage(P, b, sm, d, E_NUGGET_ONLY, a.nugget, chol, nullptr, nullptr, Hcol, Hval, W, o1, &nonfinite, pc);
        // linearise the PDE at t0 without preconditioning (white.py:42-48, latent.py:86-95)
        evaluate_ode(P, b, sm, 1.0, 1.0, Hcol, Hval);
        UpdateOut o2;
        o2.mean_out = mean; o2.chol_out = chol; o2.diff_out = nullptr; o2.ref_out = nullptr; o2.scale_by_p = false;
        update_stage(P, b, sm, P.m, P.latent ? E_NUGGET_ONLY : E_STEP_PLUS_NUGGET, a.nugget, chol, nullptr, nullptr, Hcol,
                     Hval, W, o2, &nonfinite, pc);
        if (tid == 0 && a.status) a.status[b] = nonfinite;
        if (nonfinite) {  // as in k_run: do not leave NaNs in the padding rows of the workspace for the next member
            for (size_t k = tid; k < (size_t)P.ld * (P.m + P.D); k += kThreads) W[k] = 0.0;
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_MIN_CTAS) k_init(const Problem P, const InitArgs a);
#endif

// cov_sqrtm *= sqrt(mean local diffusion)   pdefilter.py:113-116
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void k_rescale(double* chol, const double* diff_sum, double* diff_cal, int nsteps, size_t csz, int batch) {
    for (int b = blockIdx.y; b < batch; b += gridDim.y) {  // (gridDim.y is capped below the 65535 limit)
        const double cal = diff_sum[b] / nsteps;
        const double s = sqrt(cal);
        if (blockIdx.x == 0 && threadIdx.x == 0 && diff_cal) diff_cal[b] = cal;
        double* c = chol + (size_t)b * csz;
        for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < csz; k += (size_t)gridDim.x * blockDim.x) c[k] *= s;
    }
}
#else
__global__ void k_rescale(double* chol, const double* diff_sum, double* diff_cal, int nsteps, size_t csz, int batch);
#endif

// Stand-alone marginal read-out of `count` factors (D x D each): out[count][dd].
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void k_marginal_std(const double* chol, double* out, int D, int n, int dd, int count) {
    for (int b = blockIdx.x; b < count; b += gridDim.x)
        marginal_std_rows(chol + (size_t)b * D * D, D, n, dd, out + (size_t)b * dd, threadIdx.x >> 5, blockDim.x >> 5);
}
#else
__global__ void k_marginal_std(const double* chol, double* out, int D, int n, int dd, int count);
#endif

// K = Lk Lk^T (spatial Gram matrix), once per set_prior.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void k_gram(const double* Lk, double* Kg, int d) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        double acc = 0.0;
        const int kmax = r < c ? r : c;
        for (int k = 0; k <= kmax; ++k) acc = fma(Lk[(size_t)r * d + k], Lk[(size_t)c * d + k], acc);
        Kg[(size_t)r * d + c] = acc;
    }
}
#else
__global__ void k_gram(const double* Lk, double* Kg, int d);
#endif

// ---------------------------------------------------------------- dense sqrt entry points
// propagate_cholesky_factor for dense S1 (r x c1), S2 (r x c2): QR of the (c1+c2) x r stack.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void __launch_bounds__(kThreads) k_sqrt_propagate(const double* S1, const double* S2, double* out, int r,
                                                            int c1, int c2, int batch, double* Wall) {
    extern __shared__ __align__(16) double smem_raw[];
    double* vbuf = smem_raw;
    double* red = smem_raw + (c1 + c2) + 4;  // 16 doubles
    const int tid = threadIdx.x;
    const int rows = c1 + c2, ld = rows;
    const int k = rows < r ? rows : r;
    double* W = Wall + (size_t)blockIdx.x * ld * r;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int idx = tid; idx < rows * r; idx += kThreads) {
            const int col = idx / rows, row = idx - col * rows;
            W[(size_t)col * ld + row] = row < c1 ? S1[((size_t)b * r + col) * c1 + row] : S2[((size_t)b * r + col) * c2 + (row - c1)];
        }
        __syncthreads();
        Shape sh; sh.nt = rows; sh.nbot = 0; sh.ncols = r; sh.te = nullptr; sh.be = nullptr;
        householder_qr(W, ld, sh, vbuf, red);
        for (int idx = tid; idx < r * k; idx += kThreads) {  // out[i][j] = R[j][i], i < r, j < k
            const int i = idx / k, j = idx - i * k;
            out[(size_t)b * r * k + idx] = j <= i ? W[(size_t)i * ld + j] : 0.0;
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads) k_sqrt_propagate(const double* S1, const double* S2, double* out, int r,
                                                            int c1, int c2, int batch, double* Wall);
#endif

#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void __launch_bounds__(kThreads) k_sqrt_update(const double* H, const double* C, const double* E, double* C_out,
                                                         double* K_out, double* S_out, int m, int D, int batch,
                                                         double* Wall) {
    extern __shared__ __align__(16) double smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbot = E ? m : 0;
    const int rows = D + nbot, ld = D + m, ncols = m + D;
    double* vbuf = smem_raw;
    double* red = smem_raw + ld + 4;
    double* W = Wall + (size_t)blockIdx.x * ld * ncols;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Hb = H + (size_t)b * m * D;
        const double* Cb = C + (size_t)b * D * D;
        // right block: C^T (column k = row k of C), bottom zero
        for (int k = warp; k < D; k += kWarps) {
            double* col = W + (size_t)(m + k) * ld;
            for (int i = lane; i < D; i += 32) col[i] = Cb[(size_t)k * D + i];
            for (int i = D + lane; i < rows; i += 32) col[i] = 0.0;
        }
        // left block: C^T H^T on top, E^T below
        for (int r = warp; r < m; r += kWarps) {
            double* col = W + (size_t)r * ld;
            for (int i = lane; i < D; i += 32) {
                double acc = 0.0;
                for (int k = 0; k < D; ++k) acc = fma(Cb[(size_t)k * D + i], Hb[(size_t)r * D + k], acc);
                col[i] = acc;
            }
            if (E) for (int i = lane; i < m; i += 32) col[D + i] = E[((size_t)b * m + r) * m + i];
        }
        __syncthreads();
        Shape sh; sh.nt = D; sh.nbot = nbot; sh.ncols = ncols; sh.te = nullptr; sh.be = nullptr;
        householder_qr(W, ld, sh, vbuf, red);
        // S_out = R1^T, C_out = R3^T
        for (int idx = tid; idx < m * m; idx += kThreads) {
            const int i = idx / m, j = idx - i * m;
            S_out[(size_t)b * m * m + idx] = j <= i ? W[(size_t)i * ld + j] : 0.0;
        }
        for (int idx = tid; idx < D * D; idx += kThreads) {
            const int i = idx / D, j = idx - i * D;
            double v = 0.0;
            if (j <= i && m + j < rows) v = W[(size_t)(m + i) * ld + m + j];
            C_out[(size_t)b * D * D + idx] = v;
        }
        __syncthreads();
        // gain K = (R1^-1 R2)^T, one warp per column of R2 (back substitution, axpy form)
        for (int k = warp; k < D; k += kWarps) {
            double* col = W + (size_t)(m + k) * ld;
            for (int i = m - 1; i >= 0; --i) {
                const double* ri = W + (size_t)i * ld;
                const double xi = col[i] / ri[i];
                __syncwarp();
                for (int c = lane; c < i; c += 32) col[c] = fma(-xi, ri[c], col[c]);
                if (lane == 0) K_out[((size_t)b * D + k) * m + i] = xi;
                __syncwarp();
            }
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads) k_sqrt_update(const double* H, const double* C, const double* E, double* C_out,
                                                         double* K_out, double* S_out, int m, int D, int batch,
                                                         double* Wall);
#endif

// Square-root RTS smoother step (src/pnmol/base/kalman.py:49-66): new_mean = m - G (mp - m_fut); the new factor is the
// transpose of R[d:2d, d:] of the QR of the 3d x 2d matrix [[x^T, sc^T], [sq^T, 0], [0, sc_fut^T G^T]].
// All inputs row-major [batch, ...]; one CTA per member, unblocked Householder QR on an L2-resident workspace.
// (sc_fut may be any square root of the future covariance, not necessarily triangular.)
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void __launch_bounds__(kThreads) k_smoother_step(const double* m, const double* sc, const double* m_fut,
                                                           const double* sc_fut, const double* sgain, const double* sq,
                                                           const double* mp, const double* x, double* mean_out,
                                                           double* chol_out, int d, int batch, double* Wall) {
    extern __shared__ __align__(16) double smem_raw[];
    double* vbuf = smem_raw;
    double* red = smem_raw + 3 * d + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = 3 * d, ld = rows, ncols = 2 * d;
    double* W = Wall + (size_t)blockIdx.x * ld * ncols;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        const size_t o1 = (size_t)b * d, o2 = (size_t)b * d * d;
        for (int i = warp; i < d; i += kWarps) {  // new_mean[i] = m[i] - sum_k G[i][k] (mp[k] - m_fut[k])
            double acc = 0.0;
            for (int k = lane; k < d; k += 32) acc = fma(sgain[o2 + (size_t)i * d + k], mp[o1 + k] - m_fut[o1 + k], acc);
            acc = warp_sum(acc);
            if (lane == 0) mean_out[o1 + i] = m[o1 + i] - acc;
        }
        for (int idx = tid; idx < rows * ncols; idx += kThreads) {
            const int col = idx / rows, row = idx - col * rows;
            double v = 0.0;
            if (col < d) {
                if (row < d) v = x[o2 + (size_t)col * d + row];               // x^T
                else if (row < 2 * d) v = sq[o2 + (size_t)col * d + row - d];  // sq^T
            } else {
                const int c = col - d;
                if (row < d) v = sc[o2 + (size_t)c * d + row];                 // sc^T
                else if (row >= 2 * d) {  // (sc_fut^T G^T)[r][c] = sum_k sc_fut[k][r] G[c][k]
                    const int r = row - 2 * d;
                    double acc = 0.0;
                    for (int k = 0; k < d; ++k) acc = fma(sc_fut[o2 + (size_t)k * d + r], sgain[o2 + (size_t)c * d + k], acc);
                    v = acc;
                }
            }
            W[(size_t)col * ld + row] = v;
        }
        __syncthreads();
        Shape sh; sh.nt = rows; sh.nbot = 0; sh.ncols = ncols; sh.te = nullptr; sh.be = nullptr;
        householder_qr(W, ld, sh, vbuf, red);
        for (int idx = tid; idx < d * d; idx += kThreads) {  // out[i][j] = R[d + j][d + i], j <= i
            const int i = idx / d, j = idx - i * d;
            chol_out[o2 + idx] = j <= i ? W[(size_t)(d + i) * ld + d + j] : 0.0;
        }
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(kThreads) k_smoother_step(const double* m, const double* sc, const double* m_fut,
                                                           const double* sc_fut, const double* sgain, const double* sq,
                                                           const double* mp, const double* x, double* mean_out,
                                                           double* chol_out, int d, int batch, double* Wall);
#endif

}  // namespace pnmol
