// Small-state EK1 kernels (north-star kernel 1): one WARP per ensemble member, everything of a member -- the QR
// workspace included -- in the warp's own slice of shared memory, 12-24 members resident per SM.
//
// For D <~ 48 (the meshes the reference itself tests and plots: dx = 0.2, D = 18) a member-step is a few ten kflop on
// 5-20 KB of state: there is nothing to block and nothing to feed a tensor core with.  What matters is (a) that the
// per-column dependent chain of the Householder QR (reduce - scalars - update) is paid by ONE warp while the other
// warps of the SM run their own members, (b) that the code is small and rolled (a few KB of SASS: many de-synchronised
// warps share the 32 KB instruction cache), and (c) that the only global traffic is the state itself, streamed with
// coalesced accesses: B_alg = 8 [2 (D^2 + D) + d + 2] bytes per member-step.
//
// Same phases (evaluate_ode, build_predict, error_estimate, update_build_*, update_solve, update_output_*) as the
// CTA-per-member kernels, instantiated for WarpTeam; the QR is the unblocked, envelope-aware Householder of
// householder_columns with the trailing update turned around: a LANE owns a trailing column (dot product and update
// are per-lane loops over the reflector's support -- no shuffle reductions), only the column norm is a warp reduction.
// LAPACK dlarfg conventions as everywhere (src/pnmol/base/sqrt.py:21,66,88).
#pragma once
#include "ek1_kernels.cuh"

namespace pnmol {

struct SmallGeom {
    int per_warp;  // doubles of shared memory per warp
    int nwarps;    // warps (= concurrently resident members) per CTA
};

__host__ __device__ __forceinline__ size_t small_smem_doubles(int D, int m, int dd, int ld, int ldm, int wh) {
    return (size_t)ld * (m + D) + D + 3 * (size_t)m + dd + 16 + 2 * kMaxN + (size_t)m * ldm + 2 * (size_t)m * wh + 8;
}

__device__ __forceinline__ Smem carve_small(double* base, const Problem& P, double*& W) {
    Smem s;
    W = base;       base += (size_t)P.ld * (P.m + P.D);
    s.mp = base;    base += P.D;
    s.z = base;     base += P.m;
    s.y = base;     base += P.m;
    s.xw = base;    base += P.m;
    s.xat = base;   base += P.dd;
    s.red = base;   base += 16;
    s.pv = base;    base += kMaxN;
    s.pinv = base;  base += kMaxN;
    s.msq = base;   base += (size_t)P.m * P.ldm;
    s.Hval = base;  base += (size_t)P.m * P.wh;
    s.Hcol = reinterpret_cast<int32_t*>(base);
    s.Hpt = s.Hcol + (size_t)P.m * P.wh;
    s.vbuf = nullptr; s.Vs = nullptr; s.xraw = nullptr; s.sc = nullptr; s.Vr = nullptr; s.Ts = nullptr; s.Gs = nullptr;
    s.fqbase = nullptr; s.fqend = nullptr;
    s.te_p = nullptr; s.be_p = nullptr; s.te_pd = nullptr; s.te_u = nullptr; s.be_u = nullptr;
    return s;
}

// Unblocked Householder QR of the column-major matrix W (per-warp shared memory, odd leading dimension: the lanes'
// columns fall into different banks) by one warp.  Envelopes as in householder_columns; on return the upper triangle
// holds R (the reflectors stay below the diagonal and are never read again).
static __device__ __noinline__ void qr_small(double* __restrict__ W, int ld, const Shape s) {
    const int lane = threadIdx.x & 31;
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
#pragma unroll 1
    for (int j = 0; j < nref; ++j) {
        int e1, a2, e2;  // support of reflector j: rows [j, e1] and [a2, e2]
        if (j < s.nt) {
            e1 = s.te ? s.te[j] : s.nt - 1;
            if (e1 > s.nt - 1) e1 = s.nt - 1;
            if (e1 < j) e1 = j;
            a2 = s.nt;
            e2 = s.be ? s.be[j] : nrows - 1;
            if (e2 > nrows - 1) e2 = nrows - 1;
        } else {
            e1 = s.be ? s.be[j] : nrows - 1;
            if (e1 > nrows - 1) e1 = nrows - 1;
            if (e1 < j) e1 = j;
            a2 = 0;
            e2 = -1;
        }
        double* col = W + (size_t)j * ld;
        double ss = 0.0;
        for (int r = j + 1 + lane; r <= e1; r += 32) ss = fma(col[r], col[r], ss);
        for (int r = a2 + lane; r <= e2; r += 32) ss = fma(col[r], col[r], ss);
        ss = warp_sum(ss);
        const double al = col[j];
        double tau = 0.0, beta = al, scale = 0.0;
        if (ss != 0.0) {  // dlarfg: zero sub-column -> H = I
            const double s2 = fma(al, al, ss);
            const double rn = rsqrt(s2);
            const double nrm = s2 * rn;
            beta = -copysign(nrm, al);
            tau = (beta - al) * -copysign(rn, al);
            scale = __drcp_rn(al - beta);
        }
        __syncwarp();
        if (tau != 0.0) {
            for (int r = j + 1 + lane; r <= e1; r += 32) col[r] *= scale;
            for (int r = a2 + lane; r <= e2; r += 32) col[r] *= scale;
            if (lane == 0) col[j] = beta;
            __syncwarp();
            // trailing columns: one per lane; v = (1, col[j+1 .. e1], col[a2 .. e2])
            for (int k = j + 1 + lane; k < s.ncols; k += 32) {
                double* ck = W + (size_t)k * ld;
                double d0 = ck[j], d1 = 0.0;
                int r = j + 1;
#pragma unroll 2
                for (; r + 1 <= e1; r += 2) { d0 = fma(col[r], ck[r], d0); d1 = fma(col[r + 1], ck[r + 1], d1); }
                if (r <= e1) d0 = fma(col[r], ck[r], d0);
                r = a2;
#pragma unroll 4
                for (; r + 1 <= e2; r += 2) { d0 = fma(col[r], ck[r], d0); d1 = fma(col[r + 1], ck[r + 1], d1); }
                if (r <= e2) d0 = fma(col[r], ck[r], d0);
                const double w = -tau * (d0 + d1);
                ck[j] += w;
#pragma unroll 4
                for (r = j + 1; r <= e1; ++r) ck[r] = fma(w, col[r], ck[r]);
#pragma unroll 8
                for (r = a2; r <= e2; ++r) ck[r] = fma(w, col[r], ck[r]);
            }
        }
        __syncwarp();
    }
}

// update_stage (ek1_device.cuh) for one warp on its shared-memory workspace.  Returns the local diffusion.
static __device__ __noinline__ double update_stage_small(const Problem& P, int b, const Smem& sm, int mcur, EMode emode, double nugget,
                                                  const double* __restrict__ Rsrc, const int32_t* te, const int32_t* be,
                                                  double* W, const UpdateOut out, int* bad, PhaseClock& pc) {
    const int D = P.D, ld = P.ld;
    const int nbot = emode == E_NONE ? 0 : mcur;
    const int nrows = D + nbot;
    double* Wl = W + (size_t)(P.m - mcur) * ld;
    double* Wr = W + (size_t)P.m * ld;
    update_build_right(P, mcur, nrows, Rsrc, te, be, Wr, 0, 1);
    __syncwarp();
    update_build_left(P, b, mcur, nrows, emode, nugget, te, be, sm.Hcol, sm.Hval, Wl, Wr, 0, 1);
    __syncwarp();
    pc.mark(4);
    Shape sh;
    sh.nt = D; sh.nbot = nbot; sh.ncols = mcur + D; sh.te = te; sh.be = be;
    qr_small(Wl, ld, sh);
    pc.mark(5);
    const double diff = update_solve<WarpTeam>(P, sm, mcur, Wl, Wr);
    pc.mark(6);
    int flag = update_output_mean<WarpTeam>(P, sm, out, diff);
    flag |= update_output_factor(P, sm, out, mcur, nrows, Wr, 0, 1);
    if (__any_sync(0xffffffffu, flag)) *bad |= 1;
    __syncwarp();
    pc.mark(7);
    return diff;
}

__device__ __forceinline__ double ek1_step_small(const Problem& P, int b, const Smem& sm, double* W, double dt,
                                                 const double* mean_in, const double* chol_in, double* mean_out, double* chol_out,
                                                 double* err_out, double* ref_out, int flags, int* bad, PhaseClock& pc) {
    const int lane = threadIdx.x & 31;
    const int n = P.n, D = P.D;
    for (int k = lane; k < D; k += 32) {  // m = P^-1 mean, mp = A m   white.py:104-107
        const int j = k / n, i = k - j * n;
        double acc = 0.0;
        for (int s = 0; s < n; ++s) acc = fma(P.A1d[i * n + s], sm.pinv[s] * mean_in[(size_t)s * P.dd + j], acc);
        sm.mp[k] = acc;
    }
    __syncwarp();
    evaluate_ode<WarpTeam>(P, b, sm, sm.pv[0], sm.pv[1], sm.Hcol, sm.Hval);
    pc.mark(0);
    const bool dense = flags & 1;
    build_predict<WarpTeam>(P, b, sm, chol_in, dense ? P.te_pd : P.te_p, W + (size_t)P.m * P.ld, 0, 1);
    pc.mark(1);
    Shape sp;
    sp.nt = D; sp.nbot = D; sp.ncols = D; sp.te = dense ? P.te_pd : P.te_p; sp.be = P.be_p;
    qr_small(W + (size_t)P.m * P.ld, P.ld, sp);
    pc.mark(2);
    if (!P.latent && !(flags & 2)) error_estimate_smem<WarpTeam>(P, b, sm, sm.pv[1], dt, E_STEP_WHITE, sm.Hcol, sm.Hval, err_out);
    pc.mark(3);
    UpdateOut out;
    out.mean_out = mean_out; out.chol_out = chol_out; out.diff_out = nullptr;
    out.ref_out = P.latent ? nullptr : ref_out; out.scale_by_p = true;
    return update_stage_small(P, b, sm, P.m, P.latent ? E_NONE : E_STEP_WHITE, 0.0, nullptr, P.te_u, P.be_u, W, out, bad, pc);
}

#ifndef PNMOL_SMALL_THREADS
#define PNMOL_SMALL_THREADS 512
#endif

#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_SMALL)
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_run_small(const Problem P, const RunArgs a, const SmallGeom geo) {
    extern __shared__ __align__(16) double smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* W;
    const Smem sm = carve_small(smem_raw + (size_t)warp * geo.per_warp, P, W);
    for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
    __syncwarp();
    const int wid = blockIdx.x * geo.nwarps + warp, nw = gridDim.x * geo.nwarps;
    const size_t msz = (size_t)P.D, csz = (size_t)P.D * P.D;
    PhaseClock pc;
    pc.start(P.prof);
    for (int b = wid; b < P.batch; b += nw) {
        int bad = 0;
        double diffsum = 0.0, diff = 0.0;
        for (int s = 0; s < a.nsteps; ++s) {
            double dt;
            __syncwarp();
            if (a.nsteps == 1 && a.pv == nullptr) {
                if (lane < P.n) { sm.pv[lane] = a.pv0[lane]; sm.pinv[lane] = a.pinv0[lane]; }
                dt = a.dt0;
            } else {
                if (lane < P.n) { sm.pv[lane] = a.pv[(size_t)s * P.n + lane]; sm.pinv[lane] = a.pinv[(size_t)s * P.n + lane]; }
                dt = a.dts[s];
            }
            __syncwarp();
            const bool even = (s & 1) == 0;
            const double* min_ = (even ? a.mean_a : a.mean_b) + b * msz;
            const double* cin_ = (even ? a.chol_a : a.chol_b) + b * csz;
            double* mout = (even ? a.mean_b : a.mean_a) + b * msz;
            double* cout = (even ? a.chol_b : a.chol_a) + b * csz;
            const int flags = s == 0 ? a.flags : (a.flags & ~1);  // only the first step may see a dense factor
            diff = ek1_step_small(P, b, sm, W, dt, min_, cin_, mout, cout, a.err_out ? a.err_out + (size_t)b * P.d : nullptr,
                                  a.ref_out ? a.ref_out + (size_t)b * P.d : nullptr, flags, &bad, pc);
            diffsum += diff;
            if (a.mean_traj) {
                double* mt = a.mean_traj + ((size_t)s * P.batch + b) * msz;
                for (size_t k = lane; k < msz; k += 32) mt[k] = mout[k];
            }
            if (a.chol_traj) {
                double* ct = a.chol_traj + ((size_t)s * P.batch + b) * csz;
                for (size_t k = lane; k < csz; k += 32) ct[k] = cout[k];
            }
            if (a.std_traj) marginal_std_rows(cout, P.D, P.n, P.dd, a.std_traj + ((size_t)s * P.batch + b) * P.dd, 0, 1);
            __syncwarp();
        }
        if ((a.nsteps & 1) && !a.final_in_b) {  // result sits in b: bring it home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = lane; k < msz; k += 32) md[k] = ms[k];
            for (size_t k = lane; k < csz; k += 32) cd[k] = cs[k];
        }
        if (lane == 0) {
            if (a.diff_last) a.diff_last[b] = diff;
            if (a.diff_sum) a.diff_sum[b] = diffsum;
            if (a.status) a.status[b] = bad;
        }
        if (bad) {  // do not leave non-finite values in the workspace for the warp's next member
            for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
        }
        __syncwarp();
    }
}
#else
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_run_small(const Problem P, const RunArgs a, const SmallGeom geo);
#endif

// initialize(): two square-root updates on a Kronecker-structured prior factor (white.py:12-80, latent.py:20-134).
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_SMALL)
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_init_small(const Problem P, const InitArgs a, const SmallGeom geo) {
    extern __shared__ __align__(16) double smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* W;
    const Smem sm = carve_small(smem_raw + (size_t)warp * geo.per_warp, P, W);
    for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
    __syncwarp();
    const int wid = blockIdx.x * geo.nwarps + warp, nw = gridDim.x * geo.nwarps;
    const int n = P.n, d = P.d, D = P.D, nd = P.n * P.d;
    PhaseClock pc;
    pc.start(nullptr);
    for (int b = wid; b < P.batch; b += nw) {
        int bad = 0;
        double* chol = a.chol_out + (size_t)b * D * D;
        double* mean = a.mean_out + (size_t)b * D;
        const double ps = P.priorscale ? P.priorscale[b] : 1.0;
        for (int r = 0; r < D; ++r) {  // C0 = kron(Lk, c0 I_n), latent: blockdiag(., kron(E_sqrtm, c0 I_n))
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
        for (int k = lane; k < D; k += 32) sm.mp[k] = 0.0;
        for (int i = lane; i < d; i += 32) {  // update on the initial condition: H = E0, z = -y0   white.py:32-39
            sm.z[i] = -a.y0[(size_t)b * d + i];
            for (int w = 0; w < P.wh; ++w) { sm.Hcol[(size_t)i * P.wh + w] = w == 0 ? i * n : -1; sm.Hval[(size_t)i * P.wh + w] = w == 0 ? 1.0 : 0.0; }
        }
        if (lane < n) { sm.pv[lane] = 1.0; sm.pinv[lane] = 1.0; }
        __syncwarp();
        __threadfence_block();  // the factor just written is re-read (Rsrc) by other lanes of this warp
        UpdateOut o1;
        o1.mean_out = nullptr; o1.chol_out = chol; o1.diff_out = nullptr; o1.ref_out = nullptr; o1.scale_by_p = false;
        update_stage_small(P, b, sm, d, E_NUGGET_ONLY, a.nugget, chol, nullptr, nullptr, W, o1, &bad, pc);
        __threadfence_block();
        evaluate_ode<WarpTeam>(P, b, sm, 1.0, 1.0, sm.Hcol, sm.Hval);  // white.py:42-48, latent.py:86-95
        UpdateOut o2;
        o2.mean_out = mean; o2.chol_out = chol; o2.diff_out = nullptr; o2.ref_out = nullptr; o2.scale_by_p = false;
        update_stage_small(P, b, sm, P.m, P.latent ? E_NUGGET_ONLY : E_STEP_PLUS_NUGGET, a.nugget, chol, nullptr, nullptr, W, o2, &bad, pc);
        if (lane == 0 && a.status) a.status[b] = bad;
        if (bad) {
            for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
        }
        __syncwarp();
    }
}
#else
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_init_small(const Problem P, const InitArgs a, const SmallGeom geo);
#endif

// Adaptive time loop on the device, one warp per member (k_run_adaptive of ek1_kernels.cuh; src/pnmol/pdefilter.py:118-227
// with src/pnmol/odetools/step.py:58-119): accept/reject and the step-size proposal per member inside the kernel.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_SMALL)
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_run_adaptive_small(const Problem P, const AdaptiveArgs a, const SmallGeom geo) {
    extern __shared__ __align__(16) double smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* W;
    const Smem sm = carve_small(smem_raw + (size_t)warp * geo.per_warp, P, W);
    for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
    __syncwarp();
    const int wid = blockIdx.x * geo.nwarps + warp, nw = gridDim.x * geo.nwarps;
    const int nu = P.n - 1;
    const size_t msz = (size_t)P.D, csz = (size_t)P.D * P.D;
    PhaseClock pc;
    pc.start(nullptr);
    for (int b = wid; b < P.batch; b += nw) {
        int bad = 0;
        double t = a.t0, dt = a.dt0[b], diffsum = 0.0, difflast = 0.0;
        int nsteps = 0, natt = 0, cur = 0, stat = 0;
        double* err = a.err + (size_t)b * P.d;
        double* ref = a.ref + (size_t)b * P.d;
        while (t < a.tmax) {
            if (natt >= a.max_attempts) { stat |= 2; break; }
            if (!(dt >= 0.0)) { stat |= 1; break; }  // pdefilter.py:225 asserts dt >= 0 (a NaN proposal ends here too)
            __syncwarp();
            if (lane < P.n) {  // Nordsieck preconditioner p_i = |dt|^(nu - i + 1/2) / (nu - i)!   (iwp.py:55-62)
                const int k = nu - lane;
                double fact = 1.0;
                for (int q = 2; q <= k; ++q) fact *= q;
                const double pw = pow(fabs(dt), k + 0.5);
                sm.pv[lane] = pw / fact;
                sm.pinv[lane] = fact / pw;
            }
            __syncwarp();
            const double* min_ = (cur ? a.mean_b : a.mean_a) + b * msz;
            const double* cin_ = (cur ? a.chol_b : a.chol_a) + b * csz;
            double* mout = (cur ? a.mean_a : a.mean_b) + b * msz;
            double* cout = (cur ? a.chol_a : a.chol_b) + b * csz;
            const int flags = (natt == 0 || nsteps == 0) ? a.flags : (a.flags & ~1);
            int badstep = 0;
            const double diff = ek1_step_small(P, b, sm, W, dt, min_, cin_, mout, cout, err, ref, flags, &badstep, pc);
            __threadfence_block();
            __syncwarp();
            double part = 0.0;  // scaled error norm (step.py:97-108 on dt * error_estimate, pdefilter.py:208-213)
            for (int i = lane; i < P.d; i += 32) {
                const double r = dt * err[i] / (a.abstol + a.reltol * ref[i]);
                part = fma(r, r, part);
            }
            const double norm = sqrt(warp_sum(part)) / sqrt((double)P.d);
            double change = a.safety * pow(1.0 / norm, a.inv_rate);
            change = fmax(a.change_min, fmin(change, a.change_max));
            if (!(norm == norm)) change = norm;  // NaN propagates like jnp.minimum / jnp.maximum
            const double suggested = change * dt;
            ++natt;
            if (norm < 1.0) {  // accepted: the proposal becomes the state
                t = t + dt;
                cur ^= 1;
                ++nsteps;
                difflast = diff;
                diffsum += diff;
                bad |= badstep;
                if (a.mean_traj) {  // the accepted state joins the trajectory
                    if (nsteps <= a.max_traj) {
                        double* mt = a.mean_traj + ((size_t)(nsteps - 1) * P.batch + b) * msz;
                        double* ct = a.chol_traj + ((size_t)(nsteps - 1) * P.batch + b) * csz;
                        for (size_t k = lane; k < msz; k += 32) mt[k] = mout[k];
                        for (size_t k = lane; k < csz; k += 32) ct[k] = cout[k];
                        if (lane == 0) a.t_traj[(size_t)b * a.max_traj + nsteps - 1] = t;
                    } else {
                        stat |= 4;
                    }
                }
            }
            dt = fmin(suggested, a.tmax - t);
            if (!(suggested == suggested)) dt = suggested;
        }
        if (cur) {  // bring the final state home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = lane; k < msz; k += 32) md[k] = ms[k];
            for (size_t k = lane; k < csz; k += 32) cd[k] = cs[k];
        }
        if (lane == 0) {
            a.t_out[b] = t; a.dt_out[b] = dt; a.diff_sum[b] = diffsum; a.diff_last[b] = difflast;
            a.nsteps[b] = nsteps; a.nattempts[b] = natt; a.status[b] = stat | (bad ? 1 : 0);
        }
        if (bad || stat) {
            for (int k = lane; k < geo.per_warp; k += 32) smem_raw[(size_t)warp * geo.per_warp + k] = 0.0;
        }
        __syncwarp();
    }
}
#else
__global__ void __launch_bounds__(PNMOL_SMALL_THREADS, 1) k_run_adaptive_small(const Problem P, const AdaptiveArgs a, const SmallGeom geo);
#endif

}  // namespace pnmol
