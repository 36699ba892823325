// Warp-per-member EK1 kernels: every warp of the grid runs the whole filter loop of its own ensemble members
// (same phases as ek1_step / k_init in ek1_kernels.cuh, instantiated for WarpTeam, QR from qr_warp.cuh).
// There is no block-level synchronisation after the prologue; eight independent members per SM hide each
// other's latencies.  Used when every panel's compact row list has at most 256 rows (C1/C5-sized problems).
#pragma once
#include "ek1_kernels.cuh"
#include "qr_warp.cuh"

namespace pnmol {

struct WarpGeom {  // per-warp shared-memory slice, in doubles
    int ldv, per_warp, nwarps;
};

__host__ __device__ __forceinline__ int warp_ldv(int maxlen) {
    int ldv = (maxlen + 7) & ~7;
    while ((ldv & 15) != 6) ldv += 2;
    return ldv;
}

__host__ __device__ __forceinline__ int warp_smem_doubles(int D, int m, int dd, int ldm, int ldv) {
    const int big = kWNB * ldv > m * ldm ? kWNB * ldv : m * ldm;  // the reflector buffer shares space with the m x ldm scratch
    return D + 3 * m + dd + 16 + 2 * kMaxN + kWNB + kWNB * kWNB + kWNB * (kWNB + 1) + big + 8;
}

__device__ __forceinline__ Smem carve_warp(double* base, int D, int m, int dd, int ldm, int ldv, WarpQR& q) {
    Smem s;
    s.mp = base;    base += D;
    s.z = base;     base += m;
    s.y = base;     base += m;
    s.xw = base;    base += m;
    s.xat = base;   base += dd;
    s.red = base;   base += 16;
    s.pv = base;    base += kMaxN;
    s.pinv = base;  base += kMaxN;
    q.tau = base;   base += kWNB;
    q.Ts = base;    base += kWNB * kWNB;
    q.Gs = base;    base += kWNB * (kWNB + 1);
    base += ((size_t)base & 8) ? 1 : 0;  // 16-byte alignment of the big buffer
    s.msq = base;
    q.Vs = base;
    q.ldv = ldv;
    s.vbuf = nullptr; s.Vs = nullptr; s.xraw = nullptr; s.sc = nullptr; s.Vr = nullptr; s.Ts = nullptr; s.Gs = nullptr;
    return s;
}

// update_stage (ek1_device.cuh) for one warp.  Returns the local diffusion; *bad |= 1 on a non-finite result.
__device__ double update_stage_warp(const Problem& P, int b, const Smem& sm, const WarpQR& q, int mcur, EMode emode,
                                    double nugget, const double* __restrict__ Rsrc, const int32_t* te, const int32_t* be,
                                    const int32_t* Hcol, const double* Hval, double* W, const UpdateOut out, int* bad,
                                    PhaseClock& pc) {
    const int D = P.D, ld = P.ld;
    const int nbot = emode == E_NONE ? 0 : mcur;
    const int nrows = D + nbot;
    double* Wl = W + (size_t)(P.m - mcur) * ld;
    double* Wr = W + (size_t)P.m * ld;
    update_build_right(P, mcur, nrows, Rsrc, te, be, Wr, 0, 1);
    __syncwarp();
    update_build_left(P, b, mcur, nrows, emode, nugget, te, be, Hcol, Hval, Wl, Wr, 0, 1);
    __syncwarp();
    pc.mark(4);
    Shape sh;
    sh.nt = D; sh.nbot = nbot; sh.ncols = mcur + D; sh.te = te; sh.be = be;
    householder_qr_warp(Wl, ld, sh, q, pc);
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

__device__ double ek1_step_warp(const Problem& P, int b, int slot, const Smem& sm, const WarpQR& q, double dt,
                                const double* mean_in, const double* chol_in, double* mean_out, double* chol_out,
                                double* err_out, double* ref_out, int flags, int* bad, PhaseClock& pc) {
    const int lane = threadIdx.x & 31;
    const int n = P.n, D = P.D;
    double* W = P.W + (size_t)slot * P.ld * (P.m + P.D);
    int32_t* Hcol = P.Hcol + (size_t)slot * P.m * P.wh;
    double* Hval = P.Hval + (size_t)slot * P.m * P.wh;
    for (int k = lane; k < D; k += 32) {  // m = P^-1 mean, mp = A m   white.py:104-107
        const int j = k / n, i = k - j * n;
        double acc = 0.0;
        for (int s = 0; s < n; ++s) acc = fma(P.A1d[i * n + s], sm.pinv[s] * mean_in[(size_t)s * P.dd + j], acc);
        sm.mp[k] = acc;
    }
    __syncwarp();
    evaluate_ode<WarpTeam>(P, b, sm, sm.pv[0], sm.pv[1], Hcol, Hval);
    pc.mark(0);
    const bool dense = flags & 1;
    build_predict<WarpTeam>(P, b, sm, chol_in, dense ? P.te_pd : P.te_p, W + (size_t)P.m * P.ld, 0, 1);
    pc.mark(1);
    Shape sp;
    sp.nt = D; sp.nbot = D; sp.ncols = D; sp.te = dense ? P.te_pd : P.te_p; sp.be = P.be_p;
    householder_qr_warp(W + (size_t)P.m * P.ld, P.ld, sp, q, pc);
    pc.mark(2);
    if (!P.latent && !(flags & 2)) {
        error_estimate<WarpTeam>(P, b, sm, sm.pv[1], dt, E_STEP_WHITE, 0.0, Hcol, Hval, P.F + (size_t)slot * P.m * P.d,
                                 P.S + (size_t)slot * P.m * P.m, err_out);
    }
    pc.mark(3);
    UpdateOut out;
    out.mean_out = mean_out; out.chol_out = chol_out; out.diff_out = nullptr;
    out.ref_out = P.latent ? nullptr : ref_out; out.scale_by_p = true;
    return update_stage_warp(P, b, sm, q, P.m, P.latent ? E_NONE : E_STEP_WHITE, 0.0, nullptr, P.te_u, P.be_u, Hcol, Hval, W,
                             out, bad, pc);
}

__global__ void __launch_bounds__(256, 1) k_run_warp(const Problem P, const RunArgs a, const WarpGeom geo) {
    extern __shared__ double smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpQR q;
    const Smem sm = carve_warp(smem_raw + (size_t)warp * geo.per_warp, P.D, P.m, P.dd, P.ldm, geo.ldv, q);
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
            const int flags = s == 0 ? a.flags : (a.flags & ~1);
            diff = ek1_step_warp(P, b, wid, sm, q, dt, min_, cin_, mout, cout, a.err_out ? a.err_out + (size_t)b * P.d : nullptr,
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
        __syncwarp();
    }
}

// initialize(): two square-root updates on a Kronecker-structured prior factor (k_init for one warp per member).
__global__ void __launch_bounds__(256, 1) k_init_warp(const Problem P, const InitArgs a, const WarpGeom geo) {
    extern __shared__ double smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpQR q;
    const Smem sm = carve_warp(smem_raw + (size_t)warp * geo.per_warp, P.D, P.m, P.dd, P.ldm, geo.ldv, q);
    const int wid = blockIdx.x * geo.nwarps + warp, nw = gridDim.x * geo.nwarps;
    const int n = P.n, d = P.d, D = P.D, nd = P.n * P.d;
    double* W = P.W + (size_t)wid * P.ld * (P.m + P.D);
    int32_t* Hcol = P.Hcol + (size_t)wid * P.m * P.wh;
    double* Hval = P.Hval + (size_t)wid * P.m * P.wh;
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
            for (int w = 0; w < P.wh; ++w) { Hcol[(size_t)i * P.wh + w] = w == 0 ? i * n : -1; Hval[(size_t)i * P.wh + w] = w == 0 ? 1.0 : 0.0; }
        }
        if (lane < n) { sm.pv[lane] = 1.0; sm.pinv[lane] = 1.0; }
        __syncwarp();
        UpdateOut o1;
        o1.mean_out = nullptr; o1.chol_out = chol; o1.diff_out = nullptr; o1.ref_out = nullptr; o1.scale_by_p = false;
        update_stage_warp(P, b, sm, q, d, E_NUGGET_ONLY, a.nugget, chol, nullptr, nullptr, Hcol, Hval, W, o1, &bad, pc);
        evaluate_ode<WarpTeam>(P, b, sm, 1.0, 1.0, Hcol, Hval);  // white.py:42-48, latent.py:86-95
        UpdateOut o2;
        o2.mean_out = mean; o2.chol_out = chol; o2.diff_out = nullptr; o2.ref_out = nullptr; o2.scale_by_p = false;
        update_stage_warp(P, b, sm, q, P.m, P.latent ? E_NUGGET_ONLY : E_STEP_PLUS_NUGGET, a.nugget, chol, nullptr, nullptr, Hcol,
                          Hval, W, o2, &bad, pc);
        if (lane == 0 && a.status) a.status[b] = bad;
        __syncwarp();
    }
}

}  // namespace pnmol
