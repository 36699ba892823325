// Multi-CTA EK1 kernels for large state dimension: the whole grid (cooperative launch, one CTA per SM) works on one
// member at a time.  Same phases as ek1_step / k_init (ek1_kernels.cuh); the O(D^2) builds and read-outs are spread
// over all warps of the grid, the two QRs run on householder_qr_large, the O(m^2) solves stay on CTA 0.  The vectors
// the single-CTA path keeps in shared memory (mp, z, y, xw, xat) live in global scratch here (LargeQR::vec).
#pragma once
#include "ek1_kernels.cuh"
#include "qr_large.cuh"

#ifndef PNMOL_LARGE_CTAS
#define PNMOL_LARGE_CTAS 2   // CTAs per SM the multi-CTA kernels are compiled for (128 registers per thread)
#endif
namespace pnmol {

__device__ __forceinline__ Smem large_vectors(const Problem& P, const LargeQR& q, const LargeSmem& ls) {
    Smem sm;
    double* base = q.vec;
    sm.mp = base;   base += P.D;
    sm.z = base;    base += P.m;
    sm.y = base;    base += P.m;
    sm.xw = base;   base += P.m;
    sm.xat = base;
    sm.vbuf = nullptr; sm.Vs = nullptr; sm.xraw = nullptr; sm.Vr = nullptr; sm.msq = nullptr;
    sm.red = ls.red; sm.pv = ls.pv; sm.pinv = ls.pinv; sm.sc = ls.sc; sm.Ts = ls.Ts; sm.Gs = ls.Gs;
    sm.tri = (size_t)2 * P.m + kTriScratch <= (size_t)q.cap ? ls.PB : nullptr;  // (the panel buffer is idle during the solves)
    return sm;
}

// Error estimate (white.py:153-162) with the m x d / m x m assemblies spread over the grid and a right-looking
// Cholesky whose trailing update is spread over the grid (the diagonal of L is kept apart in q.Ld so that no CTA
// reads an entry another one overwrites in the same phase).  sigma and the read-out stay on CTA 0.
static __device__ void error_estimate_large(cg::grid_group& grid, const Problem& P, int b, const Smem& sm, const LargeQR& q,
                                     const LargeSmem& ls, double p1s, double dt, EMode emode, const int32_t* Hcol, const double* Hval,
                                     double* F, double* S, double* err_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * kWarps + warp, gnw = gridDim.x * kWarps;
    const int n = P.n, d = P.d, m = P.m;
    const double ps = P.priorscale ? P.priorscale[b] : 1.0;
    const double ps2 = ps * ps;
    double q00 = 0, q01 = 0, q11 = 0;
    for (int s = 0; s < n; ++s) {
        q00 = fma(P.LQ1d[s], P.LQ1d[s], q00);
        q01 = fma(P.LQ1d[s], P.LQ1d[n + s], q01);
        q11 = fma(P.LQ1d[n + s], P.LQ1d[n + s], q11);
    }
    for (int r = gw; r < m; r += gnw) {  // F = At K  (m x d)
        for (int k = lane; k < d; k += 32) {
            double acc = 0.0;
            for (int w = 0; w < P.wh; ++w) {
                const int c = Hcol[(size_t)r * P.wh + w];
                if (c >= 0 && (c % n) == 0 && c < n * d) acc = fma(Hval[(size_t)r * P.wh + w], ps2 * P.Kg[(size_t)(c / n) * d + k], acc);
            }
            F[(size_t)r * d + k] = acc;
        }
    }
    grid.sync();
    for (int r = gw; r < m; r += gnw) {  // S (m x m), row-major, full
        for (int rp = lane; rp < m; rp += 32) {
            double acc = 0.0;
            for (int w = 0; w < P.wh; ++w) {
                const int c = Hcol[(size_t)rp * P.wh + w];
                if (c >= 0 && (c % n) == 0 && c < n * d) acc = fma(F[(size_t)r * d + c / n], Hval[(size_t)rp * P.wh + w], acc);
            }
            double val = q00 * acc;
            double cross = 0.0;
            if (rp < d) cross += F[(size_t)r * d + rp];
            if (r < d) cross += F[(size_t)rp * d + r];
            val += q01 * p1s * cross;
            if (r < d && rp < d) val += q11 * p1s * p1s * ps2 * P.Kg[(size_t)r * d + rp];
            S[(size_t)r * m + rp] = val + meas_cov_entry(P, b, emode, r, rp);
        }
    }
    grid.sync();
    if (blockIdx.x == 0)
        for (int r = tid; r < m; r += kThreads) sm.y[r] = S[(size_t)r * m + r];  // diag(S) before factorisation
    // Right-looking blocked Cholesky of the lower triangle (row-major), panels of kCB = 32 columns, no serial panel:
    //   (1) every CTA factors the kb x kb diagonal block redundantly in its own shared memory (no communication),
    //   (2) the rows below the block are independent triangular solves  L21[r] = A21[r] L11^-T: one thread per row,
    //       rows dealt round-robin to the CTAs; CTA 0 writes L11 and the diagonal (q.Ld) back,
    //   (3) after one grid barrier the trailing matrix is updated in 64 x 64 tiles, one per CTA and round: the tile's
    //       two row blocks of L21 are staged in shared memory, each thread owns a 4 x 4 micro-tile.
    constexpr int kCB = 32, kCP = kCB + 1, kTile = 64;
    double* Bs = ls.PB;                      // [kCB][kCP] diagonal block
    double* Li = Bs + kCB * kCP;             // [kTile][kCP]
    double* Lj = Li + kTile * kCP;           // [kTile][kCP]
    for (int k0 = 0; k0 < m; k0 += kCB) {
        const int kb = m - k0 < kCB ? m - k0 : kCB;
        for (int idx = tid; idx < kb * kb; idx += kThreads) {
            const int r = idx / kb, c = idx - r * kb;
            if (c <= r) Bs[r * kCP + c] = S[(size_t)(k0 + r) * m + k0 + c];
        }
        __syncthreads();
        for (int kk = 0; kk < kb; ++kk) {
            const double piv = sqrt(Bs[kk * kCP + kk]);
            const double rinv = 1.0 / piv;
            __syncthreads();
            if (tid >= kk && tid < kb) Bs[tid * kCP + kk] = tid == kk ? piv : Bs[tid * kCP + kk] * rinv;
            __syncthreads();
            const int rem = kb - kk - 1;  // B[r][c] -= L[r][kk] L[c][kk],  kk < c <= r
            for (int idx = tid; idx < rem * rem; idx += kThreads) {
                const int r = kk + 1 + idx / rem, c = kk + 1 + idx % rem;
                if (c <= r) Bs[r * kCP + c] = fma(-Bs[r * kCP + kk], Bs[c * kCP + kk], Bs[r * kCP + c]);
            }
            __syncthreads();
        }
        if (blockIdx.x == 0) {
            for (int idx = tid; idx < kb * kb; idx += kThreads) {
                const int r = idx / kb, c = idx - r * kb;
                if (c <= r) S[(size_t)(k0 + r) * m + k0 + c] = Bs[r * kCP + c];
            }
            if (tid < kb) q.Ld[k0 + tid] = Bs[tid * kCP + tid];
        }
        const int c0 = k0 + kb;
        {   // (2) rows below: thread t of CTA b takes row c0 + b + t gridDim.x
            const int r = c0 + (int)blockIdx.x + tid * (int)gridDim.x;
            if (r < m) {
                double* row = S + (size_t)r * m + k0;
                double x[kCB];
#pragma unroll
                for (int j = 0; j < kCB; ++j) x[j] = j < kb ? row[j] : 0.0;
#pragma unroll
                for (int j = 0; j < kCB; ++j) {
                    if (j < kb) {
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int i = 0; i < j; i += 2) {
                            a0 = fma(x[i], Bs[j * kCP + i], a0);
                            if (i + 1 < j) a1 = fma(x[i + 1], Bs[j * kCP + i + 1], a1);
                        }
                        x[j] = (x[j] - (a0 + a1)) / Bs[j * kCP + j];
                    }
                }
#pragma unroll
                for (int j = 0; j < kCB; ++j)
                    if (j < kb) row[j] = x[j];
            }
        }
        grid.sync();
        if (c0 < m) {  // (3) S[r][c] -= sum_k L[r][k] L[c][k] for c0 <= c <= r, in 64 x 64 tiles (I >= J)
            const int nt = (m - c0 + kTile - 1) / kTile;
            const int ntiles = nt * (nt + 1) / 2;
            const int ty = tid >> 4, tx = tid & 15;
            for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
                int I = 0;
                while ((I + 1) * (I + 2) / 2 <= tl) ++I;
                const int J = tl - I * (I + 1) / 2;
                const int r0 = c0 + I * kTile, q0 = c0 + J * kTile;
                __syncthreads();
                for (int idx = tid; idx < kTile * kCB; idx += kThreads) {
                    const int rr = idx / kCB, kk = idx - rr * kCB;
                    Li[rr * kCP + kk] = (r0 + rr < m && kk < kb) ? S[(size_t)(r0 + rr) * m + k0 + kk] : 0.0;
                    Lj[rr * kCP + kk] = (q0 + rr < m && kk < kb) ? S[(size_t)(q0 + rr) * m + k0 + kk] : 0.0;
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int w = 0; w < 4; ++w) acc[u][w] = 0.0;
#pragma unroll 4
                for (int kk = 0; kk < kCB; ++kk) {
                    double a[4], bb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { a[u] = Li[(ty + 16 * u) * kCP + kk]; bb[u] = Lj[(tx + 16 * u) * kCP + kk]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w = 0; w < 4; ++w) acc[u][w] = fma(a[u], bb[w], acc[u][w]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + ty + 16 * u;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int c = q0 + tx + 16 * w;
                        if (r < m && c <= r) S[(size_t)r * m + c] -= acc[u][w];
                    }
                }
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0) {
        // forward solve L u = z  (xw <- u): blocked, vector in shared memory (the panel buffer is free again)
        if (sm.tri) {
            double* vu = sm.tri;
            for (int r = tid; r < m; r += kThreads) vu[r] = sm.z[r];
            __syncthreads();
            tri_solve_lower_rows(S, (size_t)m, q.Ld, m, vu, sm.tri + 2 * m);
            for (int r = tid; r < m; r += kThreads) sm.xw[r] = vu[r];
            __syncthreads();
        } else {
        for (int r = tid; r < m; r += kThreads) sm.xw[r] = sm.z[r];
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            if (warp == 0) {
                double acc = 0.0;
                for (int c = lane; c < k; c += 32) acc = fma(S[(size_t)k * m + c], sm.xw[c], acc);
                acc = warp_sum(acc);
                if (lane == 0) sm.xw[k] = (sm.xw[k] - acc) / q.Ld[k];
            }
            __syncthreads();
        }
        }
        double part = 0.0;
        for (int r = tid; r < m; r += kThreads) part = fma(sm.xw[r], sm.xw[r], part);
        const double sigma = sqrt(block_sum(part, sm.red) / m);
        if (err_out)
            for (int i = tid; i < d; i += kThreads) err_out[i] = dt * (sqrt(sm.y[i]) * sigma);
        __syncthreads();
    }
}

// update_stage (ek1_device.cuh) on the grid.  Ends with a grid barrier.
static __device__ void update_stage_large(cg::grid_group& grid, const Problem& P, int b, const Smem& sm, const LargeQR& q,
                                   const LargeSmem& ls, int mcur, EMode emode, double nugget, const double* Rsrc,
                                   const int32_t* te, const int32_t* be, const int32_t* Hcol, const double* Hval, double* W,
                                   const UpdateOut out, double* diff_cta0, PhaseClock& pc, int32_t* nf = nullptr) {
    if (!nf) nf = q.nf;
    const int warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kWarps + warp, gnw = gridDim.x * kWarps;
    const int D = P.D, ld = P.ld;
    const int nbot = emode == E_NONE ? 0 : mcur;
    const int nrows = D + nbot;
    double* Wl = W + (size_t)(P.m - mcur) * ld;
    double* Wr = W + (size_t)P.m * ld;
    update_build_right(P, mcur, nrows, Rsrc, te, be, Wr, gw, gnw);
    grid.sync();
    update_build_left(P, b, mcur, nrows, emode, nugget, te, be, Hcol, Hval, Wl, Wr, gw, gnw);
    grid.sync();
    Shape sh;
    sh.nt = D; sh.nbot = nbot; sh.ncols = mcur + D; sh.te = te; sh.be = be; sh.ldr = ld;  // tile-aligned panel row lists
    pc.mark(4);
    householder_qr_large(grid, Wl, ld, sh, q, ls, pc);
    int bad = 0;
    if (blockIdx.x == 0) {
        const double diff = update_solve(P, sm, mcur, Wl, Wr);
        if (threadIdx.x == 0) {
            if (out.diff_out) *out.diff_out = diff;
            if (diff_cta0) *diff_cta0 = diff;
        }
        bad = update_output_mean(P, sm, out, diff);
    }
    pc.mark(6);
    bad |= update_output_factor(P, sm, out, mcur, nrows, Wr, gw, gnw);
    if (bad) atomicOr(nf, 1);
    grid.sync();
    pc.mark(7);
}

// One EK1 step of member b on the grid (white.py:96-146 / latent.py:167-223): predict mean + linearisation on CTA 0,
// predict-stack build, QR, error estimate, update.  ls.pv / ls.pinv hold the Nordsieck preconditioner of dt.  Ends with a
// grid barrier; *diff_cta0 (shared memory of CTA 0) receives the local diffusion.
static __device__ void ek1_step_large(cg::grid_group& grid, const Problem& P, int b, const Smem& sm, const LargeQR& q,
                                      const LargeSmem& ls, double dt, const double* min_, const double* cin_, double* mout,
                                      double* cout, double* err_out, double* ref_out, int flags, double* diff_cta0,
                                      PhaseClock& pc, int32_t* nf = nullptr) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const int gw = blockIdx.x * kWarps + warp, gnw = gridDim.x * kWarps;
    const int n = P.n, D = P.D;
    double* W = P.W;
    // [predict mean + linearisation] on CTA 0   white.py:104-113
    if (blockIdx.x == 0) {
        for (int k = tid; k < D; k += kThreads) {
            const int j = k / n, i = k - j * n;
            double acc = 0.0;
            for (int r = 0; r < n; ++r) acc = fma(P.A1d[i * n + r], ls.pinv[r] * min_[(size_t)r * P.dd + j], acc);
            sm.mp[k] = acc;
        }
        __syncthreads();
        evaluate_ode(P, b, sm, ls.pv[0], ls.pv[1], P.Hcol, P.Hval);
    }
    const bool dense = flags & 1;
    pc.mark(0);
    build_predict(P, b, sm, cin_, dense ? P.te_pd : P.te_p, W + (size_t)P.m * P.ld, gw, gnw);
    grid.sync();
    pc.mark(1);
    Shape sp;
    sp.nt = D; sp.nbot = D; sp.ncols = D; sp.te = dense ? P.te_pd : P.te_p; sp.be = P.be_p; sp.ldr = P.ld;
    householder_qr_large(grid, W + (size_t)P.m * P.ld, P.ld, sp, q, ls, pc);
    if (!P.latent && !(flags & 2))
        error_estimate_large(grid, P, b, sm, q, ls, ls.pv[1], dt, E_STEP_WHITE, P.Hcol, P.Hval, P.F, P.S, err_out);
    pc.mark(3);
    UpdateOut out;
    out.mean_out = mout; out.chol_out = cout; out.diff_out = nullptr;
    out.ref_out = P.latent ? nullptr : ref_out; out.scale_by_p = true;
    update_stage_large(grid, P, b, sm, q, ls, P.m, P.latent ? E_NONE : E_STEP_WHITE, 0.0, nullptr, P.te_u, P.be_u,
                       P.Hcol, P.Hval, W, out, diff_cta0, pc, nf);
}

#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_LARGE_RUN)
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_run_large(const Problem P, const RunArgs a, const __grid_constant__ LargeQR q) {
    extern __shared__ __align__(16) double smem_raw[];
    cg::grid_group grid = cg::this_grid();
    LargeSmem ls = carve_large(smem_raw);
    __shared__ __align__(8) unsigned long long tma_bars[2];
    __shared__ unsigned tma_nload;
    large_init_barriers(ls, tma_bars, &tma_nload);
    const Smem sm = large_vectors(P, q, ls);
    __shared__ double diff_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int gw = blockIdx.x * kWarps + warp, gnw = gridDim.x * kWarps;
    const size_t gtid = (size_t)blockIdx.x * kThreads + tid, gnt = (size_t)gridDim.x * kThreads;
    const int n = P.n, D = P.D;
    const size_t msz = (size_t)D, csz = (size_t)D * D;
    double* W = P.W;
    if (blockIdx.x == 0 && tid == 0) q.nf[2] = (int32_t)cg::this_cluster().num_blocks();  // diagnostics: pnmol_b200_cluster_size
    PhaseClock pc;  // phase cycles as seen by CTA 0 (pnmol_b200_profile); indices as in ek1_step, 8..13 = QR: panel
    pc.start(blockIdx.x == 0 ? P.prof : nullptr);  // factor, barrier, partial Y, barrier, update, barrier
    for (int b = 0; b < P.batch; ++b) {
        if (blockIdx.x == 0 && tid == 0) *q.nf = 0;
        double diffsum = 0.0;
        for (int s = 0; s < a.nsteps; ++s) {
            double dt;
            __syncthreads();
            if (a.nsteps == 1 && a.pv == nullptr) {
                if (tid < n) { ls.pv[tid] = a.pv0[tid]; ls.pinv[tid] = a.pinv0[tid]; }
                dt = a.dt0;
            } else {
                if (tid < n) { ls.pv[tid] = a.pv[(size_t)s * n + tid]; ls.pinv[tid] = a.pinv[(size_t)s * n + tid]; }
                dt = a.dts[s];
            }
            __syncthreads();
            const bool even = (s & 1) == 0;
            const double* min_ = (even ? a.mean_a : a.mean_b) + b * msz;
            const double* cin_ = (even ? a.chol_a : a.chol_b) + b * csz;
            double* mout = (even ? a.mean_b : a.mean_a) + b * msz;
            double* cout = (even ? a.chol_b : a.chol_a) + b * csz;
            const int flags = s == 0 ? a.flags : (a.flags & ~1);
            ek1_step_large(grid, P, b, sm, q, ls, dt, min_, cin_, mout, cout, a.err_out ? a.err_out + (size_t)b * P.d : nullptr,
                           a.ref_out ? a.ref_out + (size_t)b * P.d : nullptr, flags, &diff_s, pc);
            if (blockIdx.x == 0) { __syncthreads(); diffsum += diff_s; }
            if (a.mean_traj) {
                double* mt = a.mean_traj + ((size_t)s * P.batch + b) * msz;
                for (size_t k = gtid; k < msz; k += gnt) mt[k] = mout[k];
            }
            if (a.chol_traj) {
                double* ct = a.chol_traj + ((size_t)s * P.batch + b) * csz;
                for (size_t k = gtid; k < csz; k += gnt) ct[k] = cout[k];
            }
            if (a.std_traj) marginal_std_rows(cout, D, n, P.dd, a.std_traj + ((size_t)s * P.batch + b) * P.dd, gw, gnw);
        }
        if ((a.nsteps & 1) && !a.final_in_b) {  // result sits in b: bring it home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = gtid; k < msz; k += gnt) md[k] = ms[k];
            for (size_t k = gtid; k < csz; k += gnt) cd[k] = cs[k];
        }
        if (blockIdx.x == 0 && tid == 0) {
            if (a.diff_last) a.diff_last[b] = diff_s;
            if (a.diff_sum) a.diff_sum[b] = diffsum;
            if (a.status) a.status[b] = *q.nf;
        }
        if (__ldcg(q.nf)) {  // the tile-aligned row lists read padding rows (times zero): leave no NaNs behind for the next member
            for (size_t k = gtid; k < (size_t)P.ld * (P.m + P.D); k += gnt) W[k] = 0.0;
        }
        grid.sync();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_run_large(const Problem P, const RunArgs a, const __grid_constant__ LargeQR q);
#endif

// Adaptive time loop on the grid (src/pnmol/pdefilter.py:192-227, src/pnmol/odetools/step.py:58-119) for the multi-CTA
// path: one member at a time, every CTA evaluates the scaled error norm and the step-size proposal redundantly from
// the error estimate / reference state that the step left in global memory (same summation order everywhere, so all
// CTAs take the same branch -- no broadcast).  The non-finite flag alternates between two words by attempt parity, so
// that a rejected non-finite proposal can be forgotten without another grid barrier.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_LARGE_ADAPTIVE)
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_run_adaptive_large(const Problem P, const AdaptiveArgs a, const __grid_constant__ LargeQR q) {
    extern __shared__ __align__(16) double smem_raw[];
    cg::grid_group grid = cg::this_grid();
    LargeSmem ls = carve_large(smem_raw);
    __shared__ __align__(8) unsigned long long tma_bars[2];
    __shared__ unsigned tma_nload;
    large_init_barriers(ls, tma_bars, &tma_nload);
    const Smem sm = large_vectors(P, q, ls);
    __shared__ double diff_s;
    const int tid = threadIdx.x;
    const size_t gtid = (size_t)blockIdx.x * kThreads + tid, gnt = (size_t)gridDim.x * kThreads;
    const int nu = P.n - 1;
    const size_t msz = (size_t)P.D, csz = (size_t)P.D * P.D;
    PhaseClock pc;
    pc.start(nullptr);
    for (int b = 0; b < P.batch; ++b) {
        if (blockIdx.x == 0 && tid == 0) { q.nf[0] = 0; q.nf[1] = 0; }
        grid.sync();
        double t = a.t0, dt = a.dt0[b], diffsum = 0.0, difflast = 0.0;
        int nsteps = 0, natt = 0, cur = 0, stat = 0, bad = 0;
        double* err = a.err + (size_t)b * P.d;
        double* ref = a.ref + (size_t)b * P.d;
        while (t < a.tmax) {
            if (natt >= a.max_attempts) { stat |= 2; break; }
            if (!(dt >= 0.0)) { stat |= 1; break; }  // pdefilter.py:225 asserts dt >= 0 (a NaN proposal ends here too)
            __syncthreads();
            if (tid < P.n) {  // Nordsieck preconditioner p_i = |dt|^(nu - i + 1/2) / (nu - i)!   (iwp.py:55-62)
                const int k = nu - tid;
                double fact = 1.0;
                for (int qq = 2; qq <= k; ++qq) fact *= qq;
                const double pw = pow(fabs(dt), k + 0.5);
                ls.pv[tid] = pw / fact;
                ls.pinv[tid] = fact / pw;
            }
            __syncthreads();
            const double* min_ = (cur ? a.mean_b : a.mean_a) + b * msz;
            const double* cin_ = (cur ? a.chol_b : a.chol_a) + b * csz;
            double* mout = (cur ? a.mean_a : a.mean_b) + b * msz;
            double* cout = (cur ? a.chol_a : a.chol_b) + b * csz;
            const int flags = (natt == 0 || nsteps == 0) ? a.flags : (a.flags & ~1);
            int32_t* nf = q.nf + (natt & 1);
            ek1_step_large(grid, P, b, sm, q, ls, dt, min_, cin_, mout, cout, err, ref, flags, &diff_s, pc, nf);   // ends with a grid barrier
            const int badstep = __ldcg(nf);
            if (blockIdx.x == 0 && tid == 0) q.nf[(natt + 1) & 1] = 0;   // the other word serves the next attempt
            // scaled error norm (step.py:97-108 on dt * error_estimate, pdefilter.py:208-213)
            double part = 0.0;
            for (int i = tid; i < P.d; i += kThreads) {
                const double r = dt * __ldcg(err + i) / (a.abstol + a.reltol * __ldcg(ref + i));
                part = fma(r, r, part);
            }
            const double norm = sqrt(block_sum(part, ls.red)) / sqrt((double)P.d);
            double change = a.safety * pow(1.0 / norm, a.inv_rate);
            change = fmax(a.change_min, fmin(change, a.change_max));
            if (!(norm == norm)) change = norm;  // NaN propagates like jnp.minimum / jnp.maximum
            const double suggested = change * dt;
            ++natt;
            if (norm < 1.0) {  // accepted: the proposal becomes the state
                t = t + dt;
                cur ^= 1;
                ++nsteps;
                bad |= badstep;
                if (blockIdx.x == 0) { difflast = diff_s; diffsum += diff_s; }
                if (a.mean_traj) {
                    if (nsteps <= a.max_traj) {
                        double* mt = a.mean_traj + ((size_t)(nsteps - 1) * P.batch + b) * msz;
                        double* ct = a.chol_traj + ((size_t)(nsteps - 1) * P.batch + b) * csz;
                        for (size_t k = gtid; k < msz; k += gnt) mt[k] = mout[k];
                        for (size_t k = gtid; k < csz; k += gnt) ct[k] = cout[k];
                        if (gtid == 0) a.t_traj[(size_t)b * a.max_traj + nsteps - 1] = t;
                    } else {
                        stat |= 4;
                    }
                }
            }
            dt = fmin(suggested, a.tmax - t);
            if (!(suggested == suggested)) dt = suggested;
        }
        grid.sync();
        if (cur) {  // bring the final state home
            const double* ms = a.mean_b + b * msz; const double* cs = a.chol_b + b * csz;
            double* md = a.mean_a + b * msz; double* cd = a.chol_a + b * csz;
            for (size_t k = gtid; k < msz; k += gnt) md[k] = ms[k];
            for (size_t k = gtid; k < csz; k += gnt) cd[k] = cs[k];
        }
        if (blockIdx.x == 0 && tid == 0) {
            a.t_out[b] = t; a.dt_out[b] = dt; a.diff_sum[b] = diffsum; a.diff_last[b] = difflast;
            a.nsteps[b] = nsteps; a.nattempts[b] = natt; a.status[b] = stat | (bad ? 1 : 0);
        }
        if (bad || stat || __ldcg(q.nf) || __ldcg(q.nf + 1)) {  // no NaNs in the workspace for the next member
            for (size_t k = gtid; k < (size_t)P.ld * (P.m + P.D); k += gnt) P.W[k] = 0.0;
        }
        grid.sync();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_run_adaptive_large(const Problem P, const AdaptiveArgs a, const __grid_constant__ LargeQR q);
#endif

// initialize() (white.py:12-80, latent.py:20-134) on the grid.
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_LARGE_INIT)
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_init_large(const Problem P, const InitArgs a, const __grid_constant__ LargeQR q) {
    extern __shared__ __align__(16) double smem_raw[];
    cg::grid_group grid = cg::this_grid();
    LargeSmem ls = carve_large(smem_raw);
    __shared__ __align__(8) unsigned long long tma_bars[2];
    __shared__ unsigned tma_nload;
    large_init_barriers(ls, tma_bars, &tma_nload);
    const Smem sm = large_vectors(P, q, ls);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * kWarps + warp, gnw = gridDim.x * kWarps;
    const int n = P.n, d = P.d, D = P.D, nd = P.n * P.d;
    double* W = P.W;
    PhaseClock pc;
    pc.start(nullptr);
    for (int b = 0; b < P.batch; ++b) {
        if (blockIdx.x == 0 && tid == 0) *q.nf = 0;
        double* chol = a.chol_out + (size_t)b * D * D;
        double* mean = a.mean_out + (size_t)b * D;
        const double ps = P.priorscale ? P.priorscale[b] : 1.0;
        for (int r = gw; r < D; r += gnw) {  // C0 = kron(Lk, c0 I_n), latent: blockdiag(., kron(E_sqrtm, c0 I_n))
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
        if (blockIdx.x == 0) {  // update on the initial condition: H = E0, z = -y0   white.py:32-39
            for (int k = tid; k < D; k += kThreads) sm.mp[k] = 0.0;
            for (int i = tid; i < d; i += kThreads) {
                sm.z[i] = -a.y0[(size_t)b * d + i];
                for (int w = 0; w < P.wh; ++w) { P.Hcol[(size_t)i * P.wh + w] = w == 0 ? i * n : -1; P.Hval[(size_t)i * P.wh + w] = w == 0 ? 1.0 : 0.0; }
            }
        }
        __syncthreads();
        if (tid < n) { ls.pv[tid] = 1.0; ls.pinv[tid] = 1.0; }
        __syncthreads();
        grid.sync();
        UpdateOut o1;
        o1.mean_out = nullptr; o1.chol_out = chol; o1.diff_out = nullptr; o1.ref_out = nullptr; o1.scale_by_p = false;
        update_stage_large(grid, P, b, sm, q, ls, d, E_NUGGET_ONLY, a.nugget, chol, nullptr, nullptr, P.Hcol, P.Hval, W, o1, nullptr, pc);
        if (blockIdx.x == 0) evaluate_ode(P, b, sm, 1.0, 1.0, P.Hcol, P.Hval);  // white.py:42-48, latent.py:86-95
        grid.sync();
        UpdateOut o2;
        o2.mean_out = mean; o2.chol_out = chol; o2.diff_out = nullptr; o2.ref_out = nullptr; o2.scale_by_p = false;
        update_stage_large(grid, P, b, sm, q, ls, P.m, P.latent ? E_NUGGET_ONLY : E_STEP_PLUS_NUGGET, a.nugget, chol, nullptr,
                           nullptr, P.Hcol, P.Hval, W, o2, nullptr, pc);
        if (blockIdx.x == 0 && tid == 0 && a.status) a.status[b] = *q.nf;
        if (__ldcg(q.nf)) {  // as in k_run_large: no NaNs in the padding rows for the next member
            const size_t gtid = (size_t)blockIdx.x * kThreads + tid, gnt = (size_t)gridDim.x * kThreads;
            for (size_t k = gtid; k < (size_t)P.ld * (P.m + P.D); k += gnt) W[k] = 0.0;
        }
        grid.sync();
    }
}
#else
__global__ void __launch_bounds__(kThreads, PNMOL_LARGE_CTAS) k_init_large(const Problem P, const InitArgs a, const __grid_constant__ LargeQR q);
#endif

}  // namespace pnmol
