// Multi-CTA blocked Householder QR for large state dimension (north-star kernel 2), sm_100a.
//
// Same mathematics, LAPACK dlarfg conventions and support envelopes as the single-CTA routines
// (householder_columns / householder_qr_blocked); what changes is who does what.  The whole grid
// (one CTA per SM, cooperative launch) works on ONE matrix that lives in the L2/HBM workspace:
//
//   per panel of kNB columns
//     S1  CTA 0 factors the panel in shared memory (the panel's compact row list x kNB columns; when
//         that does not fit, power-of-two sub-panels are factored in shared memory and applied to the
//         rest of the panel from L2), writes R back, the reflectors V to global (Vg, reflector-major,
//         compact row list) and the compact-WY T factor (dlarft, Gram matrix on the FP64 tensor pipe).
//     --  grid barrier
//     S2  every CTA takes (row chunk, 64-column group) items: stages the V chunk in shared memory and
//         each warp forms the partial  Y^T = C^T V  of its 8 columns over the chunk's rows with
//         mma.sync.m8n8k4.f64 (column data straight from global memory in the A-fragment layout);
//         partials go to global memory (no atomics: the sum order is fixed, results are reproducible).
//     --  grid barrier
//     S3  same items: sum the partials, Y' = Y^T T (DMMA), C^T -= Y' V^T (DMMA) on the chunk's rows.
//     --  grid barrier
#pragma once
#include <cooperative_groups.h>
#include <cuda.h>

namespace pnmol {

namespace cg = cooperative_groups;

constexpr int kTmaBoxes = 3;                         // tensor maps of the reflector buffer: boxes of 82 / 130 / 242 rows x kNB
__host__ __device__ constexpr int large_box_pitch(int i) { return i == 0 ? 82 : i == 1 ? 130 : 242; }   // = 2 (mod 16)
constexpr int kTmaStage = kNB * 242;                 // doubles per TMA stage (30976 bytes)

struct LargeQR {   // global scratch of the multi-CTA path (one set per handle); passed as a __grid_constant__ parameter
    alignas(64) CUtensorMap tmapV[kTmaBoxes];  // 2-D tensor maps of Vg: dim0 = compact row (lv), dim1 = reflector
    double* Vg;    // [kNB][lv]  reflectors of the current panel
    double* Tg;    // [kNB][kLdr] T factor of the current panel
    double* Yp;    // [row chunk][ycols][kNB] partial Y^T
    double* vec;   // mp | z | y | xw | xat  (the vectors the single-CTA path keeps in shared memory)
    double* Ld;    // [m] diagonal of the Cholesky factor of S (error estimate)
    int32_t* nf;   // non-finite flag of the member in flight
    int lv, ycols, cap;  // cap: panel buffer capacity (doubles of shared memory)
    int cb;              // panel width of the blocked Cholesky in the error estimate (16, 8 or 4)
};

struct LargeSmem {
    double *red, *pv, *pinv, *sc, *Ts, *Gs, *scratch, *PB;
    unsigned long long* bars;   // two mbarriers of the TMA stages (initialised once per kernel)
    unsigned* nload;            // TMA loads issued so far by this CTA (stage and phase parity of the next one)
};

// ---------------------------------------------------------------- TMA / mbarrier helpers
__device__ __forceinline__ unsigned large_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void large_tma_load(double* dst, const CUtensorMap* tmap, int x, int y, unsigned long long* bar, unsigned bytes) {
    const unsigned b = large_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(large_smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(b) : "memory");
}
__device__ __forceinline__ void large_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = large_smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}
// Once per kernel: the two mbarriers and the load counter.
__device__ __forceinline__ void large_init_barriers(LargeSmem& ls, unsigned long long* bars, unsigned* nload) {
    ls.bars = bars;
    ls.nload = nload;
    if (threadIdx.x == 0) {
        for (int j = 0; j < 2; ++j)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(large_smem_u32(&bars[j])) : "memory");
        *nload = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
}
// Which tensor map serves chunks of RC rows (RC <= 240), and the two stage buffers (128-byte aligned) in the panel buffer.
__device__ __forceinline__ int large_box_index(int RC) { return RC + 2 <= 82 ? 0 : RC + 2 <= 130 ? 1 : 2; }
__device__ __forceinline__ double* large_stage(const LargeSmem& ls, int i) {
    return reinterpret_cast<double*>((reinterpret_cast<size_t>(ls.PB) + 127) & ~(size_t)127) + (size_t)i * kTmaStage;
}
constexpr int kLargeFixed = 16 + 2 * kMaxN + 80 + 288 + 272 + kWarps * 192;

__device__ __forceinline__ LargeSmem carve_large(double* base) {
    LargeSmem s;
    s.red = base;      base += 16;
    s.pv = base;       base += kMaxN;
    s.pinv = base;     base += kMaxN;
    s.sc = base;       base += 80;
    s.Ts = base;       base += 288;
    s.Gs = base;       base += 272;
    s.scratch = base;  base += kWarps * 192;
    s.PB = base;
    return s;
}

__device__ __forceinline__ int row_cidx(const RowMap& rm, int row) { return row < rm.a2 ? row - rm.j0 : rm.len1 + (row - rm.a2); }

// Gram matrix of the panel's reflectors on the tensor pipe (upper triangle, Gs[k * 17 + i], k <= i), V(refl, c) =
// V[refl * ldv + c] zero padded to a multiple of 8 rows; per-warp partial blocks are summed in a fixed order.
static __device__ void large_gram(const double* __restrict__ V, int ldv, int len, double* __restrict__ Gs, double* __restrict__ scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (len + 7) >> 3;
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    double d00[2] = {0.0, 0.0}, d01[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
    for (int i = warp; i < ntile; i += kWarps) {
        const double* v0 = V + (size_t)g * ldv + 8 * i + 2 * t;
        const double* v1 = v0 + (size_t)8 * ldv;
        const double a0 = v0[0], a1 = v1[0], b0 = v0[1], b1 = v1[1];
        dmma884(c00[0], c00[1], a0, a0);
        dmma884(c01[0], c01[1], a0, a1);
        dmma884(c11[0], c11[1], a1, a1);
        dmma884(d00[0], d00[1], b0, b0);
        dmma884(d01[0], d01[1], b0, b1);
        dmma884(d11[0], d11[1], b1, b1);
    }
    double* mine = scratch + warp * 192;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = g * 8 + 2 * t + q;
        mine[e] = c00[q] + d00[q];
        mine[64 + e] = c01[q] + d01[q];
        mine[128 + e] = c11[q] + d11[q];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 192; e += kThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sum += scratch[w * 192 + e];
        const int blk = e >> 6, r = (e & 63) >> 3, c = e & 7;
        Gs[(r + (blk == 2 ? 8 : 0)) * 17 + c + (blk >= 1 ? 8 : 0)] = sum;
    }
    __syncthreads();
}

// S1: factor the panel (columns j0 .. j0+nbk-1) on one CTA.
//
// Two block barriers per column: the warp that updates the next column also reduces that column's norm and diagonal
// entry (published in shared memory), so every thread derives the next reflector's scalars without another
// reduction round; each warp updates two panel columns at once (one pass over the reflector).
static __device__ void large_panel_factor(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk, const RowMap rm,
                                   const LargeQR& q, const LargeSmem& ls, PhaseClock& pc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = rm.len, lp = (L + 7) & ~7, nt = s.nt;
    double* PB = ls.PB;
    double* sc = ls.sc;          // [3 i] tau, [3 i + 1] beta
    double* nxt = ls.sc + 64;    // [0] ||x||^2 below the diagonal of the next column, [1] its diagonal entry
    __shared__ int env[2 * kNB];  // own envelope of every panel column: last top row, last bottom row
    int sw = kNB;
    while (sw > 1 && (size_t)sw * lp > (size_t)q.cap) sw >>= 1;
    if (tid < nbk) { env[tid] = env_top(s, j0 + tid); env[kNB + tid] = env_bot(s, j0 + tid); }
    // reflector slots beyond the panel are zero
    for (int idx = tid; idx < (kNB - nbk) * lp; idx += kThreads) q.Vg[(size_t)(nbk + idx / lp) * q.lv + idx % lp] = 0.0;
    __syncthreads();
    for (int i0 = 0; i0 < nbk; i0 += sw) {
        const int ncs = nbk - i0 < sw ? nbk - i0 : sw;
        // ---- load the sub-panel with 8-byte asynchronous copies (global -> shared, no register staging: every thread
        // has all its copies in flight at once); entries outside a column's own envelope are zero
        for (int cc = 0; cc < ncs; ++cc) {
            const int et = env[i0 + cc], eb = env[kNB + i0 + cc];
            const double* src = W + (size_t)(j0 + i0 + cc) * ld;
            double* dstc = PB + (size_t)cc * lp;
            for (int c = tid; c < lp; c += kThreads) {
                const int row = rm.row(c);
                if (c < L && (row < nt ? row <= et : row <= eb)) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(dstc + c);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + row) : "memory");
                } else {
                    dstc[c] = 0.0;
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();
        pc.mark(14);
        {   // norm and diagonal of the sub-panel's first column
            const int cd = row_cidx(rm, j0 + i0);
            double ss = 0.0;
            for (int c = cd + 1 + tid; c < L; c += kThreads) ss = fma(PB[c], PB[c], ss);
            ss = block_sum(ss, ls.red);
            if (tid == 0) { nxt[0] = ss; nxt[1] = PB[cd]; }
            __syncthreads();
        }
        pc.mark(15);
        // ---- factor it column by column
        for (int cc = 0; cc < ncs; ++cc) {
            const int i = i0 + cc;
            const int cd = row_cidx(rm, j0 + i);
            double* x = PB + (size_t)cc * lp;
            const double ss = nxt[0], alpha = nxt[1];
            double tau = 0.0, beta = alpha, scale = 0.0;
            if (ss != 0.0) {  // dlarfg: xnorm == 0 -> H = I
                const double nrm = sqrt(fma(alpha, alpha, ss));
                beta = -copysign(nrm, alpha);  // Fortran SIGN semantics of dlarfg
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (tau != 0.0) {
                int c = cd + 1 + tid;
                for (; c + 3 * kThreads < L; c += 4 * kThreads) {
                    const double v0 = x[c], v1 = x[c + kThreads], v2 = x[c + 2 * kThreads], v3 = x[c + 3 * kThreads];
                    x[c] = v0 * scale; x[c + kThreads] = v1 * scale; x[c + 2 * kThreads] = v2 * scale; x[c + 3 * kThreads] = v3 * scale;
                }
                for (; c < L; c += kThreads) x[c] *= scale;
            }
            if (tid == 0) { x[cd] = 1.0; sc[3 * i] = tau; sc[3 * i + 1] = beta; }
            __syncthreads();  // v complete; nxt consumed by everyone
            // apply H to the remaining columns of the sub-panel, two per warp; the first one is the next pivot column
            for (int k0 = cc + 1 + warp; k0 < ncs; k0 += 2 * kWarps) {
                const int k1 = k0 + kWarps;
                const bool two = k1 < ncs;
                double* y0 = PB + (size_t)k0 * lp;
                double* y1 = PB + (size_t)(two ? k1 : k0) * lp;
                double wa = 0.0, wb = 0.0;
                if (tau != 0.0) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
                    int c = cd + lane;
                    for (; c + 96 < L; c += 128) {
                        const double x0 = x[c], x1 = x[c + 32], x2 = x[c + 64], x3 = x[c + 96];
                        a0 = fma(x0, y0[c], a0); a1 = fma(x1, y0[c + 32], a1); a2 = fma(x2, y0[c + 64], a2); a3 = fma(x3, y0[c + 96], a3);
                        b0 = fma(x0, y1[c], b0); b1 = fma(x1, y1[c + 32], b1); b2 = fma(x2, y1[c + 64], b2); b3 = fma(x3, y1[c + 96], b3);
                    }
                    for (; c < L; c += 32) { a0 = fma(x[c], y0[c], a0); b0 = fma(x[c], y1[c], b0); }
                    a0 = (a0 + a1) + (a2 + a3); b0 = (b0 + b1) + (b2 + b3);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ta = __shfl_xor_sync(0xffffffffu, a0, o), tb = __shfl_xor_sync(0xffffffffu, b0, o);
                        a0 += ta; b0 += tb;
                    }
                    wa = tau * a0; wb = tau * b0;
                }
                const bool pivot = k0 == cc + 1;  // this warp owns the next pivot column: reduce its norm on the fly
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int c = cd + lane;
                for (; c + 96 < L; c += 128) {  // loads first: the compiler cannot prove x, y0, y1 distinct
                    const double x0 = x[c], x1 = x[c + 32], x2 = x[c + 64], x3 = x[c + 96];
                    const double p0 = y0[c], p1 = y0[c + 32], p2 = y0[c + 64], p3 = y0[c + 96];
                    const double q0 = y1[c], q1 = y1[c + 32], q2 = y1[c + 64], q3 = y1[c + 96];
                    const double r0 = fma(-wa, x0, p0), r1 = fma(-wa, x1, p1), r2 = fma(-wa, x2, p2), r3 = fma(-wa, x3, p3);
                    y0[c] = r0; y0[c + 32] = r1; y0[c + 64] = r2; y0[c + 96] = r3;
                    if (two) {
                        y1[c] = fma(-wb, x0, q0); y1[c + 32] = fma(-wb, x1, q1); y1[c + 64] = fma(-wb, x2, q2); y1[c + 96] = fma(-wb, x3, q3);
                    }
                    if (pivot) {
                        if (c > cd + 1) s0 = fma(r0, r0, s0);
                        s1 = fma(r1, r1, s1); s2 = fma(r2, r2, s2); s3 = fma(r3, r3, s3);
                    }
                }
                for (; c < L; c += 32) {
                    const double xv = x[c];
                    const double ya = fma(-wa, xv, y0[c]);
                    y0[c] = ya;
                    if (two) y1[c] = fma(-wb, xv, y1[c]);
                    if (pivot && c > cd + 1) s0 = fma(ya, ya, s0);
                }
                if (pivot) {
                    const double ssn = warp_sum((s0 + s1) + (s2 + s3));
                    __syncwarp();
                    if (lane == 0) { nxt[0] = ssn; nxt[1] = y0[cd + 1]; }
                }
            }
            __syncthreads();
        }
        pc.mark(16);
        // ---- write R back, export V (the buffer keeps the pure reflectors: zero above the unit diagonal)
        for (int cc = 0; cc < ncs; ++cc) {
            const int i = i0 + cc, j = j0 + i;
            const int cd = row_cidx(rm, j), et = env[i], eb = env[kNB + i];
            const double tau = sc[3 * i], beta = sc[3 * i + 1];
            double* colp = PB + (size_t)cc * lp;
            double* wj = W + (size_t)j * ld;
            double* vg = q.Vg + (size_t)i * q.lv;
            for (int c = tid; c < lp; c += kThreads) {
                const double xv = colp[c];
                const int row = rm.row(c);
                if (c < L && (row < nt ? row <= et : row <= eb)) wj[row] = c < cd ? xv : (c == cd ? beta : 0.0);
                const double vv = (tau != 0.0 && c >= cd && c < L) ? xv : 0.0;
                colp[c] = vv;
                vg[c] = vv;
            }
        }
        __syncthreads();
        pc.mark(17);
        // ---- apply the sub-panel's reflectors to the rest of the panel (streamed from L2), one warp per column
        if (i0 + ncs < nbk) {
            for (int k = j0 + i0 + ncs + warp; k < j0 + nbk; k += kWarps) {
                double* col = W + (size_t)k * ld;
                for (int cc = 0; cc < ncs; ++cc) {
                    const double tau = sc[3 * (i0 + cc)];
                    if (tau == 0.0) continue;
                    const double* v = PB + (size_t)cc * lp;
                    const int cd = row_cidx(rm, j0 + i0 + cc);
                    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                    int c = cd + lane;
                    for (; c + 96 < L; c += 128) {
                        const double a0 = col[rm.row(c)], a1 = col[rm.row(c + 32)], a2 = col[rm.row(c + 64)], a3 = col[rm.row(c + 96)];
                        d0 = fma(v[c], a0, d0); d1 = fma(v[c + 32], a1, d1); d2 = fma(v[c + 64], a2, d2); d3 = fma(v[c + 96], a3, d3);
                    }
                    for (; c < L; c += 32) d0 = fma(v[c], col[rm.row(c)], d0);
                    const double w = tau * warp_sum((d0 + d1) + (d2 + d3));
                    c = cd + lane;
                    for (; c + 96 < L; c += 128) {
                        const int r0 = rm.row(c), r1 = rm.row(c + 32), r2 = rm.row(c + 64), r3 = rm.row(c + 96);
                        const double a0 = col[r0], a1 = col[r1], a2 = col[r2], a3 = col[r3];
                        col[r0] = fma(-w, v[c], a0); col[r1] = fma(-w, v[c + 32], a1);
                        col[r2] = fma(-w, v[c + 64], a2); col[r3] = fma(-w, v[c + 96], a3);
                    }
                    for (; c < L; c += 32) { const int r0 = rm.row(c); col[r0] = fma(-w, v[c], col[r0]); }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
    }
    pc.mark(18);
    // ---- T factor
    large_gram(sw == kNB ? PB : q.Vg, sw == kNB ? lp : q.lv, L, ls.Gs, ls.scratch);
    if (warp == 0) panel_t_factor(ls.Gs, sc, nbk, ls.Ts);
    __syncthreads();
    for (int idx = tid; idx < kNB * kLdr; idx += kThreads) q.Tg[idx] = ls.Ts[idx];
    pc.mark(19);
}

// S1 on a thread-block cluster: the panel's compact row list is split over the CS CTAs of cluster 0 (CTA r keeps rows
// [r Lc, (r+1) Lc) of all kNB columns in its shared memory), so the per-column work runs on CS SMs.  Per column one
// round: every CTA reduces its partial dot products x_i . x_k (k >= i; k = i is the norm) over its rows, all CTAs
// exchange them through distributed shared memory (CTA 0 adds the row-i entries x_k[i]), one cluster barrier, then every
// CTA sums the partials in the same fixed order, derives the dlarfg scalars redundantly and updates its own rows.
// Requires Lc >= kNB (the diagonal block lives in CTA 0) -- the caller falls back to large_panel_factor otherwise.
constexpr int kGath = 2 * kNB;  // doubles per CTA and column round: kNB partial dots + kNB row-i entries

static __device__ void large_panel_factor_cluster(cg::cluster_group& cluster, double* __restrict__ W, int ld, const Shape& s, int j0,
                                           int nbk, const RowMap rm, const LargeQR& q, const LargeSmem& ls, int Lc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CS = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
    const int L = rm.len, nt = s.nt;
    const int c_lo = r * Lc, c_hi = L < (r + 1) * Lc ? L : (r + 1) * Lc;  // this CTA's rows of the compact list
    const int nloc = c_hi > c_lo ? c_hi - c_lo : 0;
    double* PB = ls.PB;                       // [kNB][Lc] local slice, column-major
    double* gath = PB + (size_t)kNB * Lc;     // [2][CS][kGath] gathered partials (double buffered)
    double* part = gath + 2 * CS * kGath;     // [kGath] this CTA's contribution
    double* sc = ls.sc;                       // [3 i] tau, [3 i + 1] beta, [3 i + 2] scale
    __shared__ int envc[2 * kNB];
    if (tid < nbk) { envc[tid] = env_top(s, j0 + tid); envc[kNB + tid] = env_bot(s, j0 + tid); }
    __syncthreads();
    // ---- load the slice (entries outside a column's own envelope are zero)
    for (int cc = 0; cc < kNB; ++cc) {
        const bool hc = cc < nbk;
        const int et = hc ? envc[cc] : -1, eb = hc ? envc[kNB + cc] : -1;
        const double* src = W + (size_t)(j0 + (hc ? cc : 0)) * ld;
        double* dstc = PB + (size_t)cc * Lc;
        for (int cl = tid; cl < Lc; cl += kThreads) {
            const int c = c_lo + cl;
            const int row = rm.row(c);
            if (hc && c < L && (row < nt ? row <= et : row <= eb)) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(dstc + cl);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + row) : "memory");
            } else {
                dstc[cl] = 0.0;
            }
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    for (int i = 0; i < nbk; ++i) {
        const double* xi = PB + (size_t)i * Lc;
        const int lo = (i + 1 > c_lo ? i + 1 : c_lo) - c_lo;  // first local row below the diagonal
        // ---- partial dot products: warp w takes columns w and w + 8 (both in one pass over the pivot column)
        {
            const int k0 = warp, k1 = warp + kWarps;
            const bool u0 = k0 >= i && k0 < nbk, u1 = k1 >= i && k1 < nbk;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            if (u0 || u1) {
                const double* xa = PB + (size_t)(u0 ? k0 : k1) * Lc;
                const double* xb = PB + (size_t)(u1 ? k1 : k0) * Lc;
                int cl = lo + lane;
                for (; cl + 32 < nloc; cl += 64) {
                    const double p0 = xi[cl], p1 = xi[cl + 32];
                    a0 = fma(p0, xa[cl], a0); a1 = fma(p1, xa[cl + 32], a1);
                    b0 = fma(p0, xb[cl], b0); b1 = fma(p1, xb[cl + 32], b1);
                }
                if (cl < nloc) { a0 = fma(xi[cl], xa[cl], a0); b0 = fma(xi[cl], xb[cl], b0); }
                a0 += a1; b0 += b1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ta = __shfl_xor_sync(0xffffffffu, a0, o), tb = __shfl_xor_sync(0xffffffffu, b0, o);
                    a0 += ta; b0 += tb;
                }
            }
            if (lane == 0) {
                part[k0] = u0 ? a0 : 0.0;
                part[k1] = u1 ? (u0 ? b0 : a0) : 0.0;
                part[kNB + k0] = (r == 0 && k0 < nbk) ? PB[(size_t)k0 * Lc + i] : 0.0;  // row i of column k lives in CTA 0
                part[kNB + k1] = (r == 0 && k1 < nbk) ? PB[(size_t)k1 * Lc + i] : 0.0;
            }
        }
        __syncthreads();
        // ---- all-gather through distributed shared memory
        double* gbuf = gath + (i & 1) * CS * kGath;
        for (int idx = tid; idx < CS * kGath; idx += kThreads) {
            const int dst = idx / kGath, e = idx - dst * kGath;
            double* remote = cluster.map_shared_rank(gbuf, dst);
            remote[r * kGath + e] = part[e];
        }
        cluster.sync();
        // ---- totals (same order everywhere; the warp's two columns summed alongside the norm), scalars, local update
        const int k0 = warp, k1 = warp + kWarps;
        const bool v0 = k0 > i && k0 < nbk, v1 = k1 > i && k1 < nbk;
        double ss = 0.0, d0 = 0.0, d1 = 0.0;
        for (int q2 = 0; q2 < CS; ++q2) {
            const double* gq = gbuf + q2 * kGath;
            ss += gq[i]; d0 += gq[k0]; d1 += gq[k1];
        }
        const double al = gbuf[kNB + i], e0 = gbuf[kNB + k0], e1 = gbuf[kNB + k1];
        double tau = 0.0, beta = al, scale = 0.0;
        if (ss != 0.0) {  // dlarfg: xnorm == 0 -> H = I
            const double nrm = sqrt(fma(al, al, ss));
            beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg
            tau = (beta - al) / beta;
            scale = 1.0 / (al - beta);
        }
        if (tid == 0) { sc[3 * i] = tau; sc[3 * i + 1] = beta; sc[3 * i + 2] = scale; }
        if (tau != 0.0 && (v0 || v1)) {
            const double g0 = v0 ? -tau * fma(scale, d0, e0) : 0.0, g1 = v1 ? -tau * fma(scale, d1, e1) : 0.0;
            const double gs0 = g0 * scale, gs1 = g1 * scale;
            double* xa = PB + (size_t)(v0 ? k0 : k1) * Lc;
            double* xb = PB + (size_t)k1 * Lc;
            const double ga = v0 ? gs0 : gs1;
            int cl = lo + lane;
            for (; cl + 32 < nloc; cl += 64) {
                const double p0 = xi[cl], p1 = xi[cl + 32];
                const double y0 = xa[cl], y1 = xa[cl + 32];
                xa[cl] = fma(ga, p0, y0); xa[cl + 32] = fma(ga, p1, y1);
                if (v0 && v1) {
                    const double z0 = xb[cl], z1 = xb[cl + 32];
                    xb[cl] = fma(gs1, p0, z0); xb[cl + 32] = fma(gs1, p1, z1);
                }
            }
            if (cl < nloc) {
                xa[cl] = fma(ga, xi[cl], xa[cl]);
                if (v0 && v1) xb[cl] = fma(gs1, xi[cl], xb[cl]);
            }
            if (r == 0 && lane == 0) {  // v = 1 at row i
                if (v0) PB[(size_t)k0 * Lc + i] += g0;
                if (v1) PB[(size_t)k1 * Lc + i] += g1;
            }
        }
        __syncthreads();
    }
    // ---- write R back, export V (v = scale x below the diagonal, 1 on it, 0 above) to Vg and keep it in the slice
    for (int cc = 0; cc < kNB; ++cc) {
        const bool hc = cc < nbk;
        const double tau = hc ? sc[3 * cc] : 0.0, beta = hc ? sc[3 * cc + 1] : 0.0, scale = hc ? sc[3 * cc + 2] : 0.0;
        const int et = hc ? envc[cc] : -1, eb = hc ? envc[kNB + cc] : -1;
        double* colp = PB + (size_t)cc * Lc;
        double* wj = W + (size_t)(j0 + (hc ? cc : 0)) * ld;
        double* vg = q.Vg + (size_t)cc * q.lv;
        for (int cl = tid; cl < Lc; cl += kThreads) {
            const int c = c_lo + cl;
            const double xv = colp[cl];
            double vv = 0.0;
            if (hc && c < L) {
                const int row = rm.row(c);
                if (row < nt ? row <= et : row <= eb) wj[row] = c < cc ? xv : (c == cc ? beta : 0.0);
                if (tau != 0.0) vv = c > cc ? xv * scale : (c == cc ? 1.0 : 0.0);
            }
            colp[cl] = vv;
            if (c < ((L + 7) & ~7)) vg[c] = vv;
        }
    }
    __syncthreads();
    // ---- T factor: local Gram blocks (tensor pipe), summed on CTA 0 in a fixed order
    large_gram(PB, Lc, nloc, ls.Gs, ls.scratch);   // rows beyond nloc of the slice are zero; Gs = this CTA's partial
    double* gram_all = gath;                        // [CS][kNB * 17] on CTA 0 (the gather buffers are free now)
    {
        double* remote = cluster.map_shared_rank(gram_all, 0);
        for (int e = tid; e < kNB * 17; e += kThreads) remote[r * kNB * 17 + e] = ls.Gs[e];
    }
    cluster.sync();
    if (r == 0) {
        for (int e = tid; e < kNB * 17; e += kThreads) {
            double sum = 0.0;
            for (int q2 = 0; q2 < CS; ++q2) sum += gram_all[q2 * kNB * 17 + e];
            ls.Gs[e] = sum;
        }
        __syncthreads();
        if (warp == 0) panel_t_factor(ls.Gs, sc, nbk, ls.Ts);
        __syncthreads();
        for (int idx = tid; idx < kNB * kLdr; idx += kThreads) q.Tg[idx] = ls.Ts[idx];
    }
}

// 8-row tile of a lane (rows tb + 2 t, tb + 2 t + 1 of its column) for tile-aligned row lists; `vec`: even segment
// offsets and leading dimension, i.e. one 16-byte access.
__device__ __forceinline__ void large_tile_load(const double* __restrict__ cp, const RowMap& rm, int tb, int t, bool vec, double& x0, double& x1) {
    const double* p = cp + (tb < rm.len1 ? rm.j0 : rm.a2 - rm.len1) + tb + 2 * t;
    if (vec) { const double2 v = *reinterpret_cast<const double2*>(p); x0 = v.x; x1 = v.y; }
    else { x0 = p[0]; x1 = p[1]; }
}
__device__ __forceinline__ void large_tile_store(double* __restrict__ cp, const RowMap& rm, int tb, int t, bool vec, double x0, double x1) {
    double* p = cp + (tb < rm.len1 ? rm.j0 : rm.a2 - rm.len1) + tb + 2 * t;
    if (vec) *reinterpret_cast<double2*>(p) = make_double2(x0, x1);
    else { p[0] = x0; p[1] = x1; }
}

// S2: partial Y^T = C^T V per (row chunk, column group) item.  The item's chunk of the reflectors (RC rows x kNB) is
// staged in shared memory by TMA (cp.async.bulk.tensor.2d, reflector-major, pitch = 2 mod 16) into one of two stages:
// the load of the CTA's next item is in flight while the tensor cores work on the current one.
static __device__ void large_trailing_y(const double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                 const LargeQR& q, int RC, const LargeSmem& ls) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int first = j0 + nbk, ntrail = ncols - first, L = rm.len;
    const int ncg = (ntrail + 63) >> 6, nrc = (L + RC - 1) / RC;
    const int nitems = nrc * ncg;
    const int bi = large_box_index(RC), pitch = large_box_pitch(bi);
    const unsigned bytes = (unsigned)(kNB * pitch * sizeof(double));
    const CUtensorMap* tmap = &q.tmapV[bi];
    unsigned n = *ls.nload;
    const bool vec = (((rm.j0 | (rm.a2 - rm.len1) | ld) & 1) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < nitems) {
        asm volatile("fence.proxy.async;\n" ::: "memory");  // the panel team wrote the reflectors through the generic proxy
        large_tma_load(large_stage(ls, n & 1), tmap, ((int)blockIdx.x / ncg) * RC, 0, &ls.bars[n & 1], bytes);
    }
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int rcx = it / ncg, cgx = it - rcx * ncg;
        const int c0 = rcx * RC;
        const int rows = L - c0 < RC ? L - c0 : RC;
        const int nt8 = (rows + 7) >> 3;
        const int nxt = it + (int)gridDim.x;
        if (tid == 0 && nxt < nitems)
            large_tma_load(large_stage(ls, (n + 1) & 1), tmap, (nxt / ncg) * RC, 0, &ls.bars[(n + 1) & 1], bytes);
        const double* Vt = large_stage(ls, n & 1);
        large_mbar_wait(&ls.bars[n & 1], (n >> 1) & 1);
        const int cbase = (cgx * 8 + warp) * 8;
        if (cbase < ntrail) {
            const int cidx = cbase + g;
            const bool have = cidx < ntrail;
            const double* cp = W + (size_t)(first + (have ? cidx : cbase)) * ld;
            double y[2][2][2];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) { y[e][nn][0] = 0.0; y[e][nn][1] = 0.0; }
            const double* v1 = Vt + (size_t)g * pitch + 2 * t;
            for (int i0 = 0; i0 < nt8; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    const int cl = 8 * (i0 + a) + 2 * t;
                    if (rm.aligned) {  // whole tiles of one segment: (segment base) + constant, no per-element row map
                        xa[a][0] = 0.0; xa[a][1] = 0.0;
                        if (i0 + a < nt8) large_tile_load(cp, rm, c0 + 8 * (i0 + a), t, vec, xa[a][0], xa[a][1]);
                    } else {
                        xa[a][0] = (have && cl < rows) ? cp[rm.row(c0 + cl)] : 0.0;
                        xa[a][1] = (have && cl + 1 < rows) ? cp[rm.row(c0 + cl + 1)] : 0.0;
                    }
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    if (i0 + a < nt8) {
                        const double2 lo = *reinterpret_cast<const double2*>(v1 + 8 * (i0 + a));
                        const double2 hi = *reinterpret_cast<const double2*>(v1 + 8 * (i0 + a) + (size_t)8 * pitch);
                        dmma884(y[0][0][0], y[0][0][1], xa[a][0], lo.x);
                        dmma884(y[0][1][0], y[0][1][1], xa[a][0], hi.x);
                        dmma884(y[1][0][0], y[1][0][1], xa[a][1], lo.y);
                        dmma884(y[1][1][0], y[1][1][1], xa[a][1], hi.y);
                    }
                }
            }
            if (have) {
                double* yp = q.Yp + ((size_t)rcx * q.ycols + cidx) * kNB;
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    yp[8 * nn + 2 * t] = y[0][nn][0] + y[1][nn][0];
                    yp[8 * nn + 2 * t + 1] = y[0][nn][1] + y[1][nn][1];
                }
            }
        }
        __syncthreads();  // the stage is free for the load after next
        ++n;
    }
    __syncthreads();
    if (tid == 0) *ls.nload = n;
}

// S3: C^T -= (Y^T T) V^T on the rows of each item's chunk; the reflector chunk is TMA-staged as in S2.
static __device__ void large_trailing_u(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                 const LargeQR& q, int RC, const LargeSmem& ls, double* __restrict__ Ts) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int first = j0 + nbk, ntrail = ncols - first, L = rm.len;
    const int ncg = (ntrail + 63) >> 6, nrc = (L + RC - 1) / RC;
    const int nitems = nrc * ncg;
    const int bi = large_box_index(RC), pitch = large_box_pitch(bi);
    const unsigned bytes = (unsigned)(kNB * pitch * sizeof(double));
    const CUtensorMap* tmap = &q.tmapV[bi];
    unsigned n = *ls.nload;
    const bool vec = (((rm.j0 | (rm.a2 - rm.len1) | ld) & 1) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < nitems)
        large_tma_load(large_stage(ls, n & 1), tmap, ((int)blockIdx.x / ncg) * RC, 0, &ls.bars[n & 1], bytes);
    for (int idx = tid; idx < kNB * kLdr; idx += kThreads) Ts[idx] = q.Tg[idx];
    __syncthreads();
    double tf[2][2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx)
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) tf[h][sx][nn] = Ts[(8 * h + 2 * t + sx) * kLdr + g + 8 * nn];
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int rcx = it / ncg, cgx = it - rcx * ncg;
        const int c0 = rcx * RC;
        const int rows = L - c0 < RC ? L - c0 : RC;
        const int nt8 = (rows + 7) >> 3;
        const int nxt = it + (int)gridDim.x;
        if (tid == 0 && nxt < nitems)
            large_tma_load(large_stage(ls, (n + 1) & 1), tmap, (nxt / ncg) * RC, 0, &ls.bars[(n + 1) & 1], bytes);
        const double* Vs = large_stage(ls, n & 1);
        const int cbase = (cgx * 8 + warp) * 8;
        const bool active = cbase < ntrail;
        const int cidx = cbase + g;
        const bool have = active && cidx < ntrail;
        double* cp = W + (size_t)(first + (have ? cidx : (active ? cbase : 0))) * ld;
        double z[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        if (active) {  // (sums of the partial products while the reflector chunk is in flight)
            double yt[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // Y^T[col g][reflectors 8n + 2t, 8n + 2t + 1]
            if (have) {
                for (int r = 0; r < nrc; ++r) {
                    const double* yp = q.Yp + ((size_t)r * q.ycols + cidx) * kNB;
#pragma unroll
                    for (int nn = 0; nn < 2; ++nn) { yt[nn][0] += yp[8 * nn + 2 * t]; yt[nn][1] += yp[8 * nn + 2 * t + 1]; }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int sx = 0; sx < 2; ++sx)
#pragma unroll
                    for (int nn = 0; nn < 2; ++nn) dmma884(z[nn][0], z[nn][1], yt[h][sx], tf[h][sx][nn]);
#pragma unroll
            for (int nn = 0; nn < 2; ++nn) { z[nn][0] = -z[nn][0]; z[nn][1] = -z[nn][1]; }
        }
        large_mbar_wait(&ls.bars[n & 1], (n >> 1) & 1);
        if (active) {
            for (int i0 = 0; i0 < nt8; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    const int cl = 8 * (i0 + a) + 2 * t;
                    if (rm.aligned) {
                        xa[a][0] = 0.0; xa[a][1] = 0.0;
                        if (i0 + a < nt8) large_tile_load(cp, rm, c0 + 8 * (i0 + a), t, vec, xa[a][0], xa[a][1]);
                    } else {
                        xa[a][0] = (have && cl < rows) ? cp[rm.row(c0 + cl)] : 0.0;
                        xa[a][1] = (have && cl + 1 < rows) ? cp[rm.row(c0 + cl + 1)] : 0.0;
                    }
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    if (i0 + a < nt8) {
                        const double* vb = Vs + 8 * (i0 + a) + g;
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int sx = 0; sx < 2; ++sx) dmma884(xa[a][0], xa[a][1], z[h][sx], vb[(size_t)(8 * h + 2 * t + sx) * pitch]);
                    }
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    const int cl = 8 * (i0 + a) + 2 * t;
                    if (rm.aligned) {
                        if (have && i0 + a < nt8) large_tile_store(cp, rm, c0 + 8 * (i0 + a), t, vec, xa[a][0], xa[a][1]);
                    } else {
                        if (have && cl < rows) cp[rm.row(c0 + cl)] = xa[a][0];
                        if (have && cl + 1 < rows) cp[rm.row(c0 + cl + 1)] = xa[a][1];
                    }
                }
            }
        }
        __syncthreads();  // the stage is free for the load after next
        ++n;
    }
    __syncthreads();
    if (tid == 0) *ls.nload = n;
}

// Rows per chunk of the trailing phases: about two items per CTA, chunks of 64 .. 240 rows (multiple of 8; one TMA box).
__device__ __forceinline__ int large_chunk_rows(int L, int ntrail, int cap) {
    const int ncg = (ntrail + 63) >> 6;
    int want = (2 * (int)gridDim.x + ncg - 1) / ncg;
    if (want < 1) want = 1;
    int RC = (L + want - 1) / want;
    RC = (RC + 7) & ~7;
    if (RC < 64) RC = 64;
    (void)cap;
    if (RC > 240) RC = 240;
    return RC;
}

// Grid-wide blocked QR.  On return (after a grid barrier) the upper triangle holds R.
static __device__ void householder_qr_large(cg::grid_group& grid, double* __restrict__ W, int ld, const Shape s, const LargeQR& q,
                                     const LargeSmem& ls, PhaseClock& pc) {
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    for (int j0 = 0; j0 < nref; j0 += kNB) {
        const int nbk = nref - j0 < kNB ? nref - j0 : kNB;
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        // panel on the cluster when it has at least kNB rows per CTA and the slice + exchange buffers fit
        cg::cluster_group cluster = cg::this_cluster();
        const int CS = (int)cluster.num_blocks();
        const int Lc = ((rm.len + CS - 1) / CS + 7) & ~7;
        const bool on_cluster = CS > 1 && Lc >= kNB && (size_t)kNB * Lc + 2 * CS * kGath + kGath + (size_t)CS * kNB * 17 <= (size_t)q.cap;
        if (on_cluster) {
            if (blockIdx.x < CS) large_panel_factor_cluster(cluster, W, ld, s, j0, nbk, rm, q, ls, Lc);
        } else if (blockIdx.x == 0) {
            large_panel_factor(W, ld, s, j0, nbk, rm, q, ls, pc);
        }
        pc.mark(8);
        grid.sync();
        pc.mark(9);
        const int ntrail = s.ncols - (j0 + nbk);
        if (ntrail > 0) {
            const int RC = large_chunk_rows(rm.len, ntrail, q.cap);
            large_trailing_y(W, ld, s.ncols, j0, nbk, rm, q, RC, ls);
            pc.mark(10);
            grid.sync();
            pc.mark(11);
            large_trailing_u(W, ld, s.ncols, j0, nbk, rm, q, RC, ls, ls.Ts);
            pc.mark(12);
            grid.sync();
            pc.mark(13);
        }
    }
}

}  // namespace pnmol
