// Blocked, register-resident Householder QR for one CTA (sm_100a).
//
// Same mathematics and LAPACK dlarfg conventions as householder_columns (ek1_device.cuh); what
// changes is where the data lives.  Columns are processed in panels of kNB.  For a panel the
// union of the reflector supports is a short list of rows (<= 16 G, G in {4, 8, 16, 32}); a
// column restricted to those rows is held in the registers of a G-lane group (lane s keeps
// rows s, s + G, ...: up to kRPL = 16 values), two columns per group, 32/G groups per warp.
// Dot products are reduced inside a group with log2(G) shuffle stages that serve all columns
// of the warp at once; the panel's reflectors live in shared memory (Vs) and are applied to
// every trailing column while that column sits in registers, so the trailing matrix is read
// from and written to the L2-resident workspace exactly once per panel.
//
// Panel factorisation uses one reduction round per column: the owner publishes its raw
// column x_i, every group reduces (x_i . x_k over rows > i, x_k[i]) for its own columns k --
// for k = i that is the norm -- the owner derives (beta, tau, scale) from it, and all groups
// then apply  x_k -= tau (x_k[i] + scale x_i.x_k) v  with v = scale x_i below the diagonal.
#pragma once
// included from ek1_device.cuh (after householder_columns)

namespace pnmol {

constexpr int kNB = 16;   // panel width
#ifndef PNMOL_WY_TRAILING
#define PNMOL_WY_TRAILING 0
#endif
// compact-WY (blocks of 4) trailing update; measured slower than the reflector-at-a-time loop at D = 150 on B200
// (8 rows per lane make the 32-lane butterflies as expensive as the FMAs): profiles/r01_notes.md
constexpr bool kUseWyTrailing = PNMOL_WY_TRAILING != 0;
#ifndef PNMOL_DMMA_TRAILING
#define PNMOL_DMMA_TRAILING 1
#endif
constexpr bool kUseDmmaTrailing = PNMOL_DMMA_TRAILING != 0;
#ifndef PNMOL_DMMA_PAIR
#define PNMOL_DMMA_PAIR 0
#endif
constexpr bool kUseDmmaPair = PNMOL_DMMA_PAIR != 0;  // warp-pair single-pass variant (measured slower: L2 access efficiency)
#ifndef PNMOL_RPL
#define PNMOL_RPL 8
#endif
constexpr int kRPL = PNMOL_RPL;  // rows per lane (upper bound; chunks of 4 beyond the row list are skipped)
constexpr int kQ = kRPL / 4;

struct RowMap {  // compact row list of a panel: c < len1 -> j0 + c, else a2 + (c - len1)
    int j0, len1, a2, len;
    bool aligned;  // every 8-row tile [8k, 8k+8) lies in one segment (or the segments are contiguous) and len % 8 == 0
    __host__ __device__ __forceinline__ int row(int c) const { return c + (c < len1 ? j0 : a2 - len1); }
};

__device__ __forceinline__ int env_top(const Shape& s, int j) {
    int e = s.te ? s.te[j] : s.nt - 1;
    return e > s.nt - 1 ? s.nt - 1 : e;
}
__device__ __forceinline__ int env_bot(const Shape& s, int j) {
    const int last = s.nt + s.nbot - 1;
    int e = s.be ? s.be[j] : last;
    return e > last ? last : e;
}

// Row list of the panel j0 .. jl.  With s.ldr > 0 the list is aligned to the 8-row tiles of the tensor-core trailing
// update: the top segment is extended by real rows below its envelope (the reflectors are zero there, the trailing
// update rewrites them unchanged) to a multiple of 8 rows -- or, when that reaches the bottom block, to the bottom
// block itself, which makes the two segments one contiguous range -- and the end of the list likewise while the rows
// exist in the column.  rm.aligned then says that every tile lies in one segment (or the segments are contiguous).
__host__ __device__ __forceinline__ RowMap make_row_map(int nt, int ldr, int j0, int e1, int e2) {
    // j0 < nt: e1 = last top row of the panel's union, e2 = env_bot(last column); else e2 = last row of the union
    RowMap rm;
    rm.j0 = j0;
    int last;            // last workspace row of the list
    bool single_top = false;
    if (j0 < nt) {
        int len1 = e1 - j0 + 1;
        const int len2 = e2 >= nt ? e2 - nt + 1 : 0;
        if (ldr > 0 && len2 > 0) {
            const int pad1 = (-len1) & 7;
            if (j0 + len1 + pad1 >= nt) len1 = nt - j0;  // reaches the bottom block: one contiguous range
            else len1 += pad1;
        }
        rm.len1 = len1;
        rm.a2 = nt;
        rm.len = len1 + len2;
        single_top = len2 == 0;
        last = len2 > 0 ? nt + len2 - 1 : j0 + len1 - 1;
    } else {
        rm.len1 = 0;
        rm.a2 = j0;
        rm.len = e2 - j0 + 1;
        last = e2;
    }
    rm.aligned = false;
    if (ldr > 0) {
        const int pad = (-rm.len) & 7;
        if (last + pad < ldr) {
            rm.len += pad;
            if (single_top) rm.len1 += pad;
        }
        rm.aligned = (rm.len & 7) == 0 && ((rm.len1 & 7) == 0 || rm.a2 - rm.len1 == rm.j0 || rm.len1 == rm.len);
    }
    return rm;
}

__device__ __forceinline__ RowMap panel_rows(const Shape& s, int j0, int jl) {
    if (j0 < s.nt) {
        const int jt = jl < s.nt ? jl : s.nt - 1;
        int e1 = env_top(s, jt);
        if (e1 < jt) e1 = jt;
        return make_row_map(s.nt, s.ldr, j0, e1, env_bot(s, jl));
    }
    int e2 = env_bot(s, jl);
    if (e2 < jl) e2 = jl;
    return make_row_map(s.nt, s.ldr, j0, 0, e2);
}

// Interleaved butterfly reductions inside a G-lane group.
template <int G>
__device__ __forceinline__ void group_sum2(double& a, double& b) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o);
        const double tb = __shfl_xor_sync(0xffffffffu, b, o);
        a += ta;
        b += tb;
    }
}
template <int G>
__device__ __forceinline__ void group_sum3(double& a, double& b, double& c) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o);
        const double tb = __shfl_xor_sync(0xffffffffu, b, o);
        const double tc = __shfl_xor_sync(0xffffffffu, c, o);
        a += ta; b += tb; c += tc;
    }
}
template <int G>
__device__ __forceinline__ void group_sum4(double& a, double& b, double& c, double& d) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o);
        const double tb = __shfl_xor_sync(0xffffffffu, b, o);
        const double tc = __shfl_xor_sync(0xffffffffu, c, o);
        const double td = __shfl_xor_sync(0xffffffffu, d, o);
        a += ta; b += tb; c += tc; d += td;
    }
}

// Apply one reflector to the two register-resident columns of a lane group.
template <int G>
__device__ __forceinline__ void apply_reflector(const double (&v)[kRPL], double (&x0)[kRPL], double (&x1)[kRPL], int nq,
                                                double tau) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
        if (q < nq) {
#pragma unroll
            for (int rr = 0; rr < 4; rr += 2) {
                const int r = 4 * q + rr;
                a0 = fma(v[r], x0[r], a0);
                a1 = fma(v[r], x1[r], a1);
                b0 = fma(v[r + 1], x0[r + 1], b0);
                b1 = fma(v[r + 1], x1[r + 1], b1);
            }
        }
    }
    double d0 = a0 + b0, d1 = a1 + b1;
    group_sum2<G>(d0, d1);
    const double w0 = -tau * d0, w1 = -tau * d1;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
        if (q < nq) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = 4 * q + rr;
                x0[r] = fma(w0, v[r], x0[r]);
                x1[r] = fma(w1, v[r], x1[r]);
            }
        }
    }
}

// Gram entries v_k . v_i for the pairs inside each block of 4 reflectors (6 per block): gsm[6 * b + {01,02,03,12,13,23}].
__device__ __forceinline__ void panel_gram(const double* __restrict__ Vs, int vld, int len, int nbk, double* __restrict__ gsm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pair = warp; pair < 24; pair += kWarps) {
        const int b = pair / 6, q = pair - 6 * b;
        const int k = 4 * b + (q < 3 ? 0 : q < 5 ? 1 : 2);
        const int i = 4 * b + (q < 3 ? q + 1 : q < 5 ? q - 1 : 3);
        double acc = 0.0;
        if (i < nbk) {
            for (int c = lane; c < len; c += 32) acc = fma(Vs[k * vld + c], Vs[i * vld + c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) gsm[pair] = acc;
    }
}

// Trailing update in compact-WY form, 4 reflectors at a time.  GT lanes hold a column (8 rows per lane), two columns
// per lane group.  Per block: 8 independent dot products (4 reflectors x 2 columns) reduced together, the 4 x 4
// triangular recurrence  w_i = -tau_i (v_i.x + sum_{k<i} w_k v_k.v_i)  on scalars, then one fused update.
template <int GT>
__device__ __noinline__ void trailing_wy(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                         const double* __restrict__ Vs, int vld, const double* __restrict__ sc,
                                         const double* __restrict__ gsm) {
    constexpr int CPW = (32 / GT) * 2;
    constexpr int R = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / GT, sl = lane % GT;
    const double* vbase = Vs + sl;
    for (int kb = j0 + nbk + warp * CPW; kb < ncols; kb += kWarps * CPW) {
        const int k0 = kb + g * 2, k1 = k0 + 1;
        const bool h0 = k0 < ncols, h1 = k1 < ncols;
        double* c0 = W + (size_t)(h0 ? k0 : kb) * ld;
        double* c1 = W + (size_t)(h1 ? k1 : kb) * ld;
        double x0[R], x1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = sl + GT * r;
            const int row = rm.row(c);
            const bool in = c < rm.len;
            x0[r] = (in && h0) ? c0[row] : 0.0;
            x1[r] = (in && h1) ? c1[row] : 0.0;
        }
        for (int b = 0; 4 * b < nbk; ++b) {
            const int i0 = 4 * b;
            double v[4][R];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool have = i0 + q < nbk;
                const double* vp = vbase + (i0 + q) * vld;
#pragma unroll
                for (int r = 0; r < R; ++r) v[q][r] = have ? vp[GT * r] : 0.0;
            }
            double y[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) { y[q][0] = 0.0; y[q][1] = 0.0; }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    y[q][0] = fma(v[q][r], x0[r], y[q][0]);
                    y[q][1] = fma(v[q][r], x1[r], y[q][1]);
                }
            }
#pragma unroll
            for (int o = GT / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double ta = __shfl_xor_sync(0xffffffffu, y[q][0], o);
                    const double tb = __shfl_xor_sync(0xffffffffu, y[q][1], o);
                    y[q][0] += ta;
                    y[q][1] += tb;
                }
            }
            const double t0 = sc[3 * i0], t1 = i0 + 1 < nbk ? sc[3 * (i0 + 1)] : 0.0;
            const double t2 = i0 + 2 < nbk ? sc[3 * (i0 + 2)] : 0.0, t3 = i0 + 3 < nbk ? sc[3 * (i0 + 3)] : 0.0;
            const double g01 = gsm[6 * b], g02 = gsm[6 * b + 1], g03 = gsm[6 * b + 2];
            const double g12 = gsm[6 * b + 3], g13 = gsm[6 * b + 4], g23 = gsm[6 * b + 5];
            double w[4][2];
#pragma unroll
            for (int cw = 0; cw < 2; ++cw) {
                w[0][cw] = -t0 * y[0][cw];
                w[1][cw] = -t1 * fma(w[0][cw], g01, y[1][cw]);
                w[2][cw] = -t2 * fma(w[1][cw], g12, fma(w[0][cw], g02, y[2][cw]));
                w[3][cw] = -t3 * fma(w[2][cw], g23, fma(w[1][cw], g13, fma(w[0][cw], g03, y[3][cw])));
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    x0[r] = fma(w[q][0], v[q][r], x0[r]);
                    x1[r] = fma(w[q][1], v[q][r], x1[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = sl + GT * r;
            const int row = rm.row(c);
            const bool in = c < rm.len;
            if (in && h0) c0[row] = x0[r];
            if (in && h1) c1[row] = x1[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core trailing update (FP64 DMMA, mma.sync.m8n8k4.f64).  Fragment layout (verified on B200,
// tools/dmma_probe.cu): lane = 4 g + t supplies A[g][t] and B[t][g] and owns D[g][2t], D[g][2t+1].
//
// Compact WY:  C <- C - V T^T (V^T C)  for the kNB reflectors of a panel.  A warp takes 8 trailing columns; lane
// (g, t) handles column g and, of every 8-row tile i of the panel's row list, rows 8i + 2t and 8i + 2t + 1.  With
// that assignment the column data is at once the A operand of  Y^T = C^T V  (k-step (i, e) uses rows 8i + 2t + e)
// and the accumulator of  C^T -= Y'^T V^T  -- no shuffles, no layout conversion; Y'^T = Y^T T is 8 more DMMAs.
// Shared memory: Vs[refl][ldt] (ldt = 2 mod 8) feeds the B operand of the second product, Vr[row][kLdr] that of the
// first, Ts[kNB][kLdr] holds T; the strides make every fragment load bank-conflict free.
constexpr int kLdr = 18;
#ifndef PNMOL_KCH
#define PNMOL_KCH 8
#endif
constexpr int kCh = PNMOL_KCH;  // 8-row tiles fetched per chunk

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Gs[k][i] = v_k . v_i for k <= i (upper triangle of V^T V) on the tensor pipe: every warp takes a slice of the 8-row
// tiles and accumulates the three 8 x 8 blocks (0,0), (0,1), (1,1); the per-warp partial blocks go to `scratch`
// (kWarps x 192 doubles) and are summed in a fixed order (bitwise reproducible; no floating-point atomics).
// Operands: A[g][t] = V[row][g + 8a], B[t][g] = V[row][g + 8b], row = 8i + 2t + e.  Contains one __syncthreads.
__device__ __forceinline__ void panel_gram_dmma(const double* __restrict__ Vr, int len, double* __restrict__ Gs,
                                                double* __restrict__ scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (len + 7) >> 3;
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    double d00[2] = {0.0, 0.0}, d01[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
    for (int i = warp; i < ntile; i += kWarps) {
        const double* v0 = Vr + (8 * i + 2 * t) * kLdr + g;
        const double a0 = v0[0], a1 = v0[8], b0 = v0[kLdr], b1 = v0[kLdr + 8];
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
}

// T of the compact WY representation (upper triangular, LAPACK dlarft forward/columnwise):
// T[i][i] = tau_i,  T[0:i, i] = -tau_i T[0:i, 0:i] (V^T v_i)[0:i].  One warp, lane k owns row k in registers.
__device__ __forceinline__ void panel_t_factor(const double* __restrict__ Gs, const double* __restrict__ sc, int nbk,
                                               double* __restrict__ Ts) {
    const int k = threadIdx.x & 31;
    double Trow[kNB];
#pragma unroll
    for (int j = 0; j < kNB; ++j) Trow[j] = 0.0;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
        if (i < nbk) {  // (columns beyond the panel stay zero; their Gram entries were never written)
            const double tau = sc[3 * i];
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < i; j += 2) {
                a0 = fma(Trow[j], Gs[j * 17 + i], a0);
                if (j + 1 < i) a1 = fma(Trow[j + 1], Gs[(j + 1) * 17 + i], a1);
            }
            Trow[i] = k < i ? -tau * (a0 + a1) : (k == i ? tau : 0.0);
        }
    }
    if (k < kNB) {
#pragma unroll
        for (int j = 0; j < kNB; ++j) Ts[k * kLdr + j] = Trow[j];
    }
}

// Tile fetch/store of the tensor-core trailing update.  An 8-row tile that lies entirely inside one segment of the
// compact row list (the common case) is two unpredicated accesses at (segment base) + constant; only the tile that
// straddles the segments and the last, partial tile take the per-element path.  `cp` points at the lane's column
// (a valid dummy column for lanes beyond the matrix, which never store).
template <bool ALIGNED>
__device__ __forceinline__ void tile_load(const double* __restrict__ cp, const RowMap& rm, int tb, int t, double& x0, double& x1) {
    if (ALIGNED) {  // every tile is (its segment's base) + constant
        const double* p = cp + (tb < rm.len1 ? rm.j0 : rm.a2 - rm.len1) + tb + 2 * t;
        x0 = p[0]; x1 = p[1];
    } else {
        const int c0 = tb + 2 * t, c1 = c0 + 1;
        x0 = c0 < rm.len ? cp[rm.row(c0)] : 0.0;
        x1 = c1 < rm.len ? cp[rm.row(c1)] : 0.0;
    }
}
template <bool ALIGNED>
__device__ __forceinline__ void tile_store(double* __restrict__ cp, const RowMap& rm, int tb, int t, double x0, double x1) {
    if (ALIGNED) {
        double* p = cp + (tb < rm.len1 ? rm.j0 : rm.a2 - rm.len1) + tb + 2 * t;
        p[0] = x0; p[1] = x1;
    } else {
        const int c0 = tb + 2 * t, c1 = c0 + 1;
        if (c0 < rm.len) cp[rm.row(c0)] = x0;
        if (c1 < rm.len) cp[rm.row(c1)] = x1;
    }
}

// One chunk of kCh tiles of the two passes (FULL: all kCh tiles exist, no per-tile guards).
template <bool ALIGNED, bool FULL>
__device__ __forceinline__ void trailing_pass1_chunk(const double* __restrict__ cp, const RowMap& rm, int i0, int ntile, int g, int t,
                                                     const double* __restrict__ Vr, double (&y)[2][2][2]) {
    double xa[kCh][2];
#pragma unroll
    for (int a = 0; a < kCh; ++a) {
        xa[a][0] = 0.0; xa[a][1] = 0.0;
        if (FULL || i0 + a < ntile) tile_load<ALIGNED>(cp, rm, 8 * (i0 + a), t, xa[a][0], xa[a][1]);
    }
#pragma unroll
    for (int a = 0; a < kCh; ++a) {
        if (FULL || i0 + a < ntile) {
            const double* v0 = Vr + (8 * (i0 + a) + 2 * t) * kLdr + g;
            dmma884(y[0][0][0], y[0][0][1], xa[a][0], v0[0]);
            dmma884(y[0][1][0], y[0][1][1], xa[a][0], v0[8]);
            dmma884(y[1][0][0], y[1][0][1], xa[a][1], v0[kLdr]);
            dmma884(y[1][1][0], y[1][1][1], xa[a][1], v0[kLdr + 8]);
        }
    }
}
template <bool ALIGNED, bool FULL>
__device__ __forceinline__ void trailing_pass2_chunk(double* __restrict__ cp, const RowMap& rm, int i0, int ntile, int g, int t, bool have,
                                                     const double* __restrict__ Vs, int ldt, const double (&z)[2][2]) {
    double xa[kCh][2];
#pragma unroll
    for (int a = 0; a < kCh; ++a) {
        xa[a][0] = 0.0; xa[a][1] = 0.0;
        if (FULL || i0 + a < ntile) tile_load<ALIGNED>(cp, rm, 8 * (i0 + a), t, xa[a][0], xa[a][1]);
    }
    const double* vt = Vs + (2 * t) * ldt + g;
#pragma unroll
    for (int a = 0; a < kCh; ++a) {
        if (FULL || i0 + a < ntile) {
            const double* vb = vt + 8 * (i0 + a);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int sx = 0; sx < 2; ++sx) dmma884(xa[a][0], xa[a][1], z[h][sx], vb[(8 * h + sx) * ldt]);
        }
    }
    if (have) {
#pragma unroll
        for (int a = 0; a < kCh; ++a)
            if (FULL || i0 + a < ntile) tile_store<ALIGNED>(cp, rm, 8 * (i0 + a), t, xa[a][0], xa[a][1]);
    }
}

template <bool ALIGNED>
__device__ __noinline__ void trailing_dmma(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                           const double* __restrict__ Vs, int ldt, const double* __restrict__ Vr,
                                           const double* __restrict__ Ts) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (rm.len + 7) >> 3;
    // B fragments of T: T[8h + 2t + s][g + 8n]
    double tf[2][2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx)
#pragma unroll
            for (int n = 0; n < 2; ++n) tf[h][sx][n] = Ts[(8 * h + 2 * t + sx) * kLdr + g + 8 * n];
    for (int kb = j0 + nbk + warp * 8; kb < ncols; kb += kWarps * 8) {
        const int col = kb + g;
        const bool have = col < ncols;
        double* cp = W + (size_t)(have ? col : kb) * ld;
        // ---- Y^T = C^T V : accumulators [row parity e][reflector half n]; rows are fetched in chunks of kCh tiles
        double y[2][2][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int n = 0; n < 2; ++n) { y[e][n][0] = 0.0; y[e][n][1] = 0.0; }
        int i0 = 0;
        for (; i0 + kCh <= ntile; i0 += kCh) trailing_pass1_chunk<ALIGNED, true>(cp, rm, i0, ntile, g, t, Vr, y);
        if (i0 < ntile) trailing_pass1_chunk<ALIGNED, false>(cp, rm, i0, ntile, g, t, Vr, y);
        double yt[2][2];  // Y^T[col g][reflectors 8n + 2t, 8n + 2t + 1]
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int q = 0; q < 2; ++q) yt[n][q] = have ? y[0][n][q] + y[1][n][q] : 0.0;
        // ---- Y'^T = Y^T T
        double z[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int sx = 0; sx < 2; ++sx)
#pragma unroll
                for (int n = 0; n < 2; ++n) dmma884(z[n][0], z[n][1], yt[h][sx], tf[h][sx][n]);
#pragma unroll
        for (int n = 0; n < 2; ++n) { z[n][0] = -z[n][0]; z[n][1] = -z[n][1]; }
        // ---- C^T -= Y'^T V^T, again in chunks of kCh tiles
        for (i0 = 0; i0 + kCh <= ntile; i0 += kCh) trailing_pass2_chunk<ALIGNED, true>(cp, rm, i0, ntile, g, t, have, Vs, ldt, z);
        if (i0 < ntile) trailing_pass2_chunk<ALIGNED, false>(cp, rm, i0, ntile, g, t, have, Vs, ldt, z);
    }
}

// Single-pass variant for row lists of at most 8 * 2 * kHalf rows: a warp PAIR shares an 8-column group, each warp
// keeps its half of the row tiles in registers (all loads issued up front: one L2 round trip), the two partial
// Y^T are exchanged through `xch` (2 slots x kWarps x 128 doubles, double buffered: one __syncthreads per round),
// both warps apply T redundantly and update their own rows in place.  Every warp runs the same number of rounds.
constexpr int kHalf = 16;  // tiles per warp (row lists up to 256)

static __device__ __noinline__ void trailing_dmma_pair(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                                const double* __restrict__ Vs, int ldt, const double* __restrict__ Vr,
                                                const double* __restrict__ Ts, double* __restrict__ xch, PhaseClock& pc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (rm.len + 7) >> 3;
    const int nth = (ntile + 1) >> 1;          // tiles of the first half
    const int half = warp & 1, pair = warp >> 1;
    const int tbeg = half ? nth : 0, tcnt = half ? ntile - nth : nth;
    constexpr int kPairs = kWarps / 2;
    const int first = j0 + nbk;
    const int rounds = (ncols - first + 8 * kPairs - 1) / (8 * kPairs);
    for (int rd = 0; rd < rounds; ++rd) {
        const int col = first + (rd * kPairs + pair) * 8 + g;
        const bool have = col < ncols;
        double* cp = W + (size_t)(have ? col : first) * ld;
        double xa[kHalf][2];
#pragma unroll
        for (int a = 0; a < kHalf; ++a) {
            const int c0 = 8 * (tbeg + a) + 2 * t, c1 = c0 + 1;
            const bool live = a < tcnt;
            xa[a][0] = (have && live && c0 < rm.len) ? cp[rm.row(c0)] : 0.0;
            xa[a][1] = (have && live && c1 < rm.len) ? cp[rm.row(c1)] : 0.0;
        }
        double y[2][2][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int n = 0; n < 2; ++n) { y[e][n][0] = 0.0; y[e][n][1] = 0.0; }
#pragma unroll
        for (int a = 0; a < kHalf; ++a) {
            if (a < tcnt) {
                const double* v0 = Vr + (8 * (tbeg + a) + 2 * t) * kLdr + g;
                dmma884(y[0][0][0], y[0][0][1], xa[a][0], v0[0]);
                dmma884(y[0][1][0], y[0][1][1], xa[a][0], v0[8]);
                dmma884(y[1][0][0], y[1][0][1], xa[a][1], v0[kLdr]);
                dmma884(y[1][1][0], y[1][1][1], xa[a][1], v0[kLdr + 8]);
            }
        }
        pc.mark(16);
        // exchange the partial Y^T with the partner warp (fixed summation order: first half + second half)
        double* slot = xch + (rd & 1) * (kWarps * 128);
        double* mine = slot + warp * 128 + lane;
        mine[0] = y[0][0][0] + y[1][0][0];
        mine[32] = y[0][0][1] + y[1][0][1];
        mine[64] = y[0][1][0] + y[1][1][0];
        mine[96] = y[0][1][1] + y[1][1][1];
        __syncthreads();
        pc.mark(17);
        const double* lo = slot + (warp & ~1) * 128 + lane;
        const double* hi = lo + 128;
        double yt[2][2];
        yt[0][0] = lo[0] + hi[0];
        yt[0][1] = lo[32] + hi[32];
        yt[1][0] = lo[64] + hi[64];
        yt[1][1] = lo[96] + hi[96];
        double z[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int sx = 0; sx < 2; ++sx) {
                const double* tp = Ts + (8 * h + 2 * t + sx) * kLdr + g;
                dmma884(z[0][0], z[0][1], yt[h][sx], tp[0]);
                dmma884(z[1][0], z[1][1], yt[h][sx], tp[8]);
            }
#pragma unroll
        for (int n = 0; n < 2; ++n) { z[n][0] = -z[n][0]; z[n][1] = -z[n][1]; }
#pragma unroll
        for (int a = 0; a < kHalf; ++a) {
            if (a < tcnt) {
                const double* vb = Vs + 8 * (tbeg + a) + g;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) dmma884(xa[a][0], xa[a][1], z[h][sx], vb[(8 * h + 2 * t + sx) * ldt]);
            }
        }
        pc.mark(18);
#pragma unroll
        for (int a = 0; a < kHalf; ++a) {
            const int c0 = 8 * (tbeg + a) + 2 * t, c1 = c0 + 1;
            const bool live = a < tcnt;
            if (have && live && c0 < rm.len) cp[rm.row(c0)] = xa[a][0];
            if (have && live && c1 < rm.len) cp[rm.row(c1)] = xa[a][1];
        }
        pc.mark(19);
    }
}

// One panel: factor columns j0 .. j0+nbk-1 and apply their reflectors to all later columns.
// Shared memory: Vs[kNB][vld] reflectors, xraw[2][vld] raw column of the current reflector (double
// buffered), sc[3 * i] = tau of reflector i.
template <int G>
__device__ __noinline__ void qr_panel_step(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk, const RowMap rm,
                                           double* __restrict__ Vs, int vld, double* __restrict__ xraw,
                                           double* __restrict__ sc, double* __restrict__ Vr, double* __restrict__ Ts,
                                           double* __restrict__ Gs, double* __restrict__ scratch, PhaseClock& pc) {
    double* gsm = sc + 48;  // 24 Gram entries of the compact-WY blocks
    const int ldt = vld + 2;  // reflector-major stride of Vs (2 mod 8: conflict-free DMMA fragment loads)
    constexpr int CPW = (32 / G) * 2;  // trailing columns per warp (two per lane group)
    constexpr int PPW = 32 / G;        // panel columns per warp (one per lane group)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, sl = lane % G;
    const int nt = s.nt;
    // row chunks (of 4 per lane) that hold data; the driver picks the smallest G with kRPL * G >= len, so for G >= 8 more
    // than half of the kRPL rows per lane are in use and all kQ chunks are live (compile time: no guards in the loops)
    const int nq = (G >= 8 && kRPL == 8) ? kQ : (rm.len + 4 * G - 1) / (4 * G);
    double x0[kRPL], x1[kRPL];

    // ---- load this group's panel column (entries outside the column's own envelope are zero)
    // round-robin ownership (column p -> warp p % kWarps, group p / kWarps): consecutive reflectors are owned by
    // different warps, so the owner's post-work overlaps with the next owner's critical chain
    // For G = 32 (one group per warp) a group holds a SECOND panel column in x1 (column p0 + kWarps), so that panels
    // stay kNB = 16 columns wide for row lists of 129 .. 256 rows as well.
    constexpr bool kTwo = G == 32;
    // Only as many warps as the kNB columns need take part (8 for G >= 16, 4 for G = 8, 2 for G = 4): the others just
    // keep the barriers, which removes their share of the (redundant) per-column instructions.
    constexpr int NWU = kNB / PPW < kWarps ? (kNB / PPW > 0 ? kNB / PPW : 1) : kWarps;
    const int p0 = warp + NWU * g;
    const bool hp = p0 < nbk && g < PPW && warp < NWU;
    const int jp = j0 + (hp ? p0 : 0);
    const int etp = hp ? env_top(s, jp) : -1, ebp = hp ? env_bot(s, jp) : -1;
    const int p1 = p0 + NWU * PPW;
    const bool hq = kTwo && p1 < nbk && g < PPW;
    const int jq = j0 + (hq ? p1 : 0);
    const int etq = hq ? env_top(s, jq) : -1, ebq = hq ? env_bot(s, jq) : -1;
    {
        const double* c0 = W + (size_t)jp * ld;
        const double* c1 = W + (size_t)jq * ld;
#pragma unroll
        for (int r = 0; r < kRPL; ++r) {
            const int c = sl + G * r;
            const int row = rm.row(c);
            const bool ok = c < rm.len && (row < nt ? row <= etp : row <= ebp);
            x0[r] = ok ? c0[row] : 0.0;
            if (kTwo) {
                const bool ok1 = hq && c < rm.len && (row < nt ? row <= etq : row <= ebq);
                x1[r] = ok1 ? c1[row] : 0.0;
            }
        }
    }
    pc.mark(8);

    // ---- factor the panel: one barrier and one reduction round per column.  The owner publishes its column with
    // the rows c <= i zeroed (so nobody else needs per-element masks) and the diagonal entry alpha; every group
    // that still holds a live column reduces x_i . x_k and derives the reflector scalars (dlarfg) redundantly.
    int wlast = -1;  // last panel column held by this warp
#pragma unroll
    for (int gg = 0; gg < PPW * (kTwo ? 2 : 1); ++gg)
        if (warp < NWU && warp + NWU * gg < nbk) wlast = warp + NWU * gg;
    for (int i = 0; i < nbk; ++i) {
        const bool own1 = hq && p1 == i;  // (the pivot column is this group's second column)
        const bool own = (hp && p0 == i) || own1;
        const int ri = G >= kNB ? 0 : i / G, si = G >= kNB ? i : i % G;  // register slot and lane that hold row i (slot 0 for G >= 16: compile time)
        double* xr = xraw + (i & 1) * vld;  // double buffered: the next owner may publish while others still read
        if (own) {
#pragma unroll
            for (int r = 0; r < kRPL; ++r) {
                const bool gt = r > ri || (r == ri && sl > si);
                const double xo = (kTwo && own1) ? x1[r] : x0[r];
                xr[sl + G * r] = gt ? xo : 0.0;
                if (r == ri && sl == si) sc[3 * i + 1] = xo;
            }
        }
        __syncthreads();
        if (wlast < i) continue;  // no live panel column in this warp: only keep the barrier
        const double al = sc[3 * i + 1];
        double d0 = 0.0, ss = 0.0, d0b = 0.0, ssb = 0.0, d1 = 0.0, d1b = 0.0;
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
            if (q < nq) {
#pragma unroll
                for (int rr = 0; rr < 4; rr += 2) {
                    const int r = 4 * q + rr;
                    const double t = xr[sl + G * r], u = xr[sl + G * (r + 1)];
                    d0 = fma(t, x0[r], d0);
                    ss = fma(t, t, ss);
                    d0b = fma(u, x0[r + 1], d0b);
                    ssb = fma(u, u, ssb);
                    if (kTwo) { d1 = fma(t, x1[r], d1); d1b = fma(u, x1[r + 1], d1b); }
                }
            }
        }
        d0 += d0b; ss += ssb; d1 += d1b;
        // entry of this group's column at row i: slot ri of lane si
        double e0 = x0[0];
        if (G < 16 && ri == 1) e0 = x0[1];
        if (G < 8 && ri == 2) e0 = x0[2];
        if (G < 8 && ri == 3) e0 = x0[3];
        e0 = __shfl_sync(0xffffffffu, e0, (lane & ~(G - 1)) | si);
        double e1 = 0.0;
        if (kTwo) {
            e1 = __shfl_sync(0xffffffffu, x1[0], si);  // (G = 32: row i is slot 0 of lane i)
            group_sum3<G>(d0, ss, d1);
        } else {
            group_sum2<G>(d0, ss);
        }
        // dlarfg on (alpha, ||x||^2): beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta)
        double tau = 0.0, beta = al, scale = 0.0;
        if (ss != 0.0) {  // zero sub-column -> H = I
            const double s2 = fma(al, al, ss);
            const double rn = rsqrt(s2);
            const double nrm = s2 * rn;
            beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg (IEEE copysign, -0.0 counts as negative)
            tau = (beta - al) * -copysign(rn, al);
            scale = __drcp_rn(al - beta);
        }
        // (tau == 0: H = I; f, g and scale are then zero, the update below adds zeros and v is stored as zero)
        const double f0 = p0 > i && hp ? -tau * fma(scale, d0, e0) : 0.0;
        const double g0 = f0 * scale;
        const double f1 = (kTwo && p1 > i && hq) ? -tau * fma(scale, d1, e1) : 0.0;
        const double g1 = f1 * scale;
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
            if (q < nq) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int r = 4 * q + rr;
                    const double t = xr[sl + G * r];
                    x0[r] = fma(g0, t, x0[r]);
                    if (kTwo) x1[r] = fma(g1, t, x1[r]);
                }
            }
        }
        if (kTwo && sl == si) x1[0] += f1;  // (ri = 0 for G = 32)
        if (sl == si) {  // v = 1 at row i
            x0[0] += ri == 0 ? f0 : 0.0;
            if (G < 16) x0[1] += ri == 1 ? f0 : 0.0;
            if (G < 8) { x0[2] += ri == 2 ? f0 : 0.0; x0[3] += ri == 3 ? f0 : 0.0; }
        }
        if (own) {  // reflector into Vs; the column itself becomes (R entries, beta, zeros)
            if (sl == 0) sc[3 * i] = tau;
            const double one = tau != 0.0 ? 1.0 : 0.0;
#pragma unroll
            for (int r = 0; r < kRPL; ++r) {
                const bool eq = r == ri && sl == si;
                const bool ge = r > ri || (r == ri && sl >= si);
                const double vv = eq ? one : scale * xr[sl + G * r];
                Vs[i * ldt + sl + G * r] = vv;
                Vr[(sl + G * r) * kLdr + i] = vv;
                if (ge) {
                    if (kTwo && own1) x1[r] = eq ? beta : 0.0;
                    else x0[r] = eq ? beta : 0.0;
                }
            }
        }
    }
    pc.mark(9);

    // ---- write the factored panel column back
    if (hp) {
        double* c0 = W + (size_t)jp * ld;
#pragma unroll
        for (int r = 0; r < kRPL; ++r) {
            const int c = sl + G * r;
            const int row = rm.row(c);
            if (c < rm.len && (row < nt ? row <= etp : row <= ebp)) c0[row] = x0[r];
        }
    }
    if (kTwo && hq) {
        double* c1 = W + (size_t)jq * ld;
#pragma unroll
        for (int r = 0; r < kRPL; ++r) {
            const int c = sl + G * r;
            const int row = rm.row(c);
            if (c < rm.len && (row < nt ? row <= etq : row <= ebq)) c1[row] = x1[r];
        }
    }
    __syncthreads();  // all of Vs / sc written
    pc.mark(10);

    // ---- trailing columns: tensor-core compact-WY update when the row list fits the V buffers, else (or on request)
    // the DFMA paths
    if (kUseDmmaTrailing && scratch != nullptr && 8 * ((rm.len + 7) >> 3) <= vld) {
        if (j0 + nbk < s.ncols) {
            panel_gram_dmma(Vr, rm.len, Gs, scratch);
            __syncthreads();
            pc.mark(13);
            if (warp == 0) panel_t_factor(Gs, sc, nbk, Ts);
            __syncthreads();
            pc.mark(14);
            if (kUseDmmaPair && rm.len <= 16 * kHalf) trailing_dmma_pair(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, Vr, Ts, scratch, pc);
            else if (rm.aligned) trailing_dmma<true>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, Vr, Ts);
            else trailing_dmma<false>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, Vr, Ts);
        }
    } else
    if (kUseWyTrailing && rm.len <= 256) {
        panel_gram(Vs, ldt, rm.len, nbk, gsm);
        __syncthreads();
        if (rm.len <= 32) trailing_wy<4>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, sc, gsm);
        else if (rm.len <= 64) trailing_wy<8>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, sc, gsm);
        else if (rm.len <= 128) trailing_wy<16>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, sc, gsm);
        else trailing_wy<32>(W, ld, s.ncols, j0, nbk, rm, Vs, ldt, sc, gsm);
    } else {
        const double* vbase = Vs + sl;
        for (int kb = j0 + nbk + warp * CPW; kb < s.ncols; kb += kWarps * CPW) {
            const int k0 = kb + g * 2, k1 = k0 + 1;
            const bool h0 = k0 < s.ncols, h1 = k1 < s.ncols;
            double* c0 = W + (size_t)(h0 ? k0 : kb) * ld;
            double* c1 = W + (size_t)(h1 ? k1 : kb) * ld;
#pragma unroll
            for (int r = 0; r < kRPL; ++r) {
                const int c = sl + G * r;
                const int row = rm.row(c);
                const bool in = c < rm.len;
                x0[r] = (in && h0) ? c0[row] : 0.0;
                x1[r] = (in && h1) ? c1[row] : 0.0;
            }
            for (int i = 0; i < nbk; ++i) {
                const double tau = sc[3 * i];
                if (tau == 0.0) continue;
                double v[kRPL];
                const double* vp = vbase + i * ldt;
#pragma unroll
                for (int q = 0; q < kQ; ++q) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) v[4 * q + rr] = q < nq ? vp[G * (4 * q + rr)] : 0.0;
                }
                apply_reflector<G>(v, x0, x1, nq, tau);
            }
#pragma unroll
            for (int r = 0; r < kRPL; ++r) {
                const int c = sl + G * r;
                const int row = rm.row(c);
                const bool in = c < rm.len;
                if (in && h0) c0[row] = x0[r];
                if (in && h1) c1[row] = x1[r];
            }
        }
    }
    pc.mark(11);
    __syncthreads();  // Vs / sc are rewritten by the next panel; workspace writes are visible
    pc.mark(12);
}

// Blocked QR driver.  Panels whose row list exceeds 512 rows (or the smem V buffer) fall
// back to the unblocked column-by-column routine.
static __device__ void householder_qr_blocked(double* __restrict__ W, int ld, const Shape s, double* Vs, int vld, double* xraw,
                                       double* sc, double* Vr, double* Ts, double* Gs, double* scratch, double* vbuf,
                                       double* red, PhaseClock& pc) {
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    int j0 = 0;
    while (j0 < nref) {
#ifndef PNMOL_PANEL_CAP
#define PNMOL_PANEL_CAP kNB
#endif
        int nbk = nref - j0 < PNMOL_PANEL_CAP ? nref - j0 : PNMOL_PANEL_CAP;  // (tuning: narrower panels)
        RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        // lanes per column: the smallest group that holds the row list in kRPL rows per lane
        const int G = rm.len <= 4 * kRPL ? 4 : rm.len <= 8 * kRPL ? 8 : rm.len <= 16 * kRPL ? 16 : 32;
        const int cap = kWarps * (32 / G) * (G == 32 ? 2 : 1);  // panel columns the CTA can hold in registers (one per lane group, two for G = 32)
        if (nbk > cap) {
            nbk = cap;
            rm = panel_rows(s, j0, j0 + nbk - 1);
        }
        if (rm.len > 32 * kRPL || kRPL * G > vld) {
            householder_columns(W, ld, s, j0, j0 + nbk, vbuf, red);
        } else if (G == 4) {
            qr_panel_step<4>(W, ld, s, j0, nbk, rm, Vs, vld, xraw, sc, Vr, Ts, Gs, scratch, pc);
        } else if (G == 8) {
            qr_panel_step<8>(W, ld, s, j0, nbk, rm, Vs, vld, xraw, sc, Vr, Ts, Gs, scratch, pc);
        } else if (G == 16) {
            qr_panel_step<16>(W, ld, s, j0, nbk, rm, Vs, vld, xraw, sc, Vr, Ts, Gs, scratch, pc);
        } else {
            qr_panel_step<32>(W, ld, s, j0, nbk, rm, Vs, vld, xraw, sc, Vr, Ts, Gs, scratch, pc);
        }
        j0 += nbk;
    }
}

}  // namespace pnmol
