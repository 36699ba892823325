// Blocked Householder QR for one CTA, second generation (sm_100a).
//
// Same mathematics, LAPACK dlarfg/dlarft conventions, envelopes and compact row lists as qr_blocked.cuh; what changes
// is who does what, so that the dependent per-column chain costs one warp's issue slots instead of the whole CTA's:
//
//   * The panel (kNB columns restricted to the panel's row list) is staged in shared memory in a reflector-major,
//     XOR-swizzled buffer (element (r, c) at r LP + (c ^ vsw(r))).  The swizzle makes BOTH tensor-core operand
//     patterns of the trailing update bank-conflict free on the same copy (rows along k for Y^T = C^T V, reflectors
//     along k for C^T -= Y'^T V^T), so a panel needs one 16 x LP buffer instead of two layouts, and the factored
//     panel becomes the V operand in place.
//   * A panel is factored in sub-panels of 4 columns.  ONE warp holds a sub-panel in registers (lane l keeps rows
//     l, l + 32, ...) and factors its 4 columns warp-synchronously -- one fused reduction round per column (norm and
//     the dot products with the later columns in one butterfly), no block barrier -- then forms the sub-panel's 4 x 4
//     T factor.  The team applies the block reflector to the remaining panel columns in shared memory (one barrier).
//   * Trailing update: compact WY on the FP64 tensor pipe as before, but a column group whose row list fits 16 tiles is
//     processed in ONE pass with all its tiles in registers (every load issued up front, each workspace element read
//     and written once per panel).
//   * Look-ahead: while the update team applies panel k to the far trailing columns, the panel team applies it to the
//     columns of panel k+1, loads them and factors panel k+1 into the other buffer.  Teams synchronise with named
//     barriers; one block barrier per panel.
#pragma once
// included from ek1_device.cuh (after qr_blocked.cuh: RowMap, panel_rows, tile_load/tile_store, dmma884)

namespace pnmol {

// XOR swizzle of the row index inside reflector r (bits 0..2 of r -> bits 0, 3, 2 of the row index)
__device__ __forceinline__ int vsw(int r) { return (r & 1) | ((r & 2) << 2) | (r & 4); }

// A team: warps [first, first + nw) of the CTA; named barrier `bar` (0 = the whole CTA).
struct QTeam {
    int w, nw, bar;
    __device__ __forceinline__ void sync() const {
        if (bar == 0) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nw * 32) : "memory");
    }
};

template <int NV>
__device__ __forceinline__ void warp_sum_n(double (&v)[NV]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double tmp[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) tmp[k] = __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] += tmp[k];
    }
}

// ---------------------------------------------------------------- sub-panel factorisation (one warp)
// Columns c0 .. c0 + nc - 1 (nc <= 4) of the panel in `buf`; list position p is the diagonal of panel column p.
// Writes the reflectors (unit diagonal, zeros above) back to buf, tau[c0 + i], the 4 x 4 T factor to t4 and the
// finished R entries (list positions <= p of column p) to the workspace column Wp + p ld.
template <int R>
__device__ __noinline__ void subpanel_factor(unsigned buf_off, int LP, int c0, int nc, unsigned tau_off, unsigned t4_off,
                                             double* __restrict__ Wp, int ld, const RowMap rm) {
    extern __shared__ double smem_raw[];
    double* buf = smem_raw + buf_off;
    double* tau_s = smem_raw + tau_off;
    double* t4 = smem_raw + t4_off;
    const int lane = threadIdx.x & 31;
    double x[4][R];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sw = vsw(c0 + q);
        const double* col = buf + (size_t)(c0 + q) * LP;
#pragma unroll
        for (int r = 0; r < R; ++r) x[q][r] = col[(lane ^ sw) + 32 * r];  // (absent columns were loaded as zero columns)
    }
    double tauv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        {   // (an absent column, i >= nc, is a zero column: tau = 0, H = I, nothing is written to the workspace)
            const int p = c0 + i;  // diagonal: lane p, slot 0
            double t[R];
#pragma unroll
            for (int r = 0; r < R; ++r) t[r] = (r > 0 || lane > p) ? x[i][r] : 0.0;
            const double al = __shfl_sync(0xffffffffu, x[i][0], p);
            // one reduction round: ||x below the diagonal||^2 (slot i) and x_i . x_k for the later columns k (slot k)
            double red[4];
            {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; r += 2) { a0 = fma(t[r], t[r], a0); a1 = fma(t[r + 1], t[r + 1], a1); }
                red[i] = a0 + a1;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k > i) {
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int r = 0; r < R; r += 2) { a0 = fma(t[r], x[k][r], a0); a1 = fma(t[r + 1], x[k][r + 1], a1); }
                    red[k] = a0 + a1;
                }
            }
            double e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k > i) e[k] = __shfl_sync(0xffffffffu, x[k][0], p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double tmp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k >= i) tmp[k] = __shfl_xor_sync(0xffffffffu, red[k], o);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k >= i) red[k] += tmp[k];
            }
            const double ss = red[i];
            // dlarfg on (alpha, ||x||^2): beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta)
            double tau = 0.0, beta = al, scale = 0.0;
            if (ss != 0.0) {  // zero sub-column -> H = I
                const double s2 = fma(al, al, ss);
                const double rn = rsqrt(s2);
                const double nrm = s2 * rn;
                beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg
                tau = (beta - al) * -copysign(rn, al);
                scale = __drcp_rn(al - beta);
            }
            tauv[i] = tau;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k > i) {
                    const double f = -tau * fma(scale, red[k], e[k]);
                    const double g = f * scale;
#pragma unroll
                    for (int r = 0; r < R; ++r) x[k][r] = fma(g, t[r], x[k][r]);
                    if (lane == p) x[k][0] += f;
                }
            }
            // finished R entries of this column: list positions < p (slot 0 of lanes < p) and beta on the diagonal
            if (lane <= p && i < nc) Wp[(size_t)p * ld + rm.row(lane)] = lane < p ? x[i][0] : beta;
            // the column becomes the reflector: zeros above, one on the diagonal (zero when H = I), scale * x below
#pragma unroll
            for (int r = 0; r < R; ++r) x[i][r] = scale * t[r];
            if (lane == p) x[i][0] = tau != 0.0 ? 1.0 : 0.0;
        }
    }
    // T factor of the sub-panel (dlarft, forward / columnwise): T[i][i] = tau_i, T[0:k, k] = -tau_k T[0:k, 0:k] (V^T v_k)[0:k]
    double gq[6];  // v_i . v_k for (i, k) = 01 02 03 12 13 23
    {
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = i + 1; k < 4; ++k) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; r += 2) { a0 = fma(x[i][r], x[k][r], a0); a1 = fma(x[i][r + 1], x[k][r + 1], a1); }
                gq[idx++] = a0 + a1;
            }
    }
    warp_sum_n<6>(gq);
    const double t00 = tauv[0], t11 = tauv[1], t22 = tauv[2], t33 = tauv[3];
    const double t01 = -t11 * (t00 * gq[0]);
    const double t02 = -t22 * fma(t01, gq[3], t00 * gq[1]);
    const double t12 = -t22 * (t11 * gq[3]);
    const double t03 = -t33 * fma(t02, gq[5], fma(t01, gq[4], t00 * gq[2]));
    const double t13 = -t33 * fma(t12, gq[5], t11 * gq[4]);
    const double t23 = -t33 * (t22 * gq[5]);
    if (lane == 0) {
        t4[0] = t00; t4[1] = t01; t4[2] = t02; t4[3] = t03;
        t4[4] = 0.0; t4[5] = t11; t4[6] = t12; t4[7] = t13;
        t4[8] = 0.0; t4[9] = 0.0; t4[10] = t22; t4[11] = t23;
        t4[12] = 0.0; t4[13] = 0.0; t4[14] = 0.0; t4[15] = t33;
    }
    if (lane < 4) tau_s[c0 + lane] = lane == 0 ? t00 : lane == 1 ? t11 : lane == 2 ? t22 : t33;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sw = vsw(c0 + q);
        double* col = buf + (size_t)(c0 + q) * LP;
#pragma unroll
        for (int r = 0; r < R; ++r) col[(lane ^ sw) + 32 * r] = x[q][r];
    }
}

// Apply the block reflector of sub-panel c0 (4 reflectors, T factor t4) to the panel columns [cbeg, cend), one
// column per warp at a time:  x <- x - V T^T (V^T x).
template <int R>
__device__ __noinline__ void subpanel_apply(unsigned buf_off, int LP, int c0, int cbeg, int cend, unsigned t4_off, const QTeam tm) {
    extern __shared__ double smem_raw[];
    double* buf = smem_raw + buf_off;
    const double* t4 = smem_raw + t4_off;
    const int lane = threadIdx.x & 31;
    if (cbeg + tm.w >= cend) return;
    double v[4][R];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sw = vsw(c0 + q);
        const double* col = buf + (size_t)(c0 + q) * LP;
#pragma unroll
        for (int r = 0; r < R; ++r) v[q][r] = col[(lane ^ sw) + 32 * r];
    }
    const double t00 = t4[0], t01 = t4[1], t02 = t4[2], t03 = t4[3], t11 = t4[5], t12 = t4[6], t13 = t4[7], t22 = t4[10],
                 t23 = t4[11], t33 = t4[15];
    for (int c = cbeg + tm.w; c < cend; c += tm.nw) {
        const int sw = vsw(c);
        double* col = buf + (size_t)c * LP;
        double x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = col[(lane ^ sw) + 32 * r];
        double y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int r = 0; r < R; r += 2) { a0 = fma(v[q][r], x[r], a0); a1 = fma(v[q][r + 1], x[r + 1], a1); }
            y[q] = a0 + a1;
        }
        warp_sum_n<4>(y);
        // w = T^T y (T upper triangular)
        const double w0 = -(t00 * y[0]);
        const double w1 = -fma(t01, y[0], t11 * y[1]);
        const double w2 = -fma(t02, y[0], fma(t12, y[1], t22 * y[2]));
        const double w3 = -fma(t03, y[0], fma(t13, y[1], fma(t23, y[2], t33 * y[3])));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double acc = fma(w0, v[0][r], fma(w1, v[1][r], x[r]));
            col[(lane ^ sw) + 32 * r] = fma(w2, v[2][r], fma(w3, v[3][r], acc));
        }
    }
}

// ---------------------------------------------------------------- T factor of a whole panel
// Swizzled row index of tile a (8 rows) for a lane whose in-tile row is x (0..7) and whose reflector has swizzle sw:
// (8 a + x) ^ sw = 8 a + ((x ^ (sw & 7)) +- (sw & 8)), + for even a, - for odd a.  swz_even / swz_odd return the
// lane-dependent part, so that every operand address of an unrolled tile loop is (lane register) + constant.
__device__ __forceinline__ int swz_even(int x, int sw) { return (x ^ (sw & 7)) + (sw & 8); }
__device__ __forceinline__ int swz_odd(int x, int sw) { return (x ^ (sw & 7)) - (sw & 8); }

// Gram matrix V^T V (upper triangle) on the tensor pipe, at most 4 warps of the team; fixed summation order.
__device__ __noinline__ void gram16(unsigned buf_off, int LP, int ntile, unsigned gs_off, unsigned scratch_off, const QTeam tm) {
    extern __shared__ double smem_raw[];
    const double* buf = smem_raw + buf_off;
    double* Gs = smem_raw + gs_off;
    double* scratch = smem_raw + scratch_off;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int nwg = tm.nw < 4 ? tm.nw : 4;
    if (tm.w < nwg) {
        double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
        double d00[2] = {0.0, 0.0}, d01[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
        const int sw = vsw(g);
        const double* lo = buf + (size_t)g * LP;
        const double* hi = lo + (size_t)8 * LP;
        for (int i = tm.w; i < ntile; i += nwg) {
            const int r0 = (8 * i + 2 * t) ^ sw, r1 = (8 * i + 2 * t + 1) ^ sw;
            const double a0 = lo[r0], a1 = hi[r0], b0 = lo[r1], b1 = hi[r1];
            dmma884(c00[0], c00[1], a0, a0);
            dmma884(c01[0], c01[1], a0, a1);
            dmma884(c11[0], c11[1], a1, a1);
            dmma884(d00[0], d00[1], b0, b0);
            dmma884(d01[0], d01[1], b0, b1);
            dmma884(d11[0], d11[1], b1, b1);
        }
        double* mine = scratch + tm.w * 192;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = g * 8 + 2 * t + q;
            mine[e] = c00[q] + d00[q];
            mine[64 + e] = c01[q] + d01[q];
            mine[128 + e] = c11[q] + d11[q];
        }
    }
    tm.sync();
    for (int e = tm.w * 32 + lane; e < 192; e += tm.nw * 32) {
        double sum = 0.0;
        for (int w = 0; w < nwg; ++w) sum += scratch[w * 192 + e];
        const int blk = e >> 6, r = (e & 63) >> 3, c = e & 7;
        Gs[(r + (blk == 2 ? 8 : 0)) * 17 + c + (blk >= 1 ? 8 : 0)] = sum;
    }
}

// dlarft on one warp (lane k owns row k of T); tau with unit stride.
__device__ __noinline__ void t_factor16(unsigned gs_off, unsigned tau_off, unsigned ts_off) {
    extern __shared__ double smem_raw[];
    const double* Gs = smem_raw + gs_off;
    const double* tau_s = smem_raw + tau_off;
    double* Ts = smem_raw + ts_off;
    const int k = threadIdx.x & 31;
    double Trow[kNB];
#pragma unroll
    for (int j = 0; j < kNB; ++j) Trow[j] = 0.0;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
        const double tau = tau_s[i];
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < i; j += 2) {
            a0 = fma(Trow[j], Gs[j * 17 + i], a0);
            if (j + 1 < i) a1 = fma(Trow[j + 1], Gs[(j + 1) * 17 + i], a1);
        }
        Trow[i] = k < i ? -tau * (a0 + a1) : (k == i ? tau : 0.0);
    }
    if (k < kNB) {
#pragma unroll
        for (int j = 0; j < kNB; ++j) Ts[k * kLdr + j] = Trow[j];
    }
}

// ---------------------------------------------------------------- trailing update (FP64 DMMA, compact WY)
// Columns [cbeg, cend) of the workspace, 8 per warp and round.  Fragment layout as in qr_blocked.cuh; V comes from the
// swizzled reflector-major buffer.  NT > 0: single pass with all (<= NT) tiles of a column in registers; NT = 0: two
// passes in chunks of kCh tiles (row lists of more than 16 tiles).
struct TrailOps {  // lane-dependent operand bases (even / odd tiles): index + 8 a
    const double *p1e0, *p1o0, *p1e1, *p1o1;   // pass 1: reflector g (+ 8 LP: g + 8), in-tile rows 2 t, 2 t + 1
    const double *p2ea, *p2oa, *p2eb, *p2ob;   // pass 2: reflectors 2 t (a) and 2 t + 1 (b) (+ 8 LP: + 8), in-tile row g
    int hi;                                    // 8 LP
};

template <bool ALIGNED>
__device__ __forceinline__ void trail_load(const double* __restrict__ cp, const RowMap& rm, int a, int t, int nt1, int off1,
                                           int off2, double& x0, double& x1) {
    if (ALIGNED) {  // every tile lies in one segment of the row list: (segment base) + constant
        const double* q = cp + (a < nt1 ? off1 : off2) + 8 * a;
        x0 = q[0]; x1 = q[1];
    } else {
        tile_load<false>(cp, rm, 8 * a, t, x0, x1);
    }
}
template <bool ALIGNED>
__device__ __forceinline__ void trail_store(double* __restrict__ cp, const RowMap& rm, int a, int t, int nt1, int off1, int off2,
                                            double x0, double x1) {
    if (ALIGNED) {
        double* q = cp + (a < nt1 ? off1 : off2) + 8 * a;
        q[0] = x0; q[1] = x1;
    } else {
        tile_store<false>(cp, rm, 8 * a, t, x0, x1);
    }
}
// tile index a = ab + ac with ab a multiple of 2 (runtime) and ac a compile-time constant: parity(a) = parity(ac)
__device__ __forceinline__ void trail_pass1(const TrailOps& o, int ab, int ac, double x0, double x1, double (&y)[2][2][2]) {
    const double* q0 = ((ac & 1) ? o.p1o0 : o.p1e0) + 8 * ab + 8 * ac;
    const double* q1 = ((ac & 1) ? o.p1o1 : o.p1e1) + 8 * ab + 8 * ac;
    dmma884(y[0][0][0], y[0][0][1], x0, q0[0]);
    dmma884(y[0][1][0], y[0][1][1], x0, q0[o.hi]);
    dmma884(y[1][0][0], y[1][0][1], x1, q1[0]);
    dmma884(y[1][1][0], y[1][1][1], x1, q1[o.hi]);
}
__device__ __forceinline__ void trail_pass2(const TrailOps& o, int ab, int ac, const double (&z)[2][2], double& x0, double& x1) {
    const double* qa = ((ac & 1) ? o.p2oa : o.p2ea) + 8 * ab + 8 * ac;
    const double* qb = ((ac & 1) ? o.p2ob : o.p2eb) + 8 * ab + 8 * ac;
    dmma884(x0, x1, z[0][0], qa[0]);
    dmma884(x0, x1, z[0][1], qb[0]);
    dmma884(x0, x1, z[1][0], qa[o.hi]);
    dmma884(x0, x1, z[1][1], qb[o.hi]);
}
// Y'^T = -(Y^T T): yt[n][q] = Y^T[col g][reflector 8 n + 2 t + q]
__device__ __forceinline__ void trail_apply_t(const double* __restrict__ Ts, int g, int t, bool have, const double (&y)[2][2][2],
                                              double (&z)[2][2]) {
    double yt[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int q = 0; q < 2; ++q) yt[n][q] = have ? y[0][n][q] + y[1][n][q] : 0.0;
#pragma unroll
    for (int n = 0; n < 2; ++n) { z[n][0] = 0.0; z[n][1] = 0.0; }
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
}

template <bool ALIGNED, int NT>
__device__ __noinline__ void trailing_fast(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap rm, unsigned buf_off,
                                           int LP, unsigned ts_off, const QTeam tm) {
    extern __shared__ double smem_raw[];
    const double* buf = smem_raw + buf_off;
    const double* Ts = smem_raw + ts_off;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (rm.len + 7) >> 3;
    TrailOps o;
    {
        const int sw1 = vsw(g), swa = vsw(2 * t), swb = vsw(2 * t + 1);
        const double* r1 = buf + (size_t)g * LP;
        o.p1e0 = r1 + swz_even(2 * t, sw1);     o.p1o0 = r1 + swz_odd(2 * t, sw1);
        o.p1e1 = r1 + swz_even(2 * t + 1, sw1); o.p1o1 = r1 + swz_odd(2 * t + 1, sw1);
        const double* ra = buf + (size_t)(2 * t) * LP;
        const double* rb = ra + LP;
        o.p2ea = ra + swz_even(g, swa); o.p2oa = ra + swz_odd(g, swa);
        o.p2eb = rb + swz_even(g, swb); o.p2ob = rb + swz_odd(g, swb);
        o.hi = 8 * LP;
    }
    const int nt1 = (rm.len1 + 7) >> 3;                // aligned lists: tiles [0, nt1) lie in the first segment
    const int off1 = rm.j0 + 2 * t, off2 = rm.a2 - rm.len1 + 2 * t;
    for (int kb = cbeg + tm.w * 8; kb < cend; kb += tm.nw * 8) {
        const int col = kb + g;
        const bool have = col < cend;
        double* cp = W + (size_t)(have ? col : kb) * ld;
        double y[2][2][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int n = 0; n < 2; ++n) { y[e][n][0] = 0.0; y[e][n][1] = 0.0; }
        double z[2][2];
        if (NT > 0) {
            double xa[NT > 0 ? NT : 1][2];
#pragma unroll
            for (int a = 0; a < NT; ++a) {
                xa[a][0] = 0.0; xa[a][1] = 0.0;
                if (a < ntile) trail_load<ALIGNED>(cp, rm, a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
            }
#pragma unroll
            for (int a = 0; a < NT; ++a)
                if (a < ntile) trail_pass1(o, 0, a, xa[a][0], xa[a][1], y);
            trail_apply_t(Ts, g, t, have, y, z);
#pragma unroll
            for (int a = 0; a < NT; ++a)
                if (a < ntile) trail_pass2(o, 0, a, z, xa[a][0], xa[a][1]);
            if (have) {
#pragma unroll
                for (int a = 0; a < NT; ++a)
                    if (a < ntile) trail_store<ALIGNED>(cp, rm, a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
            }
        } else {
            for (int i0 = 0; i0 < ntile; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    xa[a][0] = 0.0; xa[a][1] = 0.0;
                    if (i0 + a < ntile) trail_load<ALIGNED>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a)
                    if (i0 + a < ntile) trail_pass1(o, i0, a, xa[a][0], xa[a][1], y);
            }
            trail_apply_t(Ts, g, t, have, y, z);
            for (int i0 = 0; i0 < ntile; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    xa[a][0] = 0.0; xa[a][1] = 0.0;
                    if (i0 + a < ntile) trail_load<ALIGNED>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a)
                    if (i0 + a < ntile) trail_pass2(o, i0, a, z, xa[a][0], xa[a][1]);
                if (have) {
#pragma unroll
                    for (int a = 0; a < kCh; ++a)
                        if (i0 + a < ntile) trail_store<ALIGNED>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
            }
        }
    }
}

__device__ __forceinline__ void trailing_dispatch(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap& rm,
                                                  unsigned buf_off, int LP, unsigned ts_off, const QTeam tm) {
    if (cbeg >= cend) return;
    const int ntile = (rm.len + 7) >> 3;
    if (rm.aligned) {
        if (ntile <= 8) trailing_fast<true, 8>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
        else if (ntile <= 16) trailing_fast<true, 16>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
        else trailing_fast<true, 0>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
    } else {
        trailing_fast<false, 0>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
    }
}

// ---------------------------------------------------------------- panel: load, factor, T
// Load the panel columns j0 .. j0 + nbk - 1 restricted to the row list (entries outside a column's own envelope and
// the rows [len, 32 R) are zero; absent columns nbk .. kNB - 1 are zero columns).
template <int R>
__device__ __forceinline__ void panel_load(const double* __restrict__ W, int ld, const Shape& s, int j0, int nbk,
                                           const RowMap& rm, unsigned buf_off, int LP, unsigned tau_off, const QTeam tm) {
    extern __shared__ double smem_raw[];
    double* buf = smem_raw + buf_off;
    const int lane = threadIdx.x & 31;
    for (int p = tm.w; p < kNB; p += tm.nw) {
        const bool valid = p < nbk;
        const int jp = j0 + (valid ? p : 0);
        const int et = env_top(s, jp), eb = env_bot(s, jp);
        const double* col = W + (size_t)jp * ld;
        const int sw = vsw(p);
        double* dst = buf + (size_t)p * LP;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = lane + 32 * r;
            const int row = rm.row(c);
            const bool ok = valid && c < rm.len && (row < s.nt ? row <= et : row <= eb);
            dst[(lane ^ sw) + 32 * r] = ok ? col[row] : 0.0;
        }
        if (lane == 0) smem_raw[tau_off + p] = 0.0;
    }
}

template <int R>
__device__ __forceinline__ void panel_factor(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk, const RowMap& rm,
                                             const FastQR& fq, int bi, const QTeam tm) {
    const unsigned buf_off = fq.buf[bi], tau_off = fq.tau + bi * kNB;
    const int LP = fq.LP;
    panel_load<R>(W, ld, s, j0, nbk, rm, buf_off, LP, tau_off, tm);
    tm.sync();
#pragma unroll 1
    for (int c0 = 0; c0 < nbk; c0 += 4) {
        const int nc = nbk - c0 < 4 ? nbk - c0 : 4;
        if (tm.w == 0) subpanel_factor<R>(buf_off, LP, c0, nc, tau_off, fq.t4, W + (size_t)j0 * ld, ld, rm);
        tm.sync();
        if (c0 + 4 < nbk) {
            subpanel_apply<R>(buf_off, LP, c0, c0 + 4, nbk, fq.t4, tm);
            tm.sync();
        }
    }
}

__device__ __forceinline__ void panel_factor_dispatch(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk,
                                                      const RowMap& rm, const FastQR& fq, int bi, bool need_t, const QTeam tm) {
    if (rm.len <= 64) panel_factor<2>(W, ld, s, j0, nbk, rm, fq, bi, tm);
    else if (rm.len <= 128) panel_factor<4>(W, ld, s, j0, nbk, rm, fq, bi, tm);
    else panel_factor<8>(W, ld, s, j0, nbk, rm, fq, bi, tm);
    if (need_t) {  // T factor of the whole panel (the last sub-panel's barrier precedes this)
        gram16(fq.buf[bi], fq.LP, (rm.len + 7) >> 3, fq.Gs, fq.scratch, tm);
        tm.sync();
        if (tm.w == 0) t_factor16(fq.Gs, fq.tau + bi * kNB, fq.Ts[bi]);
    }
}

// ---------------------------------------------------------------- driver
// Requires every panel row list <= fq.LP <= 256 (the host checks this when it selects the CTA-per-member path).
__device__ __noinline__ void householder_qr_fast(double* __restrict__ W, int ld, const Shape s, const FastQR fq, PhaseClock& pc) {
    const int warp = threadIdx.x >> 5;
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    const QTeam all{warp, kWarps, 0};
    int bi = 0;
    for (int j0 = 0; j0 < nref; j0 += kNB) {
        const int nbk = nref - j0 < kNB ? nref - j0 : kNB;
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        const bool trail = j0 + nbk < s.ncols;
        panel_factor_dispatch(W, ld, s, j0, nbk, rm, fq, bi, trail, all);
        __syncthreads();
        pc.mark(9);
        if (trail) trailing_dispatch(W, ld, j0 + nbk, s.ncols, rm, fq.buf[bi], fq.LP, fq.Ts[bi], all);
        __syncthreads();
        pc.mark(11);
        bi ^= 1;
    }
}

}  // namespace pnmol
