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

// XOR swizzle of the row index inside reflector r: bit 1 of r -> bit 2, (bit 0 ^ bit 2) of r -> bit 3 of the row index.
// Row pairs (2 t, 2 t + 1) stay adjacent, so the pass-1 operands of a lane are one 16-byte load; both operand patterns
// of the trailing update are bank-conflict free (128-bit loads of reflectors g, g + 8 by quarter-warps; 64-bit loads of
// reflectors 2 t + sx by half-warps).
// PNMOL_FEWER_VARIANTS = 1: fewer specialisations of the hot phases are instantiated in the panel loop (see
// panel_factor_dispatch / trailing_dispatch): the kernel's instruction working set is what limits it once the CTAs of
// the grid drift out of lock-step and stop sharing instruction-cache lines.
#ifndef PNMOL_FEWER_VARIANTS
#define PNMOL_FEWER_VARIANTS 0   // measured: robust against drift (grid 288: 305 k vs 268 k) but 331 k vs 432 k in lock-step at C5
#endif
__device__ __forceinline__ int vsw(int r) { return ((r & 2) << 1) | (((r ^ (r >> 2)) & 1) << 3); }

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
    extern __shared__ __align__(16) double smem_raw[];
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
    double gq[6];  // v_j . v_i for (j, i) = 01 02 03 12 13 23 (Gram entries of the sub-panel's T factor)
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
            // ... and, in the slots of the finished columns j < i, v_j . x_i over the same rows: the Gram entries of the
            // sub-panel's T factor ride along in this round instead of a reduction round of their own after the loop
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < i) {
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int r = 0; r < R; r += 2) { a0 = fma(x[j][r], t[r], a0); a1 = fma(x[j][r + 1], t[r + 1], a1); }
                    red[j] = a0 + a1;
                }
            }
            double e[4];   // k > i: x_k at the diagonal row; j < i: v_j at the diagonal row
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k != i) e[k] = __shfl_sync(0xffffffffu, x[k][0], p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double tmp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tmp[k] = __shfl_xor_sync(0xffffffffu, red[k], o);
#pragma unroll
                for (int k = 0; k < 4; ++k) red[k] += tmp[k];
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
            // v_j . v_i = scale_i (v_j . x_i below the diagonal) + v_j[diagonal row of i]   (v_i = 0 when H = I)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < i) gq[j == 0 ? i - 1 : j == 1 ? i + 1 : 5] = tau != 0.0 ? fma(scale, red[j], e[j]) : 0.0;
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
    extern __shared__ __align__(16) double smem_raw[];
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

// T factor of the panel (dlarft), built sub-panel by sub-panel while the factor warp is busy with the next one:
//   T[4s:4s+4, 4s:4s+4] = T4_s (from the factor warp),   T[0:4s, 4s:4s+4] = -T[0:4s, 0:4s] (V_<s^T V_s) T4_s.
// Three helper warps (h = 0, 1, 2; named barrier 4) share the Gram block V_<s^T V_s on the tensor pipe -- every warp a
// third of the 8-row tiles, partial blocks summed in a fixed order -- and helper 0 does the two small products.
static __device__ __noinline__ void t_extend(unsigned buf_off, int LP, int ntile, int sidx, unsigned ts_off, unsigned t4_off,
                                      unsigned scratch_off, int h) {
    extern __shared__ __align__(16) double smem_raw[];
    const double* buf = smem_raw + buf_off;
    double* Ts = smem_raw + ts_off;
    const double* t4 = smem_raw + t4_off;
    double* scratch = smem_raw + scratch_off;  // 3 x 64 partial Gram blocks, then 64 + 64 for the products
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int c0 = 4 * sidx;
    if (h == 0 && lane < 16) Ts[(c0 + (lane >> 2)) * kLdr + c0 + (lane & 3)] = t4[lane];
    if (sidx == 0) return;
    {
        double c_lo[2] = {0.0, 0.0}, d_lo[2] = {0.0, 0.0}, c_hi[2] = {0.0, 0.0}, d_hi[2] = {0.0, 0.0};
        const int swa = vsw(g), swb = vsw(c0 + (g & 3));
        const double* lo = buf + (size_t)g * LP;
        const double* hi = lo + (size_t)8 * LP;
        const double* pb = buf + (size_t)(c0 + (g & 3)) * LP;
        const bool hi_needed = c0 > 8;
        for (int i = h; i < ntile; i += 3) {
            const int r0 = 8 * i + 2 * t;
            const double2 a = *reinterpret_cast<const double2*>(lo + (r0 ^ swa));
            double2 bv = *reinterpret_cast<const double2*>(pb + (r0 ^ swb));
            if (g >= 4) { bv.x = 0.0; bv.y = 0.0; }
            dmma884(c_lo[0], c_lo[1], a.x, bv.x);
            dmma884(d_lo[0], d_lo[1], a.y, bv.y);
            if (hi_needed) {
                const double2 a2 = *reinterpret_cast<const double2*>(hi + (r0 ^ swa));
                dmma884(c_hi[0], c_hi[1], a2.x, bv.x);
                dmma884(d_hi[0], d_hi[1], a2.y, bv.y);
            }
        }
        if (t < 2) {  // D[g][2 t + q] = (v_g . v_{c0 + 2 t + q}) over this warp's tiles
            double* mine = scratch + h * 64;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                mine[g * 4 + 2 * t + q] = c_lo[q] + d_lo[q];
                mine[(g + 8) * 4 + 2 * t + q] = c_hi[q] + d_hi[q];
            }
        }
    }
    asm volatile("bar.sync 4, 96;" ::: "memory");
    if (h != 0) return;
    double* gs = scratch + 192;
    double* ms = scratch + 256;
    for (int idx = lane; idx < c0 * 4; idx += 32) gs[idx] = (scratch[idx] + scratch[64 + idx]) + scratch[128 + idx];
    __syncwarp();
    for (int idx = lane; idx < c0 * 4; idx += 32) {  // M = T[0:c0, 0:c0] G
        const int j = idx >> 2, c = idx & 3;
        double a0 = 0.0, a1 = 0.0;
        for (int jp = j; jp < c0; jp += 2) {
            a0 = fma(Ts[j * kLdr + jp], gs[jp * 4 + c], a0);
            if (jp + 1 < c0) a1 = fma(Ts[j * kLdr + jp + 1], gs[(jp + 1) * 4 + c], a1);
        }
        ms[idx] = a0 + a1;
    }
    __syncwarp();
    for (int idx = lane; idx < c0 * 4; idx += 32) {  // T[0:c0, c0:c0+4] = -M T4
        const int j = idx >> 2, c = idx & 3;
        double acc = 0.0;
        for (int q = 0; q <= c; ++q) acc = fma(ms[j * 4 + q], t4[q * 4 + c], acc);
        Ts[j * kLdr + c0 + c] = -acc;
    }
}

// ---------------------------------------------------------------- trailing update (FP64 DMMA, compact WY)
// Columns [cbeg, cend) of the workspace, 8 per warp and round.  Fragment layout as in qr_blocked.cuh; V comes from the
// swizzled reflector-major buffer.  NT > 0: single pass with all (<= NT) tiles of a column in registers; NT = 0: two
// passes in chunks of kCh tiles (row lists of more than 16 tiles).
struct TrailOps {  // lane-dependent operand bases (even / odd tiles): index + 8 a
    const double *p1e, *p1o;                   // pass 1: reflector g (+ 8 LP: g + 8), in-tile row pair (2 t, 2 t + 1)
    const double *p2ea, *p2oa, *p2eb, *p2ob;   // pass 2: reflectors 2 t (a) and 2 t + 1 (b) (+ 8 LP: + 8), in-tile row g
    int hi;                                    // 8 LP
};

// Workspace tile of a lane: rows 8 a + 2 t, 8 a + 2 t + 1 of its column.  MODE 2: aligned row list and even offsets
// (one 16-byte access), 1: aligned row list, 0: general row list (per-element map and guards).
template <int MODE>
__device__ __forceinline__ void trail_load(const double* __restrict__ cp, const RowMap& rm, int a, int t, int nt1, int off1,
                                           int off2, double& x0, double& x1) {
    if (MODE == 2) {
        const double2 v = *reinterpret_cast<const double2*>(cp + (a < nt1 ? off1 : off2) + 8 * a);
        x0 = v.x; x1 = v.y;
    } else if (MODE == 1) {  // every tile lies in one segment of the row list: (segment base) + constant
        const double* q = cp + (a < nt1 ? off1 : off2) + 8 * a;
        x0 = q[0]; x1 = q[1];
    } else {
        tile_load<false>(cp, rm, 8 * a, t, x0, x1);
    }
}
template <int MODE>
__device__ __forceinline__ void trail_store(double* __restrict__ cp, const RowMap& rm, int a, int t, int nt1, int off1, int off2,
                                            double x0, double x1) {
    if (MODE == 2) {
        *reinterpret_cast<double2*>(cp + (a < nt1 ? off1 : off2) + 8 * a) = make_double2(x0, x1);
    } else if (MODE == 1) {
        double* q = cp + (a < nt1 ? off1 : off2) + 8 * a;
        q[0] = x0; q[1] = x1;
    } else {
        tile_store<false>(cp, rm, 8 * a, t, x0, x1);
    }
}
// tile index a = ab + ac with ab a multiple of 2 (runtime) and ac a compile-time constant: parity(a) = parity(ac)
struct P1Ops { double2 lo, hi; };
__device__ __forceinline__ P1Ops trail_p1_fetch(const TrailOps& o, int ab, int ac) {
    const double* q = ((ac & 1) ? o.p1o : o.p1e) + 8 * ab + 8 * ac;
    P1Ops r;
    r.lo = *reinterpret_cast<const double2*>(q);
    r.hi = *reinterpret_cast<const double2*>(q + o.hi);
    return r;
}
__device__ __forceinline__ void trail_p1_mma(const P1Ops& b, double x0, double x1, double (&y)[2][2][2]) {
    dmma884(y[0][0][0], y[0][0][1], x0, b.lo.x);
    dmma884(y[0][1][0], y[0][1][1], x0, b.hi.x);
    dmma884(y[1][0][0], y[1][0][1], x1, b.lo.y);
    dmma884(y[1][1][0], y[1][1][1], x1, b.hi.y);
}
struct P2Ops { double a0, b0, a1, b1; };
__device__ __forceinline__ P2Ops trail_p2_fetch(const TrailOps& o, int ab, int ac) {
    const double* qa = ((ac & 1) ? o.p2oa : o.p2ea) + 8 * ab + 8 * ac;
    const double* qb = ((ac & 1) ? o.p2ob : o.p2eb) + 8 * ab + 8 * ac;
    P2Ops r;
    r.a0 = qa[0]; r.b0 = qb[0]; r.a1 = qa[o.hi]; r.b1 = qb[o.hi];
    return r;
}
__device__ __forceinline__ void trail_p2_mma(const P2Ops& b, const double (&z)[2][2], double& x0, double& x1) {
    dmma884(x0, x1, z[0][0], b.a0);
    dmma884(x0, x1, z[0][1], b.b0);
    dmma884(x0, x1, z[1][0], b.a1);
    dmma884(x0, x1, z[1][1], b.b1);
}
// Y'^T = -(Y^T T): yt[n][q] = Y^T[col g][reflector 8 n + 2 t + q]
__device__ __forceinline__ void trail_apply_t_yt(const double* __restrict__ Ts, int g, int t, const double (&yt)[2][2],
                                                 double (&z)[2][2]);
__device__ __forceinline__ void trail_apply_t(const double* __restrict__ Ts, int g, int t, bool have, const double (&y)[2][2][2],
                                              double (&z)[2][2]) {
    double yt[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int q = 0; q < 2; ++q) yt[n][q] = have ? y[0][n][q] + y[1][n][q] : 0.0;
    trail_apply_t_yt(Ts, g, t, yt, z);
}
__device__ __forceinline__ void trail_apply_t_yt(const double* __restrict__ Ts, int g, int t, const double (&yt)[2][2],
                                                 double (&z)[2][2]) {
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

// NT > 0: single pass, NT - 4 < ntile <= NT tiles all in registers (only the last 4 carry guards).
// NT = 0: two passes in chunks of kCh tiles (full chunks unguarded, one guarded tail chunk).
template <int MODE, int NT, bool EXTRA = false>
__device__ __noinline__ void trailing_fast(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap rm, unsigned buf_off,
                                           int LP, unsigned ts_off, const QTeam tm) {
    extern __shared__ __align__(16) double smem_raw[];
    const double* buf = smem_raw + buf_off;
    const double* Ts = smem_raw + ts_off;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (rm.len + 7) >> 3;
    TrailOps o;
    {
        const int sw1 = vsw(g), swa = vsw(2 * t), swb = vsw(2 * t + 1);
        const double* r1 = buf + (size_t)g * LP;
        o.p1e = r1 + swz_even(2 * t, sw1); o.p1o = r1 + swz_odd(2 * t, sw1);
        const double* ra = buf + (size_t)(2 * t) * LP;
        const double* rb = ra + LP;
        o.p2ea = ra + swz_even(g, swa); o.p2oa = ra + swz_odd(g, swa);
        o.p2eb = rb + swz_even(g, swb); o.p2ob = rb + swz_odd(g, swb);
        o.hi = 8 * LP;
    }
    const int nt1 = (rm.len1 + 7) >> 3;                // aligned lists: tiles [0, nt1) lie in the first segment
    const int off1 = rm.j0 + 2 * t, off2 = rm.a2 - rm.len1 + 2 * t;
    // tiles [0, NG) exist for certain (with fewer variants a variant also serves shorter row lists: every tile is guarded)
    constexpr int NG = PNMOL_FEWER_VARIANTS ? 0 : (NT > 4 ? NT - 4 : 0);
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
            if (EXTRA) {  // tiles [NT, ntile): first product now (their second product follows the main block)
                for (int i0 = NT; i0 < ntile; i0 += kCh) {
                    double xe[kCh][2];
#pragma unroll
                    for (int a = 0; a < kCh; ++a) {
                        xe[a][0] = 0.0; xe[a][1] = 0.0;
                        if (i0 + a < ntile) trail_load<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xe[a][0], xe[a][1]);
                    }
#pragma unroll
                    for (int a = 0; a < kCh; ++a)
                        if (i0 + a < ntile) trail_p1_mma(trail_p1_fetch(o, i0, a), xe[a][0], xe[a][1], y);
                }
            }
            double xa[NT > 0 ? NT : 1][2];
#pragma unroll
            for (int a = 0; a < NT; ++a) {
                xa[a][0] = 0.0; xa[a][1] = 0.0;
                if (EXTRA || a < NG || a < ntile) trail_load<MODE>(cp, rm, a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
            }
            {   // pass 1, operands fetched one tile ahead
                P1Ops nxt = trail_p1_fetch(o, 0, 0);
#pragma unroll
                for (int a = 0; a < NT; ++a) {
                    const P1Ops cur = nxt;
                    if (a + 1 < NT && (EXTRA || a + 1 < NG || a + 1 < ntile)) nxt = trail_p1_fetch(o, 0, a + 1);
                    if (EXTRA || a < NG || a < ntile) trail_p1_mma(cur, xa[a][0], xa[a][1], y);
                }
            }
            trail_apply_t(Ts, g, t, have, y, z);
            {
                P2Ops nxt = trail_p2_fetch(o, 0, 0);
#pragma unroll
                for (int a = 0; a < NT; ++a) {
                    const P2Ops cur = nxt;
                    if (a + 1 < NT && (EXTRA || a + 1 < NG || a + 1 < ntile)) nxt = trail_p2_fetch(o, 0, a + 1);
                    if (EXTRA || a < NG || a < ntile) trail_p2_mma(cur, z, xa[a][0], xa[a][1]);
                }
            }
            if (have) {
#pragma unroll
                for (int a = 0; a < NT; ++a)
                    if (EXTRA || a < NG || a < ntile) trail_store<MODE>(cp, rm, a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
            }
            if (EXTRA) {
                for (int i0 = NT; i0 < ntile; i0 += kCh) {
                    double xe[kCh][2];
#pragma unroll
                    for (int a = 0; a < kCh; ++a) {
                        xe[a][0] = 0.0; xe[a][1] = 0.0;
                        if (i0 + a < ntile) trail_load<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xe[a][0], xe[a][1]);
                    }
#pragma unroll
                    for (int a = 0; a < kCh; ++a)
                        if (i0 + a < ntile) trail_p2_mma(trail_p2_fetch(o, i0, a), z, xe[a][0], xe[a][1]);
                    if (have) {
#pragma unroll
                        for (int a = 0; a < kCh; ++a)
                            if (i0 + a < ntile) trail_store<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xe[a][0], xe[a][1]);
                    }
                }
            }
        } else {
            const int nfull = ntile / kCh * kCh;
            for (int i0 = 0; i0 < nfull; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) trail_load<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                P1Ops nxt = trail_p1_fetch(o, i0, 0);
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    const P1Ops cur = nxt;
                    if (a + 1 < kCh) nxt = trail_p1_fetch(o, i0, a + 1);
                    trail_p1_mma(cur, xa[a][0], xa[a][1], y);
                }
            }
            if (nfull < ntile) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    xa[a][0] = 0.0; xa[a][1] = 0.0;
                    if (nfull + a < ntile) trail_load<MODE>(cp, rm, nfull + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a)
                    if (nfull + a < ntile) trail_p1_mma(trail_p1_fetch(o, nfull, a), xa[a][0], xa[a][1], y);
            }
            trail_apply_t(Ts, g, t, have, y, z);
            for (int i0 = 0; i0 < nfull; i0 += kCh) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) trail_load<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                P2Ops nxt = trail_p2_fetch(o, i0, 0);
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    const P2Ops cur = nxt;
                    if (a + 1 < kCh) nxt = trail_p2_fetch(o, i0, a + 1);
                    trail_p2_mma(cur, z, xa[a][0], xa[a][1]);
                }
                if (have) {
#pragma unroll
                    for (int a = 0; a < kCh; ++a) trail_store<MODE>(cp, rm, i0 + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
            }
            if (nfull < ntile) {
                double xa[kCh][2];
#pragma unroll
                for (int a = 0; a < kCh; ++a) {
                    xa[a][0] = 0.0; xa[a][1] = 0.0;
                    if (nfull + a < ntile) trail_load<MODE>(cp, rm, nfull + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
#pragma unroll
                for (int a = 0; a < kCh; ++a)
                    if (nfull + a < ntile) trail_p2_mma(trail_p2_fetch(o, nfull, a), z, xa[a][0], xa[a][1]);
                if (have) {
#pragma unroll
                    for (int a = 0; a < kCh; ++a)
                        if (nfull + a < ntile) trail_store<MODE>(cp, rm, nfull + a, t, nt1, off1, off2, xa[a][0], xa[a][1]);
                }
            }
        }
    }
}

// Priority update: the (at most 16) columns [cbeg, cend) of the NEXT panel, which the panel team is waiting for, are
// brought up to date by ALL warps of the CTA instead of one warp per column group: a column group's row tiles are split
// over kWarps / (number of groups) warps (<= kSplitTiles tiles each, all in registers); every warp forms the partial
// Y^T of its tiles, the partials are exchanged through shared memory (`xch`: the idle panel buffer) and summed in a
// fixed order, and every warp finishes its own tiles.  Must be called by every warp of the CTA (named barrier 5).
constexpr int kSplitTiles = 8;
template <int MODE>
__device__ __noinline__ void trailing_split(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap rm, unsigned buf_off,
                                            int LP, unsigned ts_off, unsigned xch_off, PhaseClock& pc) {
    extern __shared__ __align__(16) double smem_raw[];
    const double* buf = smem_raw + buf_off;
    const double* Ts = smem_raw + ts_off;
    double* xch = smem_raw + xch_off;
    pc.mark(20);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (rm.len + 7) >> 3;
    const int ngroups = (cend - cbeg + 7) >> 3;          // 1 or 2
    const int nparts = kWarps / ngroups;
    const int group = warp % ngroups, part = warp / ngroups;
    const int per = 2 * ((ntile + 2 * nparts - 1) / (2 * nparts));   // tiles per part, even (operand parity is compile-time)
    const int a0 = part * per;
    const int a1 = a0 + per < ntile ? a0 + per : ntile;  // my tiles: [a0, a1)
    TrailOps o;
    {
        const int sw1 = vsw(g), swa = vsw(2 * t), swb = vsw(2 * t + 1);
        const double* r1 = buf + (size_t)g * LP;
        o.p1e = r1 + swz_even(2 * t, sw1); o.p1o = r1 + swz_odd(2 * t, sw1);
        const double* ra = buf + (size_t)(2 * t) * LP;
        const double* rb = ra + LP;
        o.p2ea = ra + swz_even(g, swa); o.p2oa = ra + swz_odd(g, swa);
        o.p2eb = rb + swz_even(g, swb); o.p2ob = rb + swz_odd(g, swb);
        o.hi = 8 * LP;
    }
    const int nt1 = (rm.len1 + 7) >> 3;
    const int off1 = rm.j0 + 2 * t, off2 = rm.a2 - rm.len1 + 2 * t;
    const int kb = cbeg + 8 * group;
    const int col = kb + g;
    const bool have = col < cend;
    double* cp = W + (size_t)(have ? col : kb) * ld;
    double xa[kSplitTiles][2];
#pragma unroll
    for (int i = 0; i < kSplitTiles; ++i) {
        xa[i][0] = 0.0; xa[i][1] = 0.0;
        if (a0 + i < a1) trail_load<MODE>(cp, rm, a0 + i, t, nt1, off1, off2, xa[i][0], xa[i][1]);
    }
    double y[2][2][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int n = 0; n < 2; ++n) { y[e][n][0] = 0.0; y[e][n][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < kSplitTiles; ++i)
        if (a0 + i < a1) trail_p1_mma(trail_p1_fetch(o, a0, i), xa[i][0], xa[i][1], y);
    pc.mark(16);
    {
        double2* mine = reinterpret_cast<double2*>(xch + (size_t)(part * ngroups + group) * 128 + 4 * lane);
        mine[0] = have ? make_double2(y[0][0][0] + y[1][0][0], y[0][0][1] + y[1][0][1]) : make_double2(0.0, 0.0);
        mine[1] = have ? make_double2(y[0][1][0] + y[1][1][0], y[0][1][1] + y[1][1][1]) : make_double2(0.0, 0.0);
    }
    asm volatile("bar.sync 5, %0;" ::"r"(kThreads) : "memory");
    pc.mark(17);
    double yt[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int p = 0; p < nparts; ++p) {   // fixed order: bitwise reproducible
        const double2* theirs = reinterpret_cast<const double2*>(xch + (size_t)(p * ngroups + group) * 128 + 4 * lane);
        const double2 v0 = theirs[0], v1 = theirs[1];
        yt[0][0] += v0.x; yt[0][1] += v0.y; yt[1][0] += v1.x; yt[1][1] += v1.y;
    }
    if (a0 >= a1) return;
    double z[2][2];
    trail_apply_t_yt(Ts, g, t, yt, z);
#pragma unroll
    for (int i = 0; i < kSplitTiles; ++i)
        if (a0 + i < a1) trail_p2_mma(trail_p2_fetch(o, a0, i), z, xa[i][0], xa[i][1]);
    if (have) {
#pragma unroll
        for (int i = 0; i < kSplitTiles; ++i)
            if (a0 + i < a1) trail_store<MODE>(cp, rm, a0 + i, t, nt1, off1, off2, xa[i][0], xa[i][1]);
    }
    pc.mark(18);
}

__device__ __forceinline__ void trailing_split_dispatch(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap& rm,
                                                        unsigned buf_off, int LP, unsigned ts_off, unsigned xch_off, PhaseClock& pc) {
    if (rm.aligned) {
        const bool vec = (((rm.j0 | (rm.a2 - rm.len1) | ld) & 1) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
        if (vec) trailing_split<2>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, xch_off, pc);
        else trailing_split<1>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, xch_off, pc);
    } else {
        trailing_split<0>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, xch_off, pc);
    }
}

__device__ __forceinline__ void trailing_dispatch(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap& rm,
                                                  unsigned buf_off, int LP, unsigned ts_off, const QTeam tm) {
    if (cbeg >= cend) return;
    const int ntile = (rm.len + 7) >> 3;
    if (rm.aligned) {
        // 16-byte accesses when every tile address is even (segment offsets and the leading dimension)
        const bool vec = (((rm.j0 | (rm.a2 - rm.len1) | ld) & 1) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
        if (vec) {
            if (!PNMOL_FEWER_VARIANTS && ntile <= 4) trailing_fast<2, 4>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (ntile <= 8) trailing_fast<2, 8>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (!PNMOL_FEWER_VARIANTS && ntile <= 12) trailing_fast<2, 12>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (ntile <= 16) trailing_fast<2, 16>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else trailing_fast<2, 16, true>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
        } else {
            if (!PNMOL_FEWER_VARIANTS && ntile <= 4) trailing_fast<1, 4>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (ntile <= 8) trailing_fast<1, 8>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (!PNMOL_FEWER_VARIANTS && ntile <= 12) trailing_fast<1, 12>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else if (ntile <= 16) trailing_fast<1, 16>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
            else trailing_fast<1, 16, true>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
        }
    } else {
        trailing_fast<0, 0>(W, ld, cbeg, cend, rm, buf_off, LP, ts_off, tm);
    }
}

// ---------------------------------------------------------------- driver state
// The uniform state of one blocked QR lives in SHARED memory, not in registers: the phases below are separate
// (non-inlined) functions that need the whole register file, so every value the driver keeps across a call is spilled
// to local memory and re-read after it -- ncu attributed half of all long-scoreboard stalls of the step kernel to those
// reloads (local memory does not fit the small L1 that is left next to ~190 KB of shared memory per SM).  Here every
// value is read from shared memory where it is used; the only loop-carried registers are the iteration counters.
struct QRPanel {           // state of one iteration of the panel loop
    int j0, nbk, bi;       // the panel that has been factored (reflectors in buf[bi]); nbk = 0: none yet
    int more, j1, nb1;     // the panel to factor now (into buf[bi ^ 1]), if more
    RowMap rm, rm1;        // their row lists
};
struct QRCtx {
    double* W;
    int ld, nref;
    Shape s;
    FastQR fq;
    QRPanel pn[2];         // iteration it reads pn[it & 1]; one thread writes pn[(it + 1) & 1] meanwhile
};
static_assert(sizeof(QRCtx) <= 40 * sizeof(double), "QRCtx must fit its slot in the FastQR carve-out");

__device__ __forceinline__ QRCtx& qr_ctx(unsigned ctx_off) {
    extern __shared__ __align__(16) double smem_raw[];
    return *reinterpret_cast<QRCtx*>(smem_raw + ctx_off);
}
// Teams by scheduler (a warp's scheduler is its index mod 4).  PNMOL_TEAMS_BY_SCHED = 1: the panel team is the EVEN warps
// (schedulers 0 and 2: the two chains of an SM and their helpers), the update team the ODD warps (schedulers 1 and 3), so
// that the tensor-pipe trailing updates -- a DMMA holds its scheduler's FP64 pipe for 16 cycles -- never sit on a scheduler
// that runs a dependent per-column chain.  0: panel team = warps 0 .. 3, update team = warps 4 .. 7.
#ifndef PNMOL_TEAMS_BY_SCHED
#define PNMOL_TEAMS_BY_SCHED 0   // measured: 410 k (1) vs 419 k (0) member-steps/s at C5
#endif
__device__ __forceinline__ bool in_panel_team() {
    const int warp = threadIdx.x >> 5;
    return PNMOL_TEAMS_BY_SCHED ? (warp & 1) == 0 : warp < kWarps / 2;
}
__device__ __forceinline__ int update_team_index() {
    const int warp = threadIdx.x >> 5;
    return PNMOL_TEAMS_BY_SCHED ? warp >> 1 : warp - kWarps / 2;
}
// the panel team of this CTA, team index 0 = the factor warp: warp 0 in the first CTA of an SM, warp 2 in the second
__device__ __forceinline__ QTeam panel_team(const QRCtx& cx) {
    constexpr int kHalf = kWarps / 2;
    const int warp = threadIdx.x >> 5;
    if (PNMOL_TEAMS_BY_SCHED) return QTeam{((warp >> 1) - (cx.fq.slot & 1)) & (kHalf - 1), kHalf, 1};
    return QTeam{(warp - 2 * (cx.fq.slot & 1)) & (kHalf - 1), kHalf, 1};
}
// state of the iteration after `p` (one thread)
__device__ __forceinline__ void qr_advance(const QRCtx& cx, const QRPanel& p, QRPanel& q) {
    q.j0 = p.j1; q.nbk = p.nb1; q.bi = p.bi ^ 1; q.rm = p.rm1;
    const int j1 = p.j1 + p.nb1;
    q.more = j1 < cx.nref;
    q.j1 = j1;
    q.nb1 = cx.nref - j1 < kNB ? cx.nref - j1 : kNB;
    if (q.more) q.rm1 = panel_rows(cx.s, j1, j1 + q.nb1 - 1);
}

// ---------------------------------------------------------------- panel: load, factor, T
// Load the panel columns j0 .. j0 + nbk - 1 restricted to the row list (entries outside a column's own envelope and
// the rows [len, 32 R) are zero; absent columns nbk .. kNB - 1 are zero columns).
template <int R>
__device__ __forceinline__ void panel_load(const QRCtx& cx, const QRPanel& p, const QTeam tm) {
    extern __shared__ __align__(16) double smem_raw[];
    const int bi = p.bi ^ 1;
    double* buf = smem_raw + cx.fq.buf[bi];
    const unsigned ts_off = cx.fq.Ts[bi], tau_off = cx.fq.tau + bi * kNB;
    const int LP = cx.fq.LP, ld = cx.ld, j0 = p.j1, nbk = p.nb1;
    const double* __restrict__ W = cx.W;
    const RowMap rm = p.rm1;
    const Shape s = cx.s;
    const int lane = threadIdx.x & 31;
    for (int e = tm.w * 32 + lane; e < kNB * kLdr; e += tm.nw * 32) smem_raw[ts_off + e] = 0.0;  // T starts as zero
    constexpr int CMAX = 4;  // columns per warp (teams of >= 4 warps)
    double v[CMAX][R];
    // every load of the warp is issued before the first shared-memory store (one L2 round trip per panel)
#pragma unroll
    for (int q = 0; q < CMAX; ++q) {
        const int pc_ = tm.w + q * tm.nw;
        const bool valid = pc_ < nbk;
        const int jp = j0 + (valid ? pc_ : 0);
        const int et = env_top(s, jp), eb = env_bot(s, jp);
        const double* col = W + (size_t)jp * ld;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = lane + 32 * r;
            const int row = rm.row(c);
            const bool ok = valid && c < rm.len && (row < s.nt ? row <= et : row <= eb);
            v[q][r] = ok ? col[row] : 0.0;
        }
    }
#pragma unroll
    for (int q = 0; q < CMAX; ++q) {
        const int pc_ = tm.w + q * tm.nw;
        if (pc_ < kNB) {
            const int sw = vsw(pc_);
            double* dst = buf + (size_t)pc_ * LP;
#pragma unroll
            for (int r = 0; r < R; ++r) dst[(lane ^ sw) + 32 * r] = v[q][r];
            if (lane == 0) smem_raw[tau_off + pc_] = 0.0;
        }
    }
}

// Factor the panel (p.j1, p.nb1, p.rm1) into buf[p.bi ^ 1].  Everything is re-read from shared memory at its point of
// use (see QRCtx); c0 is the only register that lives across the calls.
template <int R>
__device__ __forceinline__ void panel_factor(const QRCtx& cx, const QRPanel& p, PhaseClock& pc) {
    {
        const QTeam tm = panel_team(cx);
        panel_load<R>(cx, p, tm);
        tm.sync();
    }
    pc.mark(8);
#pragma unroll 1
    for (int c0 = 0; c0 < p.nb1; c0 += 4) {
        {
            const QTeam tm = panel_team(cx);
            const int bi = p.bi ^ 1, nbk = p.nb1, sidx = c0 >> 2;
            const unsigned buf_off = cx.fq.buf[bi];
            if (tm.w == 0) {
                subpanel_factor<R>(buf_off, cx.fq.LP, c0, nbk - c0 < 4 ? nbk - c0 : 4, cx.fq.tau + bi * kNB, cx.fq.t4 + 16 * sidx,
                                   cx.W + (size_t)p.j1 * cx.ld, cx.ld, p.rm1);
            } else if (c0 > 0) {
                // hidden behind the factor warp: the previous sub-panel's reflectors applied to the columns beyond this
                // sub-panel, and the previous sub-panel's T columns
                subpanel_apply<R>(buf_off, cx.fq.LP, c0 - 4, c0 + 4, nbk, cx.fq.t4 + 16 * (sidx - 1), QTeam{tm.w - 1, tm.nw - 1, -1});
                if (tm.w <= 3)
                    t_extend(cx.fq.buf[p.bi ^ 1], cx.fq.LP, (p.rm1.len + 7) >> 3, (c0 >> 2) - 1, cx.fq.Ts[p.bi ^ 1],
                             cx.fq.t4 + 16 * ((c0 >> 2) - 1), cx.fq.scratch, (int)panel_team(cx).w - 1);
            }
        }
        panel_team(cx).sync();
        pc.mark(9);
        if (c0 + 4 < p.nb1) {  // critical: the columns of the next sub-panel only
            const int nbk = p.nb1;
            subpanel_apply<R>(cx.fq.buf[p.bi ^ 1], cx.fq.LP, c0, c0 + 4, c0 + 8 < nbk ? c0 + 8 : nbk, cx.fq.t4 + 16 * (c0 >> 2),
                              panel_team(cx));
            panel_team(cx).sync();
            pc.mark(10);
        }
    }
    {
        const int w = panel_team(cx).w;
        if (w >= 1 && w <= 3)
            t_extend(cx.fq.buf[p.bi ^ 1], cx.fq.LP, (p.rm1.len + 7) >> 3, (p.nb1 - 1) >> 2, cx.fq.Ts[p.bi ^ 1],
                     cx.fq.t4 + 16 * ((p.nb1 - 1) >> 2), cx.fq.scratch, w - 1);
    }
    pc.mark(14);
}

__device__ __forceinline__ void panel_factor_dispatch(const QRCtx& cx, const QRPanel& p, PhaseClock& pc) {
    // PNMOL_FEWER_VARIANTS: one panel body (8 rows per lane) whenever the buffers have 256 rows -- the shorter variants
    // save a few FMAs on the first panels of a factorisation but add ~90 KB of hot code to a kernel whose instruction
    // working set already exceeds the instruction caches (instruction-cache sensitivity)
    const int len = p.rm1.len;
    if (PNMOL_FEWER_VARIANTS && cx.fq.LP >= 256) { panel_factor<8>(cx, p, pc); return; }
    if (len <= 64) panel_factor<2>(cx, p, pc);
    else if (len <= 128) panel_factor<4>(cx, p, pc);
    else panel_factor<8>(cx, p, pc);
}

// ---------------------------------------------------------------- driver
// Requires every panel row list <= fq.LP <= 256 (the host checks this when it selects the CTA-per-member path).
//
// Look-ahead with warps placed by scheduler (a warp's scheduler is its index mod 4):
//   * panel team  = warps 0 .. 3.  Its factor warp -- the one that runs the dependent per-column chain -- is warp 0 in
//     the first CTA of an SM and warp 2 in the second (fq.slot), so that the chains of two co-resident CTAs never share
//     a scheduler and its FP64 pipe (measured: 218 k -> 300 k member-steps/s);
//   * update team = warps 4 .. 7: the tensor-core trailing updates.
// Iteration k: (1) ALL warps bring the columns of panel k up to date with panel k-1 (trailing_split: the row tiles of
// the 16 columns split over the 8 warps, partial products exchanged through shared memory); (2) the panel team loads
// and factors panel k into the other buffer (T factor included) while the update team applies panel k-1 to the columns
// beyond panel k; (3) one block barrier joins the teams.
static __device__ __noinline__ void householder_qr_fast(double* __restrict__ W, int ld, const Shape s, const FastQR fq, PhaseClock& pc) {
    constexpr int kHalf = kWarps / 2;
    QRCtx& cx = qr_ctx(fq.ctx);
    if (threadIdx.x == 0) {
        cx.W = W; cx.ld = ld; cx.s = s; cx.fq = fq;
        const int nrows = s.nt + s.nbot;
        const int nref = nrows < s.ncols ? nrows : s.ncols;
        cx.nref = nref;
        QRPanel& p = cx.pn[0];   // nothing factored yet; the first panel is "next"
        p.j0 = 0; p.nbk = 0; p.bi = 1; p.rm = RowMap{0, 0, 0, 0, false};
        p.more = nref > 0; p.j1 = 0; p.nb1 = nref < kNB ? nref : kNB;
        p.rm1 = panel_rows(s, 0, p.nb1 - 1);
    }
    __syncthreads();
    int it = 0;
#pragma unroll 1
    for (;; ++it) {
        const QRPanel& p = cx.pn[it & 1];
        if (!p.more) break;
        if (threadIdx.x == kThreads - 1) qr_advance(cx, p, cx.pn[(it + 1) & 1]);
        if (p.nbk > 0)   // (the other panel buffer is idle until the panel team loads panel k into it: the exchange area)
            trailing_split_dispatch(cx.W, cx.ld, p.j1, p.j1 + p.nb1, p.rm, cx.fq.buf[p.bi], cx.fq.LP, cx.fq.Ts[p.bi],
                                    cx.fq.buf[p.bi ^ 1], pc);
        if (in_panel_team()) {
            asm volatile("bar.sync 3, %0;" ::"r"(kThreads) : "memory");   // columns of the panel are up to date
            pc.mark(11);
            panel_factor_dispatch(cx, p, pc);
        } else {
            asm volatile("bar.arrive 3, %0;" ::"r"(kThreads) : "memory");
            if (p.nbk > 0)
                trailing_dispatch(cx.W, cx.ld, p.j1 + p.nb1, cx.s.ncols, p.rm, cx.fq.buf[p.bi], cx.fq.LP, cx.fq.Ts[p.bi],
                                  QTeam{update_team_index(), kHalf, 2});
        }
        __syncthreads();
        pc.mark(12);
    }
    {   // the last panel: applied to the columns beyond it by all warps
        const QRPanel& p = cx.pn[it & 1];
        if (p.nbk > 0 && p.j0 + p.nbk < cx.s.ncols)
            trailing_dispatch(cx.W, cx.ld, p.j0 + p.nbk, cx.s.ncols, p.rm, cx.fq.buf[p.bi], cx.fq.LP, cx.fq.Ts[p.bi],
                              QTeam{(int)(threadIdx.x >> 5), kWarps, 0});
    }
    __syncthreads();
    pc.mark(11);
}

}  // namespace pnmol
