// Warp-per-member Householder QR (sm_100a): one warp owns one matrix, no block barriers anywhere.
//
// Same mathematics, LAPACK dlarfg conventions and support envelopes as householder_columns /
// householder_qr_blocked.  Panels of kWNB = 8 columns:
//   * the panel (compact row list of at most 32 * kWR = 224 rows, lane l holds rows l, l + 32, ...) is
//     factored in the warp's shared-memory slice: per column one round of dot products against the remaining panel
//     columns (the norm is the first of them), one 5-stage butterfly, dlarfg scalars, one fused update;
//   * the reflectors go to the warp's shared-memory slice (reflector-major), the compact-WY T factor
//     comes from a Gram matrix on the FP64 tensor pipe;
//   * every group of 8 trailing columns is loaded ONCE (all tiles in flight together, straight into the
//     mma.sync.m8n8k4.f64 fragment layout), Y^T = C^T V, Y' = Y^T T, C^T -= Y' V^T run on the tensor
//     pipe, and the group is stored back.
// With eight independent warps per SM the long-latency phases of different members overlap without any
// coupling; the instruction count per member-step is several times below the CTA-per-member kernel,
// whose 256 threads shared panels of only ~2 k entries.
#pragma once
// included from ek1_warp.cuh (after ek1_device.cuh / qr_blocked.cuh: RowMap, panel_rows, dmma884)

namespace pnmol {

constexpr int kWNB = 8;   // panel width
constexpr int kWR = 7;    // row slots per lane (row lists up to 224)
constexpr int kWT = 28;   // 8-row tiles of a row list

struct WarpQR {   // the warp's shared-memory slice
    double* Vs;   // [kWNB][ldv] reflectors, ldv = 6 mod 16
    double* Ts;   // [kWNB][kWNB] T factor (row-major)
    double* Gs;   // [kWNB][kWNB + 1] Gram matrix
    double* tau;  // [kWNB]
    int ldv;
};

// Padded compact row list of a panel: the top segment (rows j0 .. j0+len1-1) is padded to a multiple of 8 rows so that
// every 8-row tile of the tensor-core phases lies in ONE segment and its rows are column base + constant.  Padding
// indices ("holes") carry zeros in V; their workspace addresses are real rows that are read but never written.
// When the panel straddles the top/bottom boundary (0 < len1 < kWNB) no padding is applied (aligned == false).
struct PadMap {
    int sh1, sh2p, len1, len1p, lend, Lp;
    bool aligned;
    __device__ __forceinline__ int row(int c) const { return c + (c < len1p ? sh1 : sh2p); }
    __device__ __forceinline__ bool valid(int c) const { return c < len1 || (c >= len1p && c < lend); }
};
__host__ __device__ __forceinline__ PadMap pad_map(int j0, int len1, int a2, int len) {
    PadMap p;
    p.sh1 = j0; p.len1 = len1;
    p.aligned = len1 == 0 || len1 >= kWNB;
    p.len1p = p.aligned ? (len1 + 7) & ~7 : len1;
    p.sh2p = a2 - p.len1p;
    p.lend = p.len1p + (len - len1);
    p.Lp = (p.lend + 7) & ~7;
    return p;
}

// Factor columns j0 .. j0+nbk-1 in the warp's shared-memory slice (which then holds V); tau -> q.tau, R written back
// to W.  Lane l owns rows l, l + 32, ... of the padded row list.  All loops are rolled (a few hundred instructions in
// total): eight desynchronised warps per SM must share the instruction cache.
__device__ __noinline__ void warp_panel_factor(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk,
                                               const PadMap pm, const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int nt = s.nt, Lp = pm.Lp, ldv = q.ldv;
    double* V = q.Vs;
    double* betas = q.Gs;  // (the Gram buffer is free during the panel factorisation)
    // ---- load with 8-byte asynchronous copies (entries outside a column's own envelope, and the padding, are zero)
    for (int ci = lane; ci < Lp; ci += 32) {
        const int row = pm.row(ci);
        const bool ok = pm.valid(ci);
        const double* src = W + (size_t)j0 * ld + row;
        double* dst = V + ci;
        for (int c = 0; c < kWNB; ++c) {
            const int j = j0 + c;
            if (ok && c < nbk && (row < nt ? row <= env_top(s, j) : row <= env_bot(s, j))) {
                const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d32), "l"(src) : "memory");
            } else {
                *dst = 0.0;
            }
            src += ld;
            dst += ldv;
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncwarp();
    for (int i = 0; i < nbk; ++i) {
        double* vi = V + i * ldv;
        // dot products of the pivot column (rows below the diagonal) with itself and the later panel columns
        double d[kWNB];
#pragma unroll
        for (int k = 0; k < kWNB; ++k) d[k] = 0.0;
        double ss = 0.0;
        for (int ci = lane; ci < Lp; ci += 32) {
            const double xi = ci > i ? vi[ci] : 0.0;
            ss = fma(xi, xi, ss);
            const double* vk = V + ci;
#pragma unroll
            for (int k = 1; k < kWNB; ++k) {
                vk += ldv;
                if (k > i) d[k] = fma(xi, *vk, d[k]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
#pragma unroll
            for (int k = 1; k < kWNB; ++k)
                if (k > i) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
        }
        const double al = vi[i];
        // dlarfg on (alpha, ||x||^2): beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta)
        double tau = 0.0, beta = al, scale = 0.0;
        if (ss != 0.0) {  // zero sub-column -> H = I
            const double nrm = sqrt(fma(al, al, ss));
            beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg (IEEE copysign, -0.0 counts as negative)
            tau = (beta - al) / beta;
            scale = 1.0 / (al - beta);
        }
        // g_k = -tau (x_k[i] + v . x_k): x_k += g_k v for the later columns
#pragma unroll
        for (int k = 1; k < kWNB; ++k)
            if (k > i) d[k] = -tau * fma(scale, d[k], V[k * ldv + i]);
        __syncwarp();  // every lane has read row i and the diagonal
        if (lane == 0) { q.tau[i] = tau; betas[i] = beta; }
        for (int ci = lane; ci < Lp; ci += 32) {
            if (ci >= i) {
                double* vp = V + ci;
                const double v = ci > i ? vp[i * ldv] * scale : 1.0;
                if (ci > i) vp[i * ldv] = tau != 0.0 ? v : 0.0;
                if (tau != 0.0) {
#pragma unroll
                    for (int k = 1; k < kWNB; ++k) {
                        vp += ldv;
                        if (k > i) *vp = fma(d[k], v, *vp);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (lane >= nbk && lane < kWNB) q.tau[lane] = 0.0;
    // ---- R (beta on the diagonal, zeros below) to the workspace; the buffer keeps V (unit diagonal, zero above)
    for (int ci = lane; ci < Lp; ci += 32) {
        const int row = pm.row(ci);
        const bool ok = pm.valid(ci);
        double* dst = W + (size_t)j0 * ld + row;
        double* vp = V + ci;
        for (int c = 0; c < nbk; ++c) {
            const int j = j0 + c;
            if (ok && (row < nt ? row <= env_top(s, j) : row <= env_bot(s, j))) *dst = ci < c ? *vp : (ci == c ? betas[c] : 0.0);
            if (ci <= c) *vp = (ci == c && q.tau[c] != 0.0) ? 1.0 : 0.0;
            dst += ld;
            vp += ldv;
        }
    }
    __syncwarp();
}

// T of the compact-WY representation (dlarft forward/columnwise) from the Gram matrix V^T V (tensor pipe).
__device__ __noinline__ void warp_t_factor(int Lp, int nbk, const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = Lp >> 3;
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
    const double* vp = q.Vs + g * q.ldv + 2 * t;
    for (int i = 0; i < ntile; ++i) {
        const double a = vp[8 * i], b = vp[8 * i + 1];
        dmma884(c0[0], c0[1], a, a);
        dmma884(c1[0], c1[1], b, b);
    }
    q.Gs[g * (kWNB + 1) + 2 * t] = c0[0] + c1[0];
    q.Gs[g * (kWNB + 1) + 2 * t + 1] = c0[1] + c1[1];
    __syncwarp();
    // lane k owns row k of T: T[k][i] = -tau_i sum_{j<i} T[k][j] G[j][i] (k < i), T[i][i] = tau_i
    double Trow[kWNB];
#pragma unroll
    for (int j = 0; j < kWNB; ++j) Trow[j] = 0.0;
#pragma unroll
    for (int i = 0; i < kWNB; ++i) {
        const double tau = i < nbk ? q.tau[i] : 0.0;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < i; ++j) a = fma(Trow[j], q.Gs[j * (kWNB + 1) + i], a);
        Trow[i] = lane < i ? -tau * a : (lane == i ? tau : 0.0);
    }
    if (lane < kWNB) {
#pragma unroll
        for (int j = 0; j < kWNB; ++j) q.Ts[lane * kWNB + j] = Trow[j];
    }
    __syncwarp();
}

// Apply the panel's block reflector to the trailing columns: C <- C - V T^T (V^T C), 8 columns at a time, one pass:
// all tiles of a column group are loaded up front (straight into the mma.sync.m8n8k4.f64 fragment layout).
// ALIGNED: every tile lies in one segment of the padded row list, so its two rows per lane are (segment base of the
// column) + constant; only the last tile of each segment needs validity masks.
template <bool ALIGNED>
__device__ __noinline__ void warp_trailing(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const PadMap pm,
                                           const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = pm.Lp >> 3;
    const int n1 = pm.len1p >> 3;        // tiles of the top segment (ALIGNED)
    const int ldv = q.ldv;
    const double tf0 = q.Ts[(2 * t) * kWNB + g], tf1 = q.Ts[(2 * t + 1) * kWNB + g];  // B fragments of T: T[2t + s][g]
    const double* v1 = q.Vs + g * ldv + 2 * t;      // product 1: V[row 8a + 2t + e][refl g]
    const double* v2a = q.Vs + (2 * t) * ldv + g;   // product 2: V[row 8a + g][refl 2t + s]
    const double* v2b = v2a + ldv;
    for (int kb = j0 + nbk; kb < ncols; kb += 8) {
        const int col = kb + g;
        const bool have = col < ncols;
        double* cp = W + (size_t)(have ? col : kb) * ld + 2 * t;
        double* cp1 = cp + pm.sh1;
        double* cp2 = cp + pm.sh2p;
        double xa[kWT][2];
#pragma unroll
        for (int a = 0; a < kWT; ++a) {
            if (a < ntile) {
                if (ALIGNED) {
                    const double* p = (a < n1 ? cp1 : cp2) + 8 * a;
                    xa[a][0] = p[0];
                    xa[a][1] = p[1];
                    if (a == n1 - 1 || a == ntile - 1) {  // tail tile of a segment: padding rows read as zero
                        if (!pm.valid(8 * a + 2 * t)) xa[a][0] = 0.0;
                        if (!pm.valid(8 * a + 2 * t + 1)) xa[a][1] = 0.0;
                    }
                } else {
                    const int c0 = 8 * a + 2 * t, c1 = c0 + 1;
                    xa[a][0] = pm.valid(c0) ? cp[pm.row(c0) - 2 * t] : 0.0;
                    xa[a][1] = pm.valid(c1) ? cp[pm.row(c1) - 2 * t] : 0.0;
                }
            }
        }
        // Y^T[col g][refl 2t, 2t + 1] = sum over rows
        double y0[2] = {0.0, 0.0}, y1[2] = {0.0, 0.0};
#pragma unroll
        for (int a = 0; a < kWT; ++a) {
            if (a < ntile) {
                dmma884(y0[0], y0[1], xa[a][0], v1[8 * a]);
                dmma884(y1[0], y1[1], xa[a][1], v1[8 * a + 1]);
            }
        }
        // Y'^T = Y^T T  (k-step s uses reflector 2t + s: the accumulator layout is the A layout)
        double z[2] = {0.0, 0.0};
        dmma884(z[0], z[1], y0[0] + y1[0], tf0);
        dmma884(z[0], z[1], y0[1] + y1[1], tf1);
        z[0] = -z[0]; z[1] = -z[1];
        // C^T -= Y'^T V^T, store
#pragma unroll
        for (int a = 0; a < kWT; ++a) {
            if (a < ntile) {
                dmma884(xa[a][0], xa[a][1], z[0], v2a[8 * a]);
                dmma884(xa[a][0], xa[a][1], z[1], v2b[8 * a]);
                if (ALIGNED) {
                    double* p = (a < n1 ? cp1 : cp2) + 8 * a;
                    if (a == n1 - 1 || a == ntile - 1) {
                        if (have && pm.valid(8 * a + 2 * t)) p[0] = xa[a][0];
                        if (have && pm.valid(8 * a + 2 * t + 1)) p[1] = xa[a][1];
                    } else if (have) {
                        p[0] = xa[a][0];
                        p[1] = xa[a][1];
                    }
                } else {
                    const int c0 = 8 * a + 2 * t, c1 = c0 + 1;
                    if (have && pm.valid(c0)) cp[pm.row(c0) - 2 * t] = xa[a][0];
                    if (have && pm.valid(c1)) cp[pm.row(c1) - 2 * t] = xa[a][1];
                }
            }
        }
    }
}

// Whole QR on one warp.  On return the upper triangle of W holds R.
__device__ void householder_qr_warp(double* __restrict__ W, int ld, const Shape s, const WarpQR& q, PhaseClock& pc) {
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    for (int j0 = 0; j0 < nref; j0 += kWNB) {
        const int nbk = nref - j0 < kWNB ? nref - j0 : kWNB;
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        const PadMap pm = pad_map(rm.j0, rm.len1, rm.a2, rm.len);
        warp_panel_factor(W, ld, s, j0, nbk, pm, q);
        pc.mark(9);
        if (j0 + nbk < s.ncols) {
            warp_t_factor(pm.Lp, nbk, q);
            pc.mark(14);
            if (pm.aligned) warp_trailing<true>(W, ld, s.ncols, j0, nbk, pm, q);
            else warp_trailing<false>(W, ld, s.ncols, j0, nbk, pm, q);
            pc.mark(11);
        }
        __syncwarp();
    }
}

}  // namespace pnmol
