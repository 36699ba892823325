// Warp-per-member Householder QR (sm_100a): one warp owns one matrix, no block barriers anywhere.
//
// Same mathematics, LAPACK dlarfg conventions and support envelopes as householder_columns /
// householder_qr_blocked.  Panels of kWNB = 8 columns:
//   * the panel (compact row list of at most 32 * kWR = 224 rows, lane l holds rows l, l + 32, ...) is
//     factored entirely in registers: per column one round of dot products against the remaining panel
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

// Factor columns j0 .. j0+nbk-1 in registers; V -> Vs, tau -> q.tau, R written back to W.
__device__ __noinline__ void warp_panel_factor(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk,
                                                  const RowMap rm, const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int nt = s.nt, L = rm.len;
    const int nr = (L + 31) >> 5;  // row slots in use
    double x[kWNB][kWR];
    double* betas = q.Gs;  // (the Gram buffer is free during the panel factorisation)
    // own envelope of column `lane` of the panel (lanes < nbk), broadcast by shuffles below
    const int jl = j0 + (lane < nbk ? lane : 0);
    const int my_et = lane < nbk ? env_top(s, jl) : -1, my_eb = lane < nbk ? env_bot(s, jl) : -1;
    if (lane < kWNB) { q.tau[lane] = 0.0; betas[lane] = 0.0; }
    double* const wp = W + (size_t)j0 * ld;
#pragma unroll
    for (int c = 0; c < kWNB; ++c) {
        const int et = __shfl_sync(0xffffffffu, my_et, c), eb = __shfl_sync(0xffffffffu, my_eb, c);
        const double* col = wp + (size_t)(c < nbk ? c : 0) * ld;
#pragma unroll
        for (int r = 0; r < kWR; ++r) {
            const int ci = lane + 32 * r;
            const int row = rm.row(ci);
            const bool ok = ci < L && (row < nt ? row <= et : row <= eb);
            x[c][r] = ok ? col[row] : 0.0;
        }
    }
#pragma unroll
    for (int i = 0; i < kWNB; ++i) {
        if (i < nbk) {
            // dot products of the pivot column (rows below the diagonal) with itself and the later panel columns
            double d[kWNB];
#pragma unroll
            for (int k = i; k < kWNB; ++k) {
                double a0 = lane > i ? x[i][0] * x[k][0] : 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 1; r < kWR; r += 2) {
                    if (r < nr) a1 = fma(x[i][r], x[k][r], a1);
                    if (r + 1 < kWR && r + 1 < nr) a0 = fma(x[i][r + 1], x[k][r + 1], a0);
                }
                d[k] = a0 + a1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = i; k < kWNB; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
            }
            const double al = __shfl_sync(0xffffffffu, x[i][0], i);
            const double ss = d[i];
            // dlarfg on (alpha, ||x||^2): beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta)
            double tau = 0.0, beta = al, scale = 0.0;
            if (ss != 0.0) {  // zero sub-column -> H = I
                const double s2 = fma(al, al, ss);
                const double nrm = sqrt(s2);
                beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg (IEEE copysign, -0.0 counts as negative)
                tau = (beta - al) / beta;
                scale = 1.0 / (al - beta);
            }
            if (lane == 0) { q.tau[i] = tau; betas[i] = beta; }
            if (tau != 0.0) {
                // v = scale * x below the diagonal (kept in x[i]); x_k -= tau (x_k[i] + v . x_k) v for the later columns
#pragma unroll
                for (int r = 0; r < kWR; ++r)
                    if (r < nr) x[i][r] = (r > 0 || lane > i) ? x[i][r] * scale : x[i][r];
#pragma unroll
                for (int k = i + 1; k < kWNB; ++k) {
                    const double ek = __shfl_sync(0xffffffffu, x[k][0], i);
                    const double gk = -tau * fma(scale, d[k], ek);
                    x[k][0] = lane > i ? fma(gk, x[i][0], x[k][0]) : (lane == i ? x[k][0] + gk : x[k][0]);
#pragma unroll
                    for (int r = 1; r < kWR; ++r)
                        if (r < nr) x[k][r] = fma(gk, x[i][r], x[k][r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < kWR; ++r)
                    if (r > 0 || lane > i) x[i][r] = 0.0;  // H = I: no reflector (the entries below the diagonal are zero)
            }
        }
    }
    // export: V (unit diagonal, zero above) to shared memory, R (beta on the diagonal, zeros below) to the workspace
    __syncwarp();
#pragma unroll
    for (int c = 0; c < kWNB; ++c) {
        const bool hc = c < nbk;
        const int et = __shfl_sync(0xffffffffu, my_et, c), eb = __shfl_sync(0xffffffffu, my_eb, c);
        double* col = wp + (size_t)(hc ? c : 0) * ld;
        double* vs = q.Vs + c * q.ldv;
        const bool live = hc && q.tau[c] != 0.0;
        const double beta = betas[c];
#pragma unroll
        for (int r = 0; r < kWR; ++r) {
            const int ci = lane + 32 * r;
            if (ci < q.ldv) vs[ci] = (live && ci < L) ? (ci > c ? x[c][r] : (ci == c ? 1.0 : 0.0)) : 0.0;
            const int row = rm.row(ci);
            if (hc && ci < L && (row < nt ? row <= et : row <= eb)) col[row] = ci < c ? x[c][r] : (ci == c ? beta : 0.0);
        }
    }
    for (int ci = 32 * kWR + lane; ci < q.ldv; ci += 32) {  // tail of the reflector rows (tiles read up to 8 * ntile)
#pragma unroll
        for (int c = 0; c < kWNB; ++c) q.Vs[c * q.ldv + ci] = 0.0;
    }
    __syncwarp();
}

// T of the compact-WY representation (dlarft forward/columnwise) from the Gram matrix V^T V (tensor pipe).
__device__ __noinline__ void warp_t_factor(int L, int nbk, const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (L + 7) >> 3;
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

// Apply the panel's block reflector to the trailing columns: C <- C - V T^T (V^T C), 8 columns at a time, one pass.
__device__ __noinline__ void warp_trailing(double* __restrict__ W, int ld, int ncols, int j0, int nbk, const RowMap rm,
                                              const WarpQR& q) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int L = rm.len, ntile = (L + 7) >> 3;
    const int sh1 = rm.j0, sh2 = rm.a2 - rm.len1, len1 = rm.len1;
    const int ldv = q.ldv;
    double tf[2];  // B fragments of T: T[2t + s][g]
    tf[0] = q.Ts[(2 * t) * kWNB + g];
    tf[1] = q.Ts[(2 * t + 1) * kWNB + g];
    const double* v1 = q.Vs + g * ldv + 2 * t;  // product 1: V[row 8i + 2t + e][refl g]
    const double* v2 = q.Vs + (2 * t) * ldv + g;  // product 2: V[row 8i + g][refl 2t + s]
    for (int kb = j0 + nbk; kb < ncols; kb += 8) {
        const int col = kb + g;
        const bool have = col < ncols;
        double* cp = W + (size_t)(have ? col : kb) * ld;
        double xa[kWT][2];
#pragma unroll
        for (int i4 = 0; i4 < kWT; i4 += 4) {
            if (i4 < ntile) {
#pragma unroll
                for (int a = i4; a < i4 + 4; ++a) {
                    const int c0 = 8 * a + 2 * t, c1 = c0 + 1;
                    xa[a][0] = (have && c0 < L) ? cp[c0 + (c0 < len1 ? sh1 : sh2)] : 0.0;
                    xa[a][1] = (have && c1 < L) ? cp[c1 + (c1 < len1 ? sh1 : sh2)] : 0.0;
                }
            }
        }
        // Y^T[col g][refl 2t, 2t + 1] = sum over rows
        double y0[2] = {0.0, 0.0}, y1[2] = {0.0, 0.0};
#pragma unroll
        for (int i4 = 0; i4 < kWT; i4 += 4) {
            if (i4 < ntile) {
#pragma unroll
                for (int a = i4; a < i4 + 4; ++a) {
                    if (a < ntile) {
                        dmma884(y0[0], y0[1], xa[a][0], v1[8 * a]);
                        dmma884(y1[0], y1[1], xa[a][1], v1[8 * a + 1]);
                    }
                }
            }
        }
        const double yt0 = y0[0] + y1[0], yt1 = y0[1] + y1[1];
        // Y'^T = Y^T T  (k-step s uses reflector 2t + s: the accumulator layout is the A layout)
        double z[2] = {0.0, 0.0};
        dmma884(z[0], z[1], yt0, tf[0]);
        dmma884(z[0], z[1], yt1, tf[1]);
        z[0] = -z[0]; z[1] = -z[1];
        // C^T -= Y'^T V^T, store
#pragma unroll
        for (int i4 = 0; i4 < kWT; i4 += 4) {
            if (i4 < ntile) {
#pragma unroll
                for (int a = i4; a < i4 + 4; ++a) {
                    if (a < ntile) {
                        dmma884(xa[a][0], xa[a][1], z[0], v2[8 * a]);
                        dmma884(xa[a][0], xa[a][1], z[1], v2[8 * a + ldv]);
                    }
                    const int c0 = 8 * a + 2 * t, c1 = c0 + 1;
                    if (have && c0 < L) cp[c0 + (c0 < len1 ? sh1 : sh2)] = xa[a][0];
                    if (have && c1 < L) cp[c1 + (c1 < len1 ? sh1 : sh2)] = xa[a][1];
                }
            }
        }
    }
}

// Whole QR on one warp.  On return the upper triangle of W holds R.
__device__ void householder_qr_warp(double* __restrict__ W, int ld, const Shape s, const WarpQR& q) {
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    for (int j0 = 0; j0 < nref; j0 += kWNB) {
        const int nbk = nref - j0 < kWNB ? nref - j0 : kWNB;
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        warp_panel_factor(W, ld, s, j0, nbk, rm, q);
        if (j0 + nbk < s.ncols) {
            warp_t_factor(rm.len, nbk, q);
            warp_trailing(W, ld, s.ncols, j0, nbk, rm, q);
        }
        __syncwarp();
    }
}

}  // namespace pnmol
