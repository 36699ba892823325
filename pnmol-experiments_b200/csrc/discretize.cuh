// Batched problem set-up on the device: the step BEFORE the filter loop, for sweeps over
// kernel hyper-parameters (one set per ensemble member).
//
//   k_fd_coefficients   probabilistic finite-difference stencils (src/pnmol/discretize.py:177-201 fd_coefficients, batched
//                       as in fd_probabilistic :61-77): one thread per (member, mesh point) builds the s x s kernel Gram
//                       matrix of the stencil (s <= 8) and the differentiated kernel row with closed-form 1-D kernel
//                       derivatives (the reference differentiates with JAX autodiff), solves the system by LU with
//                       partial pivoting (what jnp.linalg.solve / dgesv does) and returns the weights (a row of L) and
//                       the posterior variance (diagonal of E_sqrtm, quirk Q3).
//   k_gram_cholesky     spatial Gram matrix k(X, X) + nugget I and its Cholesky factor (src/pnmol/white.py:82-94
//                       initialize_iwp), one CTA per member; optionally the Gaussian log-likelihood of data y under that
//                       Gram matrix (src/pnmol/kernels.py:186-211 mle_input_scale / log_likelihood).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pnmol {

// kind: 0 SquareExponential, 1 Matern52, 2 Polynomial.  par: [input_scale, output_scale, order, const].
struct KPar { double input_scale, output_scale, order, cst; };

__device__ __forceinline__ double ipow(double b, int e) {  // b^e, e >= 0; 0 for e < 0 (a vanished derivative term)
    if (e < 0) return 0.0;
    double r = 1.0;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}

// k(x, y) in one dimension (kernels.py:107-111, 114-124, 127-144)
__device__ __forceinline__ double kernel_value(int kind, const KPar& p, double x, double y) {
    const double u = x - y;
    if (kind == 0) return p.output_scale * p.output_scale * exp(-(u * u) * (p.input_scale * p.input_scale) / 2.0);
    if (kind == 1) {
        const double a = sqrt(5.0 * (u * u) * (p.input_scale * p.input_scale));
        return p.output_scale * p.output_scale * (1.0 + a + a * a / 3.0) * exp(-a);
    }
    return ipow(x * y + p.cst, (int)p.order);
}

// which: 0 d/dx, 1 d2/dx2, 2 d2/dxdy, 3 d4/dx2dy2.  Matern52 returns NaN at x == y like the reference's autodiff
// (the caller substitutes the reference's constants, discretize.py:184-197).
__device__ __forceinline__ double kernel_derivative(int kind, const KPar& p, int which, double x, double y) {
    const double u = x - y;
    if (kind == 0) {
        const double q = p.input_scale * p.input_scale;
        const double k = p.output_scale * p.output_scale * exp(-q * u * u / 2.0);
        const double poly = which == 0 ? -q * u
                          : which == 1 ? q * q * u * u - q
                          : which == 2 ? q - q * q * u * u
                                       : 3.0 * q * q - 6.0 * q * q * q * u * u + q * q * q * q * u * u * u * u;
        return poly * k;
    }
    if (kind == 1) {
        if (u == 0.0) return nan("");
        const double c = sqrt(5.0) * p.input_scale;
        const double a = c * fabs(u);
        const double s2 = p.output_scale * p.output_scale, ex = exp(-a);
        if (which == 0) return -s2 / 3.0 * c * (u > 0.0 ? 1.0 : -1.0) * a * (1.0 + a) * ex;
        if (which == 1) return -s2 / 3.0 * c * c * (1.0 + a - a * a) * ex;
        if (which == 2) return s2 / 3.0 * c * c * (1.0 + a - a * a) * ex;
        return -s2 / 3.0 * c * c * c * c * (-a * a + 5.0 * a - 3.0) * ex;
    }
    const int o = (int)p.order;
    const double b = x * y + p.cst;
    if (which == 0) return o * y * ipow(b, o - 1);
    if (which == 1) return o * (o - 1) * y * y * ipow(b, o - 2);
    if (which == 2) return o * ipow(b, o - 1) + o * (o - 1) * x * y * ipow(b, o - 2);
    return o * (o - 1) * (2.0 * ipow(b, o - 2) + 4.0 * x * y * (o - 2) * ipow(b, o - 3)
                          + x * x * y * y * (o - 2) * (o - 3) * ipow(b, o - 4));
}

constexpr int kMaxStencil = 8;

// diffop: 0 gradient (first = d/dx, second = d2/dxdy), 1 laplace (first = d2/dx2, second = d4/dx2dy2).
// par [nbatch][4]; x [P]; nbrs [P][s]; weights [nbatch][P][s]; unc [nbatch][P].
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void k_fd_coefficients(int kind, const double* __restrict__ par, int nbatch, const double* __restrict__ x,
                                  const double* __restrict__ nbrs, int P, int s, int diffop, double nugget,
                                  double* __restrict__ weights, double* __restrict__ unc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nbatch * P) return;
    const int b = (int)(idx / P), pt = (int)(idx - (size_t)b * P);
    KPar kp;
    kp.input_scale = par[4 * b]; kp.output_scale = par[4 * b + 1]; kp.order = par[4 * b + 2]; kp.cst = par[4 * b + 3];
    const int first = diffop == 0 ? 0 : 1, second = diffop == 0 ? 2 : 3;
    double G[kMaxStencil][kMaxStencil], rhs[kMaxStencil], w[kMaxStencil], xn[kMaxStencil];
    const double xp = x[pt];
    for (int i = 0; i < s; ++i) xn[i] = nbrs[(size_t)pt * s + i];
    double mat_a = 0.0, mat_b = 0.0;
    if (kind == 1) {  // Taylor-series constants of the Matern52 at zero (discretize.py:184-197)
        const double r = kp.input_scale, sc = kp.output_scale;
        mat_a = r * r * sc * sc * 2.5 / (1.0 - 2.5);
        mat_b = sc * sc * r * r * r * r * 3.0 * 2.5 * 2.5 / (2.0 - 3.0 * 2.5 + 2.5 * 2.5);
    }
    for (int i = 0; i < s; ++i) {
        for (int j = 0; j < s; ++j) G[i][j] = kernel_value(kind, kp, xn[i], xn[j]) + (i == j ? nugget : 0.0);
        double d = kernel_derivative(kind, kp, first, xp, xn[i]);
        if (kind == 1 && isnan(d)) d = mat_a;
        rhs[i] = d;
        w[i] = d;
    }
    // LU with partial pivoting (dgesv), right-hand side carried along
    for (int k = 0; k < s; ++k) {
        int piv = k;
        double best = fabs(G[k][k]);
        for (int i = k + 1; i < s; ++i)
            if (fabs(G[i][k]) > best) { best = fabs(G[i][k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < s; ++j) { const double t = G[k][j]; G[k][j] = G[piv][j]; G[piv][j] = t; }
            const double t = w[k]; w[k] = w[piv]; w[piv] = t;
        }
        const double inv = 1.0 / G[k][k];
        for (int i = k + 1; i < s; ++i) {
            const double l = G[i][k] * inv;
            for (int j = k + 1; j < s; ++j) G[i][j] = fma(-l, G[k][j], G[i][j]);
            w[i] = fma(-l, w[k], w[i]);
        }
    }
    for (int k = s - 1; k >= 0; --k) {
        double acc = w[k];
        for (int j = k + 1; j < s; ++j) acc = fma(-G[k][j], w[j], acc);
        w[k] = acc / G[k][k];
    }
    double top = kernel_derivative(kind, kp, second, xp, xp);
    if (kind == 1 && isnan(top)) top = mat_b;
    double dot = 0.0;
    for (int i = 0; i < s; ++i) dot = fma(w[i], rhs[i], dot);
    for (int i = 0; i < s; ++i) weights[((size_t)b * P + pt) * s + i] = w[i];
    unc[(size_t)b * P + pt] = top - dot;
}
#else
__global__ void k_fd_coefficients(int kind, const double* __restrict__ par, int nbatch, const double* __restrict__ x,
                                  const double* __restrict__ nbrs, int P, int s, int diffop, double nugget,
                                  double* __restrict__ weights, double* __restrict__ unc);
#endif

// One CTA per member: A = k(X, X) + (white^2 + nugget) I, A = L L^T (lower factor, zeros above the diagonal written),
// optional log-likelihood of y:  -0.5 (y^T A^-1 y + log det A + n log 2 pi).  The factorisation runs in shared memory
// when `in_smem` (d * (d + 1) doubles fit), otherwise in place in the output buffer (L2).
#if defined(PNMOL_TU_ALL) || defined(PNMOL_TU_MISC)
__global__ void __launch_bounds__(256) k_gram_cholesky(int kind, const double* __restrict__ par, int nbatch,
                                                      const double* __restrict__ X, int d, double diag_add,
                                                      const double* __restrict__ y, double* __restrict__ Lout,
                                                      double* __restrict__ loglik, int32_t* __restrict__ status, int in_smem) {
    extern __shared__ __align__(16) double sm_raw[];
    __shared__ double red[16];
    __shared__ int bad;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        KPar kp;
        kp.input_scale = par[4 * b]; kp.output_scale = par[4 * b + 1]; kp.order = par[4 * b + 2]; kp.cst = par[4 * b + 3];
        double* out = Lout + (size_t)b * d * d;
        double* A = in_smem ? sm_raw : out;
        const int lda = in_smem ? d + 1 : d;  // odd pitch: conflict-free column walks in shared memory
        if (tid == 0) bad = 0;
        for (int e = tid; e < d * d; e += nthr) {
            const int i = e / d, j = e - i * d;
            if (j <= i) A[i * lda + j] = kernel_value(kind, kp, X[i], X[j]) + (i == j ? diag_add : 0.0);
        }
        __syncthreads();
        for (int j = 0; j < d; ++j) {
            // column j: L[j][j] = sqrt(A[j][j]), L[i][j] = A[i][j] / L[j][j]
            const double ajj = A[j * lda + j];
            __syncthreads();
            const double ljj = sqrt(ajj);
            if (tid == 0) {
                A[j * lda + j] = ljj;
                if (!(ajj > 0.0)) bad = 1;  // not positive definite (NaN factor, like numpy raises / jax returns NaN)
            }
            const double inv = 1.0 / ljj;
            for (int i = j + 1 + tid; i < d; i += nthr) A[i * lda + j] *= inv;
            __syncthreads();
            // trailing update of the lower triangle: A[i][k] -= L[i][j] L[k][j], j < k <= i
            const int nrem = d - j - 1;
            for (int e = tid; e < nrem * nrem; e += nthr) {
                const int ii = e / nrem, kk = e - ii * nrem;
                if (kk <= ii) {
                    const int i = j + 1 + ii, k = j + 1 + kk;
                    A[i * lda + k] = fma(-A[i * lda + j], A[k * lda + j], A[i * lda + k]);
                }
            }
            __syncthreads();
        }
        if (loglik) {  // forward substitution L w = y (thread 0 walks the rows; the dot products are spread over the CTA)
            double* wv = in_smem ? sm_raw + (size_t)d * lda : nullptr;
            double quad = 0.0, logdet = 0.0;
            if (wv) {
                for (int i = 0; i < d; ++i) {
                    double part = 0.0;
                    for (int k = tid; k < i; k += nthr) part = fma(A[i * lda + k], wv[k], part);
                    // block sum in a fixed order
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if ((tid & 31) == 0) red[tid >> 5] = part;
                    __syncthreads();
                    if (tid == 0) {
                        double s = 0.0;
                        for (int w8 = 0; w8 < (nthr >> 5); ++w8) s += red[w8];
                        wv[i] = (y[i] - s) / A[i * lda + i];
                    }
                    __syncthreads();
                }
                if (tid == 0) {
                    for (int i = 0; i < d; ++i) { quad = fma(wv[i], wv[i], quad); logdet += 2.0 * log(A[i * lda + i]); }
                    loglik[b] = -0.5 * (quad + logdet + d * log(2.0 * 3.14159265358979323846));
                }
            }  // (the host rejects a log-likelihood request when the factorisation does not fit shared memory)
        }
        __syncthreads();
        if (in_smem) {
            for (int e = tid; e < d * d; e += nthr) {
                const int i = e / d, j = e - i * d;
                out[e] = j <= i ? A[i * lda + j] : 0.0;
            }
        } else {
            for (int e = tid; e < d * d; e += nthr) {
                const int i = e / d, j = e - i * d;
                if (j > i) out[e] = 0.0;
            }
        }
        if (tid == 0 && status) status[b] = bad;
        __syncthreads();
    }
}
#else
__global__ void __launch_bounds__(256) k_gram_cholesky(int kind, const double* __restrict__ par, int nbatch,
                                                      const double* __restrict__ X, int d, double diag_add,
                                                      const double* __restrict__ y, double* __restrict__ Lout,
                                                      double* __restrict__ loglik, int32_t* __restrict__ status, int in_smem);
#endif

}  // namespace pnmol
