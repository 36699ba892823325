// Device code of the EK1 filter loop (one CTA per ensemble member).
//
// Reference path restated here (file:line in schmidtjonathan/pnmol-experiments):
//   attempt_step      src/pnmol/white.py:96-146, src/pnmol/latent.py:155-225
//   evaluate_ode      src/pnmol/white.py:169-208, src/pnmol/latent.py:237-292
//   estimate_error    src/pnmol/white.py:153-162
//   sqrt propagation  src/pnmol/base/sqrt.py:9-23, update src/pnmol/base/sqrt.py:34-95
//   initialize        src/pnmol/white.py:12-80, src/pnmol/latent.py:20-134
//
// Layout: each CTA owns a column-major workspace W (leading dimension ld = 2D) of
// m + D columns that stays L2-resident.  Columns m..m+D-1 first hold the predict stack
// [(A P^-1 Cl)^T ; Ql^T] (2D x D); its R factor (= Clp^T) is then, in place, the top-right
// block of the update matrix [[Clp^T H^T, Clp^T], [E^T, 0]] whose left m columns live in
// W's columns 0..m-1.  Householder QR follows LAPACK dlarfg conventions
// (beta = -sign(alpha) * norm, H = I for a zero sub-column) on the reference's own
// matrix layout, touching only the rows inside per-column support envelopes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pnmol {

#ifndef PNMOL_THREADS
#define PNMOL_THREADS 256
#endif
#ifndef PNMOL_MIN_CTAS
#define PNMOL_MIN_CTAS 2
#endif
#ifdef PNMOL_PLAIN_STATE_IO   // tuning: plain instead of streaming (evict-first) accesses to the filter state
#define PNMOL_STATE_LOAD(p) (*(p))
#define PNMOL_STATE_STORE(p, v) (*(p) = (v))
#else
#define PNMOL_STATE_LOAD(p) __ldcs(p)
#define PNMOL_STATE_STORE(p, v) __stcs(p, v)
#endif
constexpr int kThreads = PNMOL_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxN = 8;     // num_derivatives + 1
constexpr int kMaxComp = 4;  // PDE components

enum EMode {
    E_STEP_WHITE = 0,      // E = blockdiag(E_sqrtm, R_sqrtm)           white.py:184
    E_NONE = 1,            // no measurement noise                       latent.py:197
    E_NUGGET_ONLY = 2,     // nugget * I                                 white.py:33-38, latent.py:71-76,98-103
    E_STEP_PLUS_NUGGET = 3 // blockdiag(E_sqrtm, R_sqrtm) + nugget * I   white.py:51-56
};

struct Problem {
    int latent, semilinear, reaction;
    int d, n, nb, ncomp, npts, dd, D, m;
    int wl, wb, wh, ld, batch, nparams, vld, ldm;  // ldm > 0: an m x ldm shared-memory scratch exists (small m)
    int whs;                                       // = wh when the rows of H live in shared memory, else 0
    const int32_t* Lcol; const double* Lval; const double* Ediag;
    const int32_t* Bcol; const double* Bval; const double* Rsq;
    const double* A1d; const double* LQ1d; const double* Lk; const double* Kg;
    const double* diffscale; const double* priorscale; const double* rparams;
    const int32_t* te_p; const int32_t* be_p; const int32_t* te_u; const int32_t* be_u;     // triangular input factor
    const int32_t* te_pd;                                                                    // dense input factor
    double* W; int32_t* Hcol; double* Hval; double* F; double* S;
    unsigned long long* prof;  // optional clock64 phase accumulators (diagnostics)
    int* smslot;               // [number of SMs] zeroed before every launch: CTAs count themselves per SM
    unsigned* gsync;           // pace-keeping counter of this launch (k_run), zeroed before the launch; nullptr: off
};

// Phase timer: thread 0 of every CTA adds the cycles since the previous mark to prof[idx].
// Compiled in only with -DPNMOL_PROFILE=1 (tools/build_variant.sh prof -DPNMOL_PROFILE=1): the object is passed by reference
// through __noinline__ functions, i.e. it lives in local memory, and even a disabled mark() costs a local-memory load
// on the critical path of the panel chain.  The shipped library carries no marks (pnmol_b200_profile then reports zeros).
#ifndef PNMOL_PROFILE
#define PNMOL_PROFILE 0
#endif
struct PhaseClock {
#if !PNMOL_PROFILE
    __device__ __forceinline__ void start(unsigned long long*) {}
    __device__ __forceinline__ void mark(int) {}
#else
    unsigned long long* prof;
    long long last;
    __device__ __forceinline__ void start(unsigned long long* p) { prof = p; if (prof && threadIdx.x == 0) last = clock64(); }
    __device__ __forceinline__ void mark(int idx) {
        if (prof && threadIdx.x == 0) {
            const long long now = clock64();
            // fire-and-forget reduction (a generic-pointer atomicAdd waits for its round trip and inflates the next phase)
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(prof + idx), "l"((unsigned long long)(now - last)) : "memory");
            last = now;
        }
    }
#endif
};

struct Shape {  // QR row structure: rows [0,nt) top, [nt, nt+nbot) bottom
    int nt, nbot, ncols;
    const int32_t* te;  // nullptr = dense
    const int32_t* be;
    int ldr = 0;  // > 0: rows that exist in every workspace column; panel row lists are then aligned to 8-row tiles
};

struct FastQR {       // shared-memory buffers of the blocked QR (qr_fast.cuh), as offsets in doubles from the base of the
    unsigned buf[2];   // dynamic shared memory: 2 x (16 x LP) panel / reflector buffers, XOR-swizzled reflector-major
    unsigned Ts[2];    // 2 x (16 x 18): T factors
    unsigned Gs;       // 16 x 17 Gram matrix V^T V (upper triangle)
    unsigned scratch;  // 4 x 192 per-warp Gram partials
    unsigned tau;      // 2 x 16
    unsigned t4;       // 4 x (4 x 4): T factors of the sub-panels of the current panel
    unsigned ctx;      // 40 doubles: QRCtx, the driver's uniform state (qr_fast.cuh)
    int LP;            // rows of a buffer: 64, 128 or 256, >= every panel row list
    int slot;          // which co-resident CTA of its SM this is (0, 1, ...): decides the factor warp's scheduler
};

struct Smem {
    double *vbuf, *mp, *z, *y, *xw, *xat, *red, *pv, *pinv, *Vs, *xraw, *sc, *msq, *Vr, *Ts, *Gs;
    double* ekey = nullptr;  // 4 doubles: key of the cached error-estimate factor (error_estimate_smem), or nullptr
    FastQR fq;
    double *fqbase, *fqend;
    int32_t *te_p, *be_p, *te_pd, *te_u, *be_u;  // shared-memory copies of the QR envelopes (Problem::te_p ...)
    double* Hval; int32_t *Hcol, *Hpt;           // sparse rows of H (m x wh; Hpt = mesh point of an entry) when they fit in
                                                 // shared memory, else nullptr
    double* tri = nullptr;                       // multi-CTA path: shared-memory scratch of the blocked triangular solves
                                                 // (2 m + kTriScratch doubles), else nullptr
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) s += red[i];
    return s;
}

// A "team" is the set of threads that works on one ensemble member: the whole CTA on the blocked path, one warp on
// the warp-per-member path (ek1_warp.cuh).  The O(D^2) phases below are written once against this interface.
struct CtaTeam {
    static constexpr int size = kThreads, nwarps = kWarps;
    static __device__ __forceinline__ int tid() { return threadIdx.x; }
    static __device__ __forceinline__ int warp() { return threadIdx.x >> 5; }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ double sum(double v, double* red) { return block_sum(v, red); }
};
struct WarpTeam {
    static constexpr int size = 32, nwarps = 1;
    static __device__ __forceinline__ int tid() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int warp() { return 0; }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ double sum(double v, double*) { return warp_sum(v); }
};

__host__ __device__ __forceinline__ size_t fastqr_doubles(int LP) {
    return 2 * (size_t)16 * LP + 2 * (size_t)16 * 18 + 4 * 192 + 2 * 16 + 64 + 40;
}

// vld = rows of a panel buffer (LP of the blocked QR: 64, 128 or 256)
// wh: ELL width of the sparse rows of H (> 0: they live in shared memory; 0: in the global scratch Problem::Hcol/Hval)
__host__ __device__ __forceinline__ size_t smem_doubles(int D, int m, int dd, int vld, int ldm, int wh) {
    // (the m x ldm scratch of the error estimate and of the triangular solves aliases the panel buffers of the QR, which are
    // idle then; it only takes extra room beyond their 2 x 16 x vld doubles)
    const size_t msq = (size_t)m * ldm, bufs = 2 * (size_t)16 * vld;
    return (size_t)(2 * D + 4) + D + 3 * (size_t)m + dd + 16 + 2 * kMaxN + 80 + 4 + 1 + fastqr_doubles(vld) + (msq > bufs ? msq - bufs : 0) + 8 +
           (3 * (size_t)D + 2 * ((size_t)m + D) + 2) / 2 + 2 * (size_t)m * wh;
}

__device__ __forceinline__ Smem carve(double* base, int D, int m, int dd, int vld, int ldm, int wh = 0) {
    Smem s;
    double* const base0 = base;
    s.vbuf = base;              base += 2 * D + 4;
    s.mp = base;                base += D;
    s.z = base;                 base += m;
    s.y = base;                 base += m;
    s.xw = base;                base += m;
    s.xat = base;               base += dd;
    s.red = base;               base += 16;
    s.pv = base;                base += kMaxN;
    s.pinv = base;              base += kMaxN;
    s.sc = base;                base += 80;
    s.ekey = base;              base += 4;
    s.xraw = nullptr; s.Vs = nullptr; s.Vr = nullptr; s.Ts = nullptr; s.Gs = nullptr;
    s.fq.LP = vld;
    s.fq.slot = 0;
    if ((base - base0) & 1) ++base;  // 16-byte alignment of the panel buffers (128-bit operand loads)
    s.fqbase = base;            // (zeroed at kernel start: reflector buffers start finite)
    s.fq.buf[0] = (unsigned)(base - base0); base += (size_t)16 * vld;
    s.fq.buf[1] = (unsigned)(base - base0); base += (size_t)16 * vld;
    s.fq.Ts[0] = (unsigned)(base - base0);  base += 16 * 18;
    s.fq.Ts[1] = (unsigned)(base - base0);  base += 16 * 18;
    s.fq.Gs = (unsigned)(base - base0);     // (unused by the CTA-per-member path; no room reserved)
    s.fq.scratch = (unsigned)(base - base0); base += 4 * 192;
    s.fq.tau = (unsigned)(base - base0);    base += 2 * 16;
    s.fq.t4 = (unsigned)(base - base0);     base += 64;
    s.fq.ctx = (unsigned)(base - base0);    base += 40;
    s.fqend = base;
    s.msq = ldm > 0 ? s.fqbase : nullptr;   // aliases the panel buffers (never live at the same time)
    {
        const size_t msq = (size_t)m * ldm, bufs = 2 * (size_t)16 * vld;
        base += (msq > bufs ? msq - bufs : 0) + 8;
    }
    int32_t* ib = reinterpret_cast<int32_t*>(base);
    s.te_p = ib;  ib += D;
    s.be_p = ib;  ib += D;
    s.te_pd = ib; ib += D;
    s.te_u = ib;  ib += m + D;
    s.be_u = ib;  ib += m + D;
    ib += (3 * D + 2 * (m + D)) & 1;
    s.Hval = wh > 0 ? reinterpret_cast<double*>(ib) : nullptr;
    s.Hcol = wh > 0 ? reinterpret_cast<int32_t*>(s.Hval + (size_t)m * wh) : nullptr;
    s.Hpt = wh > 0 ? s.Hcol + (size_t)m * wh : nullptr;
    return s;
}

// Which co-resident CTA of this SM am I?  (Order of arrival on the SM; block-uniform, contains a block barrier.)
__device__ __forceinline__ int sm_slot(const Problem& P) {
    __shared__ int slot_s;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        slot_s = P.smslot ? atomicAdd(&P.smslot[smid], 1) : 0;
    }
    __syncthreads();
    return slot_s;
}

// Copy the envelope arrays of the problem into shared memory (once per kernel; followed by a block barrier at the caller).
__device__ __forceinline__ void load_envelopes(const Problem& P, const Smem& sm) {
    for (int i = threadIdx.x; i < P.D; i += kThreads) { sm.te_p[i] = P.te_p[i]; sm.be_p[i] = P.be_p[i]; sm.te_pd[i] = P.te_pd[i]; }
    for (int i = threadIdx.x; i < P.m + P.D; i += kThreads) { sm.te_u[i] = P.te_u[i]; sm.be_u[i] = P.be_u[i]; }
}

// ---------------------------------------------------------------- Householder QR
// Unblocked, structure-aware, on the global (L2-resident) workspace.  After return the
// upper triangle holds R and every entry below the diagonal inside the support envelope
// is exactly zero.
static __device__ void householder_columns(double* __restrict__ W, int ld, const Shape& s, int jbeg, int jend, double* vbuf,
                                    double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nrows = s.nt + s.nbot;
    for (int j = jbeg; j < jend; ++j) {
        int a1 = j, e1, a2, e2;
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
        const int len1 = e1 - a1 + 1;
        const int len2 = e2 >= a2 ? e2 - a2 + 1 : 0;
        const int len = len1 + len2;
        double* col = W + (size_t)j * ld;
        double ss = 0.0;
        for (int c = tid; c < len; c += kThreads) {
            const int r = c < len1 ? a1 + c : a2 + (c - len1);
            const double x = col[r];
            vbuf[c] = x;
            if (c > 0) ss += x * x;
        }
        ss = block_sum(ss, red);
        const double alpha = vbuf[0];
        double tau = 0.0, beta = alpha, scale = 0.0;
        if (ss != 0.0) {  // dlarfg: xnorm == 0 -> H = I
            const double nrm = sqrt(alpha * alpha + ss);
            beta = -copysign(nrm, alpha);  // Fortran SIGN semantics of dlarfg
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncthreads();
        if (tau != 0.0) {
            for (int c = tid; c < len; c += kThreads) {
                const int r = c < len1 ? a1 + c : a2 + (c - len1);
                vbuf[c] = c == 0 ? 1.0 : vbuf[c] * scale;
                col[r] = c == 0 ? beta : 0.0;
            }
        }
        __syncthreads();
        if (tau != 0.0) {
            for (int k = j + 1 + warp; k < s.ncols; k += kWarps) {
                double* ck = W + (size_t)k * ld;
                double dot = 0.0;
                for (int c = lane; c < len; c += 32) {
                    const int r = c < len1 ? a1 + c : a2 + (c - len1);
                    dot = fma(vbuf[c], ck[r], dot);
                }
                dot = warp_sum(dot);
                const double w = tau * dot;
                if (w != 0.0) {
                    for (int c = lane; c < len; c += 32) {
                        const int r = c < len1 ? a1 + c : a2 + (c - len1);
                        ck[r] = fma(-w, vbuf[c], ck[r]);
                    }
                }
            }
        }
        __syncthreads();
    }
}

static __device__ void householder_qr(double* __restrict__ W, int ld, const Shape s, double* vbuf, double* red) {
    const int nrows = s.nt + s.nbot;
    householder_columns(W, ld, s, 0, nrows < s.ncols ? nrows : s.ncols, vbuf, red);
}

// Scratch for the tensor-core trailing update (per-warp partial tiles): the m x ldm buffer is free during a QR.
__device__ __forceinline__ double* qr_scratch(const Problem& P, const Smem& sm) {
    return (sm.msq != nullptr && P.m * P.ldm >= 2 * kWarps * 128) ? sm.msq : nullptr;
}

}  // namespace pnmol

#include "qr_blocked.cuh"
#include "qr_fast.cuh"

namespace pnmol {

// ---------------------------------------------------------------- reaction terms
// f and df of src/pnmol/pde/examples.py:151-165 (SIR), :228-235 (Lotka-Volterra),
// :311-315 (spruce budworm); all point-wise, so df is ncomp x ncomp per mesh point.
__device__ __forceinline__ void reaction_point(int id, const double* prm, const double* x, double* f, double* J) {
    if (id == 1) {
        const double g = prm[0];
        f[0] = g * x[0] * (1.0 - x[0]);
        J[0] = g * (1.0 - 2.0 * x[0]);
    } else if (id == 2) {
        const double beta = prm[0], gamma = prm[1];
        const double s = x[0], i = x[1], r = x[2];
        const double tot = s + i + r;
        const double g = beta * s * i / tot;
        const double gs = beta * i / tot - g / tot;
        const double gi = beta * s / tot - g / tot;
        const double gr = -g / tot;
        f[0] = -g; f[1] = g - gamma * i; f[2] = gamma * i;
        J[0] = -gs; J[1] = -gi; J[2] = -gr;
        J[3] = gs;  J[4] = gi - gamma; J[5] = gr;
        J[6] = 0.0; J[7] = gamma; J[8] = 0.0;
    } else if (id == 3) {
        const double a = prm[0], b = prm[1], c = prm[2], dd = prm[3];
        const double u = x[0], v = x[1];
        f[0] = a * u - b * u * v;
        f[1] = c * u * v - dd * v;
        J[0] = a - b * v; J[1] = -b * u;
        J[2] = c * v;     J[3] = c * u - dd;
    }
}

// ---------------------------------------------------------------- evaluate_ode
// Builds z (smem) and the sparse rows of H (global scratch, ELL width wh) from the
// predicted mean mp (smem), with projections p0 = E0 P, p1 = E1 P reduced to the two
// scalars p0s, p1s.  white.py:169-208, latent.py:237-292.
template <class T = CtaTeam>
__device__ void evaluate_ode(const Problem& P, int b, const Smem& sm, double p0s, double p1s, int32_t* Hcol,
                             double* Hval) {
    const int tid = T::tid();
    const int n = P.n, d = P.d;
    for (int j = tid; j < P.dd; j += T::size) sm.xat[j] = p0s * sm.mp[j * n];
    T::sync();
    const double* prm = P.rparams ? P.rparams + (size_t)b * P.nparams : nullptr;
    for (int i = tid; i < d; i += T::size) {
        const int comp = i / P.npts, pt = i - comp * P.npts;
        const double ds = P.diffscale ? P.diffscale[(size_t)b * P.ncomp + comp] : 1.0;
        int32_t* hc = Hcol + (size_t)i * P.wh;
        double* hv = Hval + (size_t)i * P.wh;
        double acc = 0.0;
        int w = 0;
        for (; w < P.wl; ++w) {
            const int c = P.Lcol[(size_t)i * P.wl + w];
            if (c >= 0) {
                const double lv = ds * P.Lval[(size_t)i * P.wl + w];
                acc = fma(lv, sm.xat[c], acc);
                hc[w] = c * n;
                hv[w] = -p0s * lv;
            } else {
                hc[w] = -1;
                hv[w] = 0.0;
            }
        }
        double shift = 0.0;
        if (P.semilinear) {
            double x[kMaxComp], f[kMaxComp], J[kMaxComp * kMaxComp];
            for (int c = 0; c < P.ncomp; ++c) x[c] = sm.xat[c * P.npts + pt];
            reaction_point(P.reaction, prm, x, f, J);
            double jx = 0.0;
            for (int c = 0; c < P.ncomp; ++c) {
                const double jv = J[comp * P.ncomp + c];
                jx = fma(jv, x[c], jx);
                hc[w] = (c * P.npts + pt) * n;
                hv[w] = -p0s * jv;
                ++w;
            }
            acc += jx;
            shift = jx - f[comp];
        }
        hc[w] = i * n + 1;
        hv[w] = p1s;
        ++w;
        double hz = p1s * sm.mp[i * n + 1] - acc;
        if (P.latent) {
            hc[w] = (d + i) * n;
            hv[w] = -p0s;
            ++w;
            hz -= sm.xat[d + i];
        }
        for (; w < P.wh; ++w) { hc[w] = -1; hv[w] = 0.0; }
        sm.z[i] = hz + shift;
    }
    for (int r = tid; r < P.nb; r += T::size) {
        int32_t* hc = Hcol + (size_t)(d + r) * P.wh;
        double* hv = Hval + (size_t)(d + r) * P.wh;
        double acc = 0.0;
        int w = 0;
        for (; w < P.wb; ++w) {
            const int c = P.Bcol[(size_t)r * P.wb + w];
            if (c >= 0) {
                const double bv = P.Bval[(size_t)r * P.wb + w];
                acc = fma(bv, sm.xat[c], acc);
                hc[w] = c * n;
                hv[w] = p0s * bv;
            } else {
                hc[w] = -1;
                hv[w] = 0.0;
            }
        }
        for (; w < P.wh; ++w) { hc[w] = -1; hv[w] = 0.0; }
        sm.z[d + r] = acc;
    }
    T::sync();
}

// ---------------------------------------------------------------- predict stack
// Columns m..m+D-1 of W: top D rows (A P^-1 Cl)^T, bottom D rows Ql^T.
// white.py:104,118 / latent.py:179,194 and iwp.py:32-53, stacked_ssm.py:16-26.
// (w0, nw): this warp's index and the number of warps sharing the loop (CTA-local by default; grid-wide on the
// multi-CTA path, which synchronises with a grid barrier afterwards).
// pv_prev != nullptr (fused time loop, CTA-per-member path): the input factor is NOT read from the state buffer but taken
// in place from where the previous step's update QR left it -- row r of the factor is column m + r of R, rows m..m + r
// (see update_output_factor), i.e. Cl[r][k] = pv_prev[r % n] * Wp[r ld + m + k] with the previous step's Nordsieck
// scaling pv_prev -- so that a step that is neither the last one nor part of a requested trajectory never writes and
// re-reads the D x D factor.  Same operations in the same order as the route through the state buffer (bitwise equal).
template <class T = CtaTeam>
__device__ void build_predict(const Problem& P, int b, const Smem& sm, const double* __restrict__ Cl,
                              const int32_t* te, double* Wp, int w0 = T::warp(), int nw = T::nwarps,
                              const double* pv_prev = nullptr, int nrows_prev = 0) {
    const int lane = T::tid() & 31;
    const int n = P.n, D = P.D, nd = P.n * P.d;
    const double ps = P.priorscale ? P.priorscale[b] : 1.0;
    double pvp[kMaxN];
#pragma unroll
    for (int s = 0; s < kMaxN; ++s) pvp[s] = (pv_prev && s < n) ? pv_prev[s] : 1.0;
    // One n x n block of columns per warp and round: the n columns i = blk n + ii are combinations of the SAME n rows
    // blk n + s of the input factor (A = I (x) A_1d), so those rows are read once; two lane-chunks per round keep
    // 2 n independent (streaming, HBM) loads in flight.
    const int nblk = D / n;
    for (int blk = w0; blk < nblk; blk += nw) {
        const int i0 = blk * n;
        const int tend = te[i0];  // (the columns of a block share their top envelope)
        const double* src = Cl + (size_t)i0 * D;
        double* col0 = Wp + (size_t)i0 * P.ld;
        if (pv_prev && n == 3 && tend < 192) {
            // common case (two derivatives, fused time loop): ALL rows of the block are read before the first store --
            // one L2 round trip per block instead of one per 64 rows; same operations in the same order as below
            double a[3][3];
#pragma unroll
            for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                for (int q = 0; q < 3; ++q) a[ii][q] = P.A1d[ii * 3 + q];
            const double pi0 = sm.pinv[0], pi1 = sm.pinv[1], pi2 = sm.pinv[2];
            double v[6][3];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int k = 32 * u + lane;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    double c = 0.0;
                    if (k <= i0 + q && P.m + k < nrows_prev) c = pvp[q] * col0[(size_t)q * P.ld + P.m + k];
                    v[u][q] = k <= tend ? (q == 0 ? pi0 : q == 1 ? pi1 : pi2) * c : 0.0;
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int k = 32 * u + lane;
                if (k <= tend) {
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii) {
                        double acc = 0.0;
                        acc = fma(a[ii][0], v[u][0], acc);
                        acc = fma(a[ii][1], v[u][1], acc);
                        acc = fma(a[ii][2], v[u][2], acc);
                        col0[(size_t)ii * P.ld + k] = acc;
                    }
                }
            }
        } else
        for (int k0 = 0; k0 <= tend; k0 += 64) {
            double v[2][kMaxN];
            if (pv_prev) {
                // in place: rows m + k of the block's own columns -> rows k (every load of a round before its first store;
                // a later round reads rows >= m + k0 + 64, which no earlier round has written)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int k = k0 + 32 * u + lane;
#pragma unroll
                    for (int s = 0; s < kMaxN; ++s) {
                        double c = 0.0;
                        if (s < n && k <= i0 + s && P.m + k < nrows_prev) c = pvp[s] * col0[(size_t)s * P.ld + P.m + k];
                        v[u][s] = (s < n && k <= tend) ? sm.pinv[s] * c : 0.0;
                    }
                }
                __syncwarp();
            } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = k0 + 32 * u + lane;
#pragma unroll
                for (int s = 0; s < kMaxN; ++s) v[u][s] = (s < n && k <= tend) ? sm.pinv[s] * PNMOL_STATE_LOAD(src + (size_t)s * D + k) : 0.0;
            }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = k0 + 32 * u + lane;
                if (k <= tend) {
                    for (int ii = 0; ii < n; ++ii) {
                        double acc = 0.0;
#pragma unroll
                        for (int s = 0; s < kMaxN; ++s)
                            if (s < n) acc = fma(P.A1d[ii * n + s], v[u][s], acc);
                        col0[(size_t)ii * P.ld + k] = acc;
                    }
                }
            }
        }
        // Ql^T columns i0 .. i0 + n - 1 = rows of Ql, entries 0 .. i  (lanes over the n x n blocks)
        if (i0 < nd) {
            for (int kb = lane; kb <= blk; kb += 32) {
                const double lk = ps * P.Lk[(size_t)blk * P.d + kb];
                for (int ii = 0; ii < n; ++ii) {
                    double* col = col0 + (size_t)ii * P.ld + D;
                    for (int kk = 0; kk < n; ++kk) {
                        const int k = kb * n + kk;
                        if (k <= i0 + ii) col[k] = lk * P.LQ1d[ii * n + kk];
                    }
                }
            }
        } else {
            const int comp = (blk - P.d) / P.npts;
            const double ds = P.diffscale ? P.diffscale[(size_t)b * P.ncomp + comp] : 1.0;
            const double eb = ds * P.Ediag[blk - P.d];
            for (int kb = lane; kb <= blk; kb += 32) {
                for (int ii = 0; ii < n; ++ii) {
                    double* col = col0 + (size_t)ii * P.ld + D;
                    for (int kk = 0; kk < n; ++kk) {
                        const int k = kb * n + kk;
                        if (k <= i0 + ii) col[k] = kb == blk ? eb * P.LQ1d[ii * n + kk] : 0.0;
                    }
                }
            }
        }
    }
    T::sync();
}

// ---------------------------------------------------------------- error estimate
// white.py:153-162 in closed form: with H = At E0 + p1 It E1 (At = order-0 entries of H,
// It = [I_d; 0]) and Ql Ql^T = K (x) q,  H Q H^T = q00 At K At^T + q01 p1 (At K It^T + It K At^T)
// + q11 p1^2 It K It^T.  sigma^2 = z^T S^-1 z / m through a Cholesky factorisation of S.
// (E E^T)[r][rp] for the step's measurement noise E = blockdiag(E_sqrtm, R_sqrtm)   white.py:157,184
__device__ __forceinline__ double meas_cov_entry(const Problem& P, int b, EMode emode, int r, int rp) {
    double ee = 0.0;
    if (emode == E_STEP_WHITE) {
        if (r < P.d) {
            if (rp == r) {
                const int comp = r / P.npts;
                const double ds = P.diffscale ? P.diffscale[(size_t)b * P.ncomp + comp] : 1.0;
                const double e = ds * P.Ediag[r];
                ee = e * e;
            }
        } else if (rp >= P.d) {
            for (int k = 0; k < P.nb; ++k) ee = fma(P.Rsq[(size_t)(r - P.d) * P.nb + k], P.Rsq[(size_t)(rp - P.d) * P.nb + k], ee);
        }
    }
    return ee;
}

template <class T = CtaTeam>
__device__ void error_estimate_global(const Problem& P, int b, const Smem& sm, double p1s, double dt, EMode emode,
                                      double nugget, const int32_t* Hcol, const double* Hval, double* F, double* S,
                                      double* err_out) {
    const int tid = T::tid(), lane = tid & 31, warp = T::warp();
    const int n = P.n, d = P.d, m = P.m;
    const double ps = P.priorscale ? P.priorscale[b] : 1.0;
    const double ps2 = ps * ps;
    double q00 = 0, q01 = 0, q11 = 0;
    for (int s = 0; s < n; ++s) {
        q00 = fma(P.LQ1d[s], P.LQ1d[s], q00);
        q01 = fma(P.LQ1d[s], P.LQ1d[n + s], q01);
        q11 = fma(P.LQ1d[n + s], P.LQ1d[n + s], q11);
    }
    // F = At K  (m x d)
    for (int r = warp; r < m; r += T::nwarps) {
        for (int k = lane; k < d; k += 32) {
            double acc = 0.0;
            for (int w = 0; w < P.wh; ++w) {
                const int c = Hcol[(size_t)r * P.wh + w];
                if (c >= 0 && (c % n) == 0 && c < n * d) acc = fma(Hval[(size_t)r * P.wh + w], ps2 * P.Kg[(size_t)(c / n) * d + k], acc);
            }
            F[(size_t)r * d + k] = acc;
        }
    }
    T::sync();
    // S (m x m), row-major, full
    for (int r = warp; r < m; r += T::nwarps) {
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
    T::sync();
    for (int r = tid; r < m; r += T::size) sm.y[r] = S[(size_t)r * m + r];  // diag(S) before factorisation
    T::sync();
    // right-looking Cholesky on the lower triangle (row-major)
    for (int k = 0; k < m; ++k) {
        const double piv = sqrt(S[(size_t)k * m + k]);
        T::sync();
        for (int r = k + tid; r < m; r += T::size) S[(size_t)r * m + k] = r == k ? piv : S[(size_t)r * m + k] / piv;
        T::sync();
        for (int r = k + 1 + warp; r < m; r += T::nwarps) {
            const double lrk = S[(size_t)r * m + k];
            for (int c = k + 1 + lane; c <= r; c += 32) S[(size_t)r * m + c] = fma(-lrk, S[(size_t)c * m + k], S[(size_t)r * m + c]);
        }
        T::sync();
    }
    // forward solve L u = z  (xw <- u), row-oriented dot form
    for (int r = tid; r < m; r += T::size) sm.xw[r] = sm.z[r];
    T::sync();
    for (int k = 0; k < m; ++k) {
        if (warp == 0) {
            double acc = 0.0;
            for (int c = lane; c < k; c += 32) acc = fma(S[(size_t)k * m + c], sm.xw[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) sm.xw[k] = (sm.xw[k] - acc) / S[(size_t)k * m + k];
        }
        T::sync();
    }
    double part = 0.0;
    for (int r = tid; r < m; r += T::size) part = fma(sm.xw[r], sm.xw[r], part);
    const double sigma = sqrt(T::sum(part, sm.red) / m);
    if (err_out)
        for (int i = tid; i < d; i += T::size) err_out[i] = dt * (sqrt(sm.y[i]) * sigma);
    T::sync();
}

__device__ __forceinline__ void error_estimate_solve_warp(const Problem& P, const Smem& sm, bool inverse);
// room for S and X = L^-1 side by side in the idle panel buffers (2 x 16 x vld doubles)
__device__ __forceinline__ bool err_inverse_fits(const Problem& P) { return 2 * P.m * P.ldm <= 32 * P.vld; }

// Shared-memory variant for small m (P.ldm > 0): S is assembled entry by entry straight from the sparse rows of
// At and the L2-resident Gram matrix (all loads independent), factorised and solved in sm.msq.
template <class T = CtaTeam>
__device__ void error_estimate_smem(const Problem& P, int b, const Smem& sm, double p1s, double dt, EMode emode,
                                    const int32_t* Hcol, const double* Hval, double* err_out, double* Sg = nullptr,
                                    double* Fg = nullptr, double* key = nullptr, bool defer_solve = false,
                                    bool use_inverse = false) {
    const int tid = T::tid(), lane = tid & 31, warp = T::warp();
    const int n = P.n, d = P.d, m = P.m, ldm = P.ldm;
    // Loop-invariant factor: for a linear PDE the rows of H depend on the member and on the Nordsieck scaling (dt) only,
    // so S = H Q H^T + E E^T and its Cholesky factor are the same in every step of a constant-step run.  The CTA keeps
    // the factor of its current member in its global scratch (Sg: lower triangle + diagonal of L, Fg: diag(S)) and
    // re-reads it while the key (member, scalings, noise mode) in shared memory is unchanged -- same numbers, bit for bit.
    // use_inverse (persistent constant-step loop): the cache holds X = L^-1 instead of L, so that every later step applies
    // it as a matrix-vector product instead of a forward substitution (52 dependent steps on one warp)
    const bool cacheable = key != nullptr && Sg != nullptr && Fg != nullptr && !P.semilinear;
    const bool inverse = cacheable && use_inverse && err_inverse_fits(P);
    const bool hit = cacheable && key[0] == (double)b && key[1] == sm.pv[0] && key[2] == p1s &&
                     key[3] == (double)(emode + (inverse ? 16 : 0));
    if (hit) {
        double* S = sm.msq;
        for (int idx = tid; idx < m * m; idx += T::size) {
            const int r = idx / m, c = idx - r * m;
            if (c < r) S[r * ldm + c] = Sg[idx];
            else if (c == r) sm.xw[r] = Sg[idx];
        }
        for (int r = tid; r < m; r += T::size) sm.y[r] = Fg[r];
    } else {
    const double ps = P.priorscale ? P.priorscale[b] : 1.0;
    const double ps2 = ps * ps;
    double q00 = 0, q01 = 0, q11 = 0;
    for (int s = 0; s < n; ++s) {
        q00 = fma(P.LQ1d[s], P.LQ1d[s], q00);
        q01 = fma(P.LQ1d[s], P.LQ1d[n + s], q01);
        q11 = fma(P.LQ1d[n + s], P.LQ1d[n + s], q11);
    }
    double* S = sm.msq;
    const int na_ode = P.wl + (P.semilinear ? P.ncomp : 0);  // order-0 entries come first in a row of H
    const int ntri = m * (m + 1) / 2;
    // mesh-point index c / n of every order-0 entry, once (the entry loops below are then division-free); -1 = padding
    int32_t* pt = sm.Hpt;
    const bool havept = pt != nullptr;
    if (havept) {
        for (int e = tid; e < m * P.wh; e += T::size) {
            const int c = Hcol[e];
            pt[e] = c >= 0 ? c / n : -1;
        }
        T::sync();
    }
    for (int idx = tid; idx < ntri; idx += T::size) {
        int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (r * (r + 1) / 2 > idx) --r;
        while ((r + 1) * (r + 2) / 2 <= idx) ++r;
        const int rp = idx - r * (r + 1) / 2;  // rp <= r
        const int32_t* hr = (havept ? pt : Hcol) + r * P.wh;
        const int32_t* hp = (havept ? pt : Hcol) + rp * P.wh;
        const double* vr = Hval + (size_t)r * P.wh;
        const double* vp = Hval + (size_t)rp * P.wh;
        const int nar = r < d ? na_ode : P.wb, nap = rp < d ? na_ode : P.wb;
        double acc = 0.0, cross = 0.0;
        for (int u = 0; u < nar; ++u) {
            int cu = hr[u];
            if (cu < 0) continue;
            if (!havept) cu /= n;
            const double* krow = P.Kg + (size_t)cu * d;
            double inner = 0.0;
            for (int w = 0; w < nap; ++w) {
                int cw = hp[w];
                if (cw < 0) continue;
                if (!havept) cw /= n;
                inner = fma(vp[w], krow[cw], inner);
            }
            acc = fma(vr[u], inner, acc);
            if (rp < d) cross = fma(vr[u], krow[rp], cross);
        }
        if (r < d) {
            for (int w = 0; w < nap; ++w) {
                int cw = hp[w];
                if (cw < 0) continue;
                if (!havept) cw /= n;
                cross = fma(vp[w], P.Kg[(size_t)cw * d + r], cross);
            }
        }
        double val = q00 * acc + q01 * p1s * cross;
        if (r < d && rp < d) val = fma(q11 * p1s * p1s, P.Kg[(size_t)r * d + rp], val);
        val = ps2 * val + meas_cov_entry(P, b, emode, r, rp);
        S[r * ldm + rp] = val;
        if (r == rp) sm.y[r] = val;  // diag(S) before factorisation
    }
    T::sync();
    // right-looking Cholesky of the lower triangle, two barriers per column; the diagonal of L goes to sm.xw so
    // that S[k][k] is never written while other threads may still read it
    for (int k = 0; k < m; ++k) {
        T::sync();
        const double skk = S[k * ldm + k];
        const double rinv = rsqrt(skk);
        for (int r = k + 1 + tid; r < m; r += T::size) S[r * ldm + k] *= rinv;
        if (tid == 0) sm.xw[k] = skk * rinv;
        T::sync();
        for (int r = k + 1 + warp; r < m; r += T::nwarps) {  // (warp per row, lanes over the columns: no divisions)
            const double lrk = S[r * ldm + k];
            for (int c = k + 1 + lane; c <= r; c += 32) S[r * ldm + c] = fma(-lrk, S[c * ldm + k], S[r * ldm + c]);
        }
    }
    if (inverse) {
        // X = L^-1, one column per thread by forward substitution into a second buffer (behind S in the idle panel
        // buffers), then back into the layout of L: strictly lower part in S, diagonal in sm.xw
        T::sync();
        double* X = S + (size_t)m * ldm;
        for (int c = tid; c < m; c += T::size) {
            X[c * ldm + c] = 1.0 / sm.xw[c];
            for (int r = c + 1; r < m; ++r) {
                double a0 = 0.0, a1 = 0.0;
                int j = c;
                for (; j + 1 < r; j += 2) { a0 = fma(S[r * ldm + j], X[j * ldm + c], a0); a1 = fma(S[r * ldm + j + 1], X[(j + 1) * ldm + c], a1); }
                if (j < r) a0 = fma(S[r * ldm + j], X[j * ldm + c], a0);
                X[r * ldm + c] = -(a0 + a1) / sm.xw[r];
            }
        }
        T::sync();
        for (int idx = tid; idx < m * m; idx += T::size) {
            const int r = idx / m, c = idx - r * m;
            if (c < r) S[r * ldm + c] = X[r * ldm + c];
            else if (c == r) sm.xw[r] = X[r * ldm + r];
        }
    }
    if (cacheable) {
        T::sync();
        for (int idx = tid; idx < m * m; idx += T::size) {
            const int r = idx / m, c = idx - r * m;
            if (c < r) Sg[idx] = S[r * ldm + c];
            else if (c == r) Sg[idx] = sm.xw[r];
        }
        for (int r = tid; r < m; r += T::size) Fg[r] = sm.y[r];
        if (tid == 0) { key[0] = (double)b; key[1] = sm.pv[0]; key[2] = p1s; key[3] = (double)(emode + (inverse ? 16 : 0)); }
    }
    }  // (factor computed or re-read)
    double* S = sm.msq;
    T::sync();
    // (defer_solve: the caller runs error_estimate_solve_warp on one warp next to other work and writes err_out itself)
    if (defer_solve) return;
    // forward solve L u = z by warp 0 (column oriented: each lane owns rows lane, lane + 32, lane + 64)
    if (warp == 0) error_estimate_solve_warp(P, sm, inverse);
    T::sync();
    const double sigma = sm.red[15];
    if (err_out)
        for (int i = tid; i < d; i += T::size) err_out[i] = dt * (sqrt(sm.y[i]) * sigma);
    T::sync();
}

// The forward solve of error_estimate_smem on ONE warp (S, the diagonal of L in sm.xw and z are in shared memory):
// sigma = sqrt(|L^-1 z|^2 / m) -> sm.red[15].  Same operations as the solve inside error_estimate_smem.
// inverse: S / sm.xw hold X = L^-1 (strictly lower part / diagonal): u = X z, one or two rows per lane.
__device__ __forceinline__ void error_estimate_solve_warp(const Problem& P, const Smem& sm, bool inverse) {
    const int lane = threadIdx.x & 31;
    const int m = P.m, ldm = P.ldm;
    const double* S = sm.msq;
    if (inverse) {
        double part = 0.0;
        for (int r = lane; r < m; r += 32) {
            const double* row = S + r * ldm;
            double a0 = sm.xw[r] * sm.z[r], a1 = 0.0;
            int c = 0;
            for (; c + 1 < r; c += 2) { a0 = fma(row[c], sm.z[c], a0); a1 = fma(row[c + 1], sm.z[c + 1], a1); }
            if (c < r) a0 = fma(row[c], sm.z[c], a0);
            const double ur = a0 + a1;
            part = fma(ur, ur, part);
        }
        part = warp_sum(part);
        if (lane == 0) sm.red[15] = sqrt(part / m);
        return;
    }
    // (the reciprocals of the diagonal are formed up front, one per lane and chunk: a division inside the dependent
    // chain of the substitution costs more than everything else in a step of it)
    double u[3], acc[3], rd[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int r = lane + 32 * q;
        u[q] = r < m ? sm.z[r] : 0.0; acc[q] = 0.0;
        rd[q] = r < m ? 1.0 / sm.xw[r] : 0.0;
    }
    for (int k = 0; k < m; ++k) {
        const int q = k >> 5;
        const double cand = q == 0 ? (u[0] - acc[0]) * rd[0] : q == 1 ? (u[1] - acc[1]) * rd[1] : (u[2] - acc[2]) * rd[2];
        const double uk = __shfl_sync(0xffffffffu, cand, k & 31);
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
            const int r = lane + 32 * qq;
            if (r > k && r < m) acc[qq] = fma(S[r * ldm + k], uk, acc[qq]);
            if (r == k) u[qq] = uk;
        }
    }
    double part = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int r = lane + 32 * q; if (r < m) part = fma(u[q], u[q], part); }
    part = warp_sum(part);
    if (lane == 0) sm.red[15] = sqrt(part / m);
}

template <class T = CtaTeam>
__device__ void error_estimate(const Problem& P, int b, const Smem& sm, double p1s, double dt, EMode emode, double nugget,
                               const int32_t* Hcol, const double* Hval, double* F, double* S, double* err_out,
                               bool defer_solve = false, bool use_inverse = false) {
    if (P.ldm > 0 && P.m <= 96)
        error_estimate_smem<T>(P, b, sm, p1s, dt, emode, Hcol, Hval, err_out, S, F, T::size == kThreads ? sm.ekey : nullptr,
                               defer_solve, use_inverse);
    else
        error_estimate_global<T>(P, b, sm, p1s, dt, emode, nugget, Hcol, Hval, F, S, err_out);
}

// ---------------------------------------------------------------- update stage
// E^T entry (row rp of column r of the bottom-left block), see EMode.
__device__ __forceinline__ double meas_sqrtm_entry(const Problem& P, int b, EMode emode, double nugget, int r,
                                                   int rp) {
    double v = 0.0;
    if (emode == E_STEP_WHITE || emode == E_STEP_PLUS_NUGGET) {
        if (r < P.d) {
            if (rp == r) {
                const int comp = r / P.npts;
                const double ds = P.diffscale ? P.diffscale[(size_t)b * P.ncomp + comp] : 1.0;
                v = ds * P.Ediag[r];
            }
        } else if (rp >= P.d) {
            v = P.Rsq[(size_t)(r - P.d) * P.nb + (rp - P.d)];
        }
    }
    if ((emode == E_NUGGET_ONLY || emode == E_STEP_PLUS_NUGGET) && rp == r) v += nugget;
    return v;
}

// Assemble the update block matrix of sqrt.py:60-65 / 82-87 in W (compact form: the
// D - m structurally zero rows of the reference's 2D x (m+D) matrix are dropped), factor
// it, and apply the result to the predicted mean.  On entry the top-right block holds
// R = Clp^T in place (Rsrc == nullptr) or is read from the row-major lower-triangular
// factor Rsrc (D x D).  mcur = number of measurement rows of this update.
struct UpdateOut {
    double* mean_out;  // (n, dd) or nullptr (then the new flat mean is left in sm.mp)
    double* chol_out;  // (D, D) row-major
    double* diff_out;  // scalar or nullptr
    double* ref_out;   // (d) or nullptr
    bool scale_by_p;   // multiply by the Nordsieck preconditioner on output
    // Deferred error estimate (error_estimate(..., defer_solve = true) has left S in shared memory): its forward solve
    // runs on warp 0 while the other warps assemble the left block of the update matrix.
    bool err_solve = false;
    bool err_inverse = false;   // the cached factor is X = L^-1 (error_estimate_smem, use_inverse)
    double err_dt = 0.0;
    double* err_out = nullptr;
};

// ---- part 1a: right block (rows 0..k hold R, copied from Rsrc if given; the rest of the envelope is zero)
static __device__ void update_build_right(const Problem& P, int mcur, int nrows, const double* __restrict__ Rsrc,
                                   const int32_t* te, const int32_t* be, double* Wr, int w0, int nw) {
    const int lane = threadIdx.x & 31;
    const int D = P.D, ld = P.ld;
    for (int k = w0; k < D; k += nw) {
        double* col = Wr + (size_t)k * ld;
        int tend = te ? te[mcur + k] : D - 1;
        if (tend > D - 1) tend = D - 1;
        if (Rsrc) {
            for (int i = lane; i <= k; i += 32) col[i] = Rsrc[(size_t)k * D + i];
        }
        for (int i = k + 1 + lane; i <= tend; i += 32) col[i] = 0.0;
        int bend = be ? be[mcur + k] : nrows - 1;
        if (bend > nrows - 1) bend = nrows - 1;
        for (int i = D + lane; i <= bend; i += 32) col[i] = 0.0;
    }
}

// ---- part 1b: left block: top = R H^T (column r = sum over the sparse row r of H), bottom = E^T
static __device__ void update_build_left(const Problem& P, int b, int mcur, int nrows, EMode emode, double nugget,
                                  const int32_t* te, const int32_t* be, const int32_t* Hcol, const double* Hval,
                                  double* Wl, const double* Wr, int w0, int nw) {
    const int lane = threadIdx.x & 31;
    const int D = P.D, ld = P.ld;
    for (int r = w0; r < mcur; r += nw) {
        double* col = Wl + (size_t)r * ld;
        int tend = te ? te[r] : D - 1;
        if (tend > D - 1) tend = D - 1;
        const int32_t* hc = Hcol + (size_t)r * P.wh;
        const double* hv = Hval + (size_t)r * P.wh;
        for (int i0 = 0; i0 <= tend; i0 += 128) {  // four lane-chunks per round: 4 wh independent loads in flight
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int w = 0; w < P.wh; ++w) {
                const int c = hc[w];
                if (c < 0) continue;
                const double hvw = hv[w];
                const double* rc = Wr + (size_t)c * ld;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 32 * u + lane;
                    if (i <= c && i <= tend) acc[u] = fma(hvw, rc[i], acc[u]);  // R[i][c], zero for i > c
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 32 * u + lane;
                if (i <= tend) col[i] = acc[u];
            }
        }
        int bend = be ? be[r] : nrows - 1;
        if (bend > nrows - 1) bend = nrows - 1;
        for (int i = D + lane; i <= bend; i += 32) col[i] = meas_sqrtm_entry(P, b, emode, nugget, r, i - D);
    }
}

// ---- blocked triangular solves on one CTA (large m: the multi-CTA path), vector in shared memory.
// Blocks of kTriB unknowns: the off-diagonal part is a matrix-vector product spread over all warps (independent,
// coalesced loads), the kTriB x kTriB diagonal block is staged in shared memory (leading dimension kTriB + 1) and solved
// by one warp with register-resident right-hand sides -- instead of one block barrier and one L2 round trip per unknown.
constexpr int kTriB = 64;
constexpr int kTriScratch = kTriB * (kTriB + 1) + kTriB;

// Lower-triangular system  L u = v  where row k of L is T + k rs (entries 0..k contiguous; diag != nullptr overrides
// the diagonal).  v: right-hand side on entry, solution on exit.
static __device__ void tri_solve_lower_rows(const double* __restrict__ T, size_t rs, const double* __restrict__ diag, int m,
                                            double* __restrict__ v, double* __restrict__ scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Bs = scratch;
    double* part = scratch + kTriB * (kTriB + 1);
    for (int jb = 0; jb < m; jb += kTriB) {
        const int nb = m - jb < kTriB ? m - jb : kTriB;
        for (int r = warp; r < nb; r += kWarps) {  // part[r] = sum_{c < jb} L[jb + r][c] u[c]
            const double* row = T + (size_t)(jb + r) * rs;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int c = lane;
            for (; c + 96 < jb; c += 128) {
                a0 = fma(row[c], v[c], a0); a1 = fma(row[c + 32], v[c + 32], a1);
                a2 = fma(row[c + 64], v[c + 64], a2); a3 = fma(row[c + 96], v[c + 96], a3);
            }
            for (; c < jb; c += 32) a0 = fma(row[c], v[c], a0);
            const double sum = warp_sum((a0 + a1) + (a2 + a3));
            if (lane == 0) part[r] = sum;
        }
        for (int idx = tid; idx < nb * nb; idx += kThreads) {
            const int r = idx / nb, c = idx - r * nb;
            if (c <= r) Bs[r * (kTriB + 1) + c] = (c == r && diag) ? diag[jb + r] : T[(size_t)(jb + r) * rs + jb + c];
        }
        __syncthreads();
        if (warp == 0) {
            const int r0 = lane, r1 = lane + 32;
            double rhs0 = r0 < nb ? v[jb + r0] - part[r0] : 0.0, rhs1 = r1 < nb ? v[jb + r1] - part[r1] : 0.0;
            for (int k = 0; k < nb; ++k) {
                const double cand = (k < 32 ? rhs0 : rhs1) / Bs[k * (kTriB + 1) + k];
                const double uk = __shfl_sync(0xffffffffu, cand, k & 31);
                if (r0 > k && r0 < nb) rhs0 = fma(-Bs[r0 * (kTriB + 1) + k], uk, rhs0);
                if (r1 > k && r1 < nb) rhs1 = fma(-Bs[r1 * (kTriB + 1) + k], uk, rhs1);
                if (lane == (k & 31)) { if (k < 32) rhs0 = uk; else rhs1 = uk; }
            }
            if (r0 < nb) v[jb + r0] = rhs0;
            if (r1 < nb) v[jb + r1] = rhs1;
        }
        __syncthreads();
    }
}

// Upper-triangular system  R x = v  where column k of R is T + k cs (entries 0..k contiguous).
static __device__ void tri_solve_upper_cols(const double* __restrict__ T, size_t cs, int m, double* __restrict__ v,
                                            double* __restrict__ scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Bs = scratch;
    for (int jb = ((m - 1) / kTriB) * kTriB; jb >= 0; jb -= kTriB) {
        const int nb = m - jb < kTriB ? m - jb : kTriB;
        for (int idx = tid; idx < nb * nb; idx += kThreads) {
            const int c = idx / nb, r = idx - c * nb;
            if (r <= c) Bs[r * (kTriB + 1) + c] = T[(size_t)(jb + c) * cs + jb + r];
        }
        __syncthreads();
        if (warp == 0) {
            const int r0 = lane, r1 = lane + 32;
            double rhs0 = r0 < nb ? v[jb + r0] : 0.0, rhs1 = r1 < nb ? v[jb + r1] : 0.0;
            for (int k = nb - 1; k >= 0; --k) {
                const double cand = (k < 32 ? rhs0 : rhs1) / Bs[k * (kTriB + 1) + k];
                const double xk = __shfl_sync(0xffffffffu, cand, k & 31);
                if (r0 < k) rhs0 = fma(-Bs[r0 * (kTriB + 1) + k], xk, rhs0);
                if (r1 < k) rhs1 = fma(-Bs[r1 * (kTriB + 1) + k], xk, rhs1);
                if (lane == (k & 31)) { if (k < 32) rhs0 = xk; else rhs1 = xk; }
            }
            if (r0 < nb) v[jb + r0] = rhs0;
            if (r1 < nb) v[jb + r1] = rhs1;
        }
        __syncthreads();
        for (int c = tid; c < jb; c += kThreads) {  // v[c] -= sum_k R[c][jb + k] x[jb + k]
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = 0;
            for (; k + 3 < nb; k += 4) {
                a0 = fma(T[(size_t)(jb + k) * cs + c], v[jb + k], a0);
                a1 = fma(T[(size_t)(jb + k + 1) * cs + c], v[jb + k + 1], a1);
                a2 = fma(T[(size_t)(jb + k + 2) * cs + c], v[jb + k + 2], a2);
                a3 = fma(T[(size_t)(jb + k + 3) * cs + c], v[jb + k + 3], a3);
            }
            for (; k < nb; ++k) a0 = fma(T[(size_t)(jb + k) * cs + c], v[jb + k], a0);
            v[c] -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
}

// ---- part 2 (one CTA): R1 = Wl[0:m, 0:m] (upper, column-major).  y = R1^-T z (for the mean),  x = R1^-1 z (quirk Q1,
// white.py:125 / latent.py:204);  m_new = mp - R2^T y (white.py:123, sqrt.py:72) is left in sm.mp.  Returns the
// local diffusion x.x / m.
template <class T = CtaTeam>
__device__ double update_solve(const Problem& P, const Smem& sm, int mcur, const double* Wl, const double* Wr) {
    const int tid = T::tid(), lane = tid & 31, warp = T::warp();
    const int D = P.D, ld = P.ld;
    double diff;
    if (P.ldm > 0 && mcur <= 96) {
        // stage R1 row-major in shared memory (odd leading dimension: conflict-free rows and columns), then
        // warp 0 solves R1^T y = z and warp 1 solves R1 x = z concurrently, column oriented
        const int ldm = P.ldm;
        double* Rs = sm.msq;
        for (int idx = tid; idx < mcur * mcur; idx += T::size) {
            const int k = idx / mcur, c = idx - k * mcur;  // column k, row c
            if (c <= k) Rs[c * ldm + k] = Wl[(size_t)k * ld + c];
        }
        T::sync();
        if (warp == 0) {  // forward: y_k = (z_k - sum_{c<k} R1[c][k] y_c) / R1[k][k]   (reciprocals of the diagonal up front)
            double zz[3], acc[3], rd[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int j = lane + 32 * q;
                zz[q] = j < mcur ? sm.z[j] : 0.0; acc[q] = 0.0;
                rd[q] = j < mcur ? 1.0 / Rs[j * ldm + j] : 0.0;
            }
            for (int k = 0; k < mcur; ++k) {
                const int q = k >> 5;
                const double cand = q == 0 ? (zz[0] - acc[0]) * rd[0] : q == 1 ? (zz[1] - acc[1]) * rd[1] : (zz[2] - acc[2]) * rd[2];
                const double yk = __shfl_sync(0xffffffffu, cand, k & 31);
#pragma unroll
                for (int qq = 0; qq < 3; ++qq) {
                    const int j = lane + 32 * qq;
                    if (j > k && j < mcur) acc[qq] = fma(Rs[k * ldm + j], yk, acc[qq]);
                    if (j == k) sm.y[j] = yk;
                }
            }
        }
        if (warp == (T::nwarps > 1 ? 1 : 0)) {  // backward: x_k = (z_k - sum_{c>k} R1[k][c] x_c) / R1[k][k]
            double zz[3], rd[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int j = lane + 32 * q;
                zz[q] = j < mcur ? sm.z[j] : 0.0;
                rd[q] = j < mcur ? 1.0 / Rs[j * ldm + j] : 0.0;
            }
            double part = 0.0;
            for (int k = mcur - 1; k >= 0; --k) {
                const int q = k >> 5;
                const double cand = q == 0 ? zz[0] * rd[0] : q == 1 ? zz[1] * rd[1] : zz[2] * rd[2];
                const double xk = __shfl_sync(0xffffffffu, cand, k & 31);
                if (lane == 0) part = fma(xk, xk, part);
#pragma unroll
                for (int qq = 0; qq < 3; ++qq) {
                    const int c = lane + 32 * qq;
                    if (c < k) zz[qq] = fma(-Rs[c * ldm + k], xk, zz[qq]);
                }
            }
            if (lane == 0) sm.red[14] = part / mcur;
        }
        T::sync();
        diff = sm.red[14];
    } else if (sm.tri && T::size == kThreads) {
        // multi-CTA path (large m): blocked solves with the vectors in shared memory
        double* vy = sm.tri;
        double* vx = sm.tri + mcur;
        double* scratch = sm.tri + 2 * mcur;
        for (int r = tid; r < mcur; r += kThreads) { const double zr = sm.z[r]; vy[r] = zr; vx[r] = zr; }
        __syncthreads();
        tri_solve_lower_rows(Wl, (size_t)ld, nullptr, mcur, vy, scratch);   // R1^T y = z: row k of R1^T is column k of R1
        tri_solve_upper_cols(Wl, (size_t)ld, mcur, vx, scratch);            // R1 x = z
        double part = 0.0;
        for (int r = tid; r < mcur; r += kThreads) { sm.y[r] = vy[r]; part = fma(vx[r], vx[r], part); }
        diff = T::sum(part, sm.red) / mcur;
    } else {
        for (int r = tid; r < mcur; r += T::size) { sm.y[r] = sm.z[r]; sm.xw[r] = sm.z[r]; }
        T::sync();
        for (int k = 0; k < mcur; ++k) {  // forward: dot form over contiguous column k of R1
            if (warp == 0) {
                const double* ck = Wl + (size_t)k * ld;
                double acc = 0.0;
                for (int c = lane; c < k; c += 32) acc = fma(ck[c], sm.y[c], acc);
                acc = warp_sum(acc);
                if (lane == 0) sm.y[k] = (sm.y[k] - acc) / ck[k];
            }
            T::sync();
        }
        for (int k = mcur - 1; k >= 0; --k) {  // backward: axpy form over contiguous column k of R1
            const double* ck = Wl + (size_t)k * ld;
            const double xk = sm.xw[k] / ck[k];
            T::sync();
            for (int c = tid; c < k; c += T::size) sm.xw[c] = fma(-xk, ck[c], sm.xw[c]);
            if (tid == 0) sm.xw[k] = xk;
            T::sync();
        }
        double part = 0.0;
        for (int r = tid; r < mcur; r += T::size) part = fma(sm.xw[r], sm.xw[r], part);
        diff = T::sum(part, sm.red) / mcur;
    }
    // m_new = mp - R2^T y   (white.py:123, sqrt.py:72); four columns per warp and round (independent loads)
    for (int k0 = 4 * warp; k0 < D; k0 += 4 * T::nwarps) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = lane; i < mcur; i += 32) {
            const double yi = sm.y[i];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < D) acc[u] = fma(Wr[(size_t)(k0 + u) * ld + i], yi, acc[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        }
        if (lane < 4 && k0 + lane < D) sm.mp[k0 + lane] -= lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    }
    T::sync();
    return diff;
}

// ---- part 3a (one CTA): mean (n, dd) = P m_new reshaped (white.py:132-135).  Returns 1 on a non-finite value.
template <class T = CtaTeam>
__device__ int update_output_mean(const Problem& P, const Smem& sm, const UpdateOut& out, double diff) {
    const int n = P.n, D = P.D;
    int bad = 0;
    if (!(diff == diff) || isinf(diff)) bad = 1;
    for (int k = T::tid(); k < D; k += T::size) {
        const int j = k / n, i = k - j * n;
        const double v = out.scale_by_p ? sm.pv[i] * sm.mp[k] : sm.mp[k];
        if (out.mean_out) out.mean_out[(size_t)i * P.dd + j] = v;
        if (out.ref_out && i == 0 && j < P.d) out.ref_out[j] = fabs(v);
        if (!isfinite(v)) bad = 1;
    }
    return bad;
}

// ---- part 3b: factor = P R3^T (white.py:132): row r of the factor = column m + r of R, rows m..
// (R3[c][r] = W[(m + r) ld + m + c]).  Returns 1 on a non-finite value.
static __device__ int update_output_factor(const Problem& P, const Smem& sm, const UpdateOut& out, int mcur, int nrows,
                                    const double* Wr, int w0, int nw) {
    const int lane = threadIdx.x & 31;
    const int n = P.n, D = P.D, ld = P.ld;
    int bad = 0;
    for (int r = w0; r < D; r += nw) {
        const double* col = Wr + (size_t)r * ld + mcur;
        const double pr = out.scale_by_p ? sm.pv[r % n] : 1.0;
        double* orow = out.chol_out + (size_t)r * D;
        for (int c0 = 0; c0 < D; c0 += 128) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + 32 * u + lane;
                v[u] = (c <= r && mcur + c < nrows) ? pr * col[c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + 32 * u + lane;
                if (c < D) {
                    PNMOL_STATE_STORE(orow + c, v[u]);  // streaming by default: keep the workspaces, not the state, resident in L2
                    if (!isfinite(v[u])) bad = 1;
                }
            }
        }
    }
    return bad;
}

static __device__ void update_stage(const Problem& P, int b, const Smem& sm, int mcur, EMode emode, double nugget,
                             const double* __restrict__ Rsrc, const int32_t* te, const int32_t* be,
                             const int32_t* Hcol, const double* Hval, double* W, const UpdateOut out,
                             int* nonfinite, PhaseClock& pc) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = P.D, ld = P.ld;
    const int nbot = emode == E_NONE ? 0 : mcur;
    const int nrows = D + nbot;
    // The right block always starts at column P.m of the workspace (where the predict QR
    // left R); an update with fewer measurement rows (initialisation on y0) uses the
    // columns P.m - mcur .. P.m - 1 for its left block.
    double* Wl = W + (size_t)(P.m - mcur) * ld;
    double* Wr = W + (size_t)P.m * ld;
    const int ncols = mcur + D;

    update_build_right(P, mcur, nrows, Rsrc, te, be, Wr, warp, kWarps);
    __syncthreads();
    pc.mark(13);
    if (out.err_solve) {
        if (warp == 0) { error_estimate_solve_warp(P, sm, out.err_inverse); pc.mark(15); }
        else update_build_left(P, b, mcur, nrows, emode, nugget, te, be, Hcol, Hval, Wl, Wr, warp - 1, kWarps - 1);
        __syncthreads();
        if (out.err_out) {   // white.py:160-162 (sm.y = diag(S) is overwritten only by update_solve, after the QR)
            const double sigma = sm.red[15];
            for (int i = tid; i < P.d; i += kThreads) out.err_out[i] = out.err_dt * (sqrt(sm.y[i]) * sigma);
        }
    } else {
        update_build_left(P, b, mcur, nrows, emode, nugget, te, be, Hcol, Hval, Wl, Wr, warp, kWarps);
        __syncthreads();
    }

    pc.mark(4);
    Shape sh;
    sh.nt = D; sh.nbot = nbot; sh.ncols = ncols; sh.te = te; sh.be = be; sh.ldr = ld;
    householder_qr_fast(Wl, ld, sh, sm.fq, pc);
    pc.mark(5);

    const double diff = update_solve(P, sm, mcur, Wl, Wr);
    if (out.diff_out && tid == 0) *out.diff_out = diff;
    pc.mark(6);
    int bad = update_output_mean(P, sm, out, diff);
    if (out.chol_out) bad |= update_output_factor(P, sm, out, mcur, nrows, Wr, warp, kWarps);  // (nullptr: the next step takes R3 in place)
    if (bad) atomicOr(nonfinite, 1);
    __syncthreads();
    pc.mark(7);
}

}  // namespace pnmol
