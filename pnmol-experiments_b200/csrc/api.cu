// C ABI of libpnmol_b200.so (see include/pnmol_b200.h).  Host-side orchestration only:
// argument checking, device-buffer ownership, envelope computation, kernel launches.
#include "../../include/pnmol_b200.h"
#include "ek1_kernels.cuh"
#include "ek1_large.cuh"
#include "ek1_small.cuh"
#include "discretize.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace pnmol;

namespace {
constexpr unsigned kPaceRing = 16;   // pace-keeping counters per handle (one per launch in flight)

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                std::to_string(__LINE__) + ")");                                   \
    } while (0)

void compute_structure(int latent, int semilinear, int d, int n, int nb, int ncomp, const int32_t* Lcol, int wl,
                       const int32_t* Bcol, int wb, int dense, int32_t* te_p, int32_t* be_p, int32_t* te_u,
                       int32_t* be_u) {
    const int dd = latent ? 2 * d : d, D = n * dd, m = d + nb, npts = d / ncomp;
    for (int i = 0; i < D; ++i) {
        te_p[i] = dense ? D - 1 : std::min(D - 1, n * (i / n) + n - 1);
        be_p[i] = D + i;
    }
    int run = 0, brun = D - 1;
    for (int r = 0; r < m; ++r) {
        int last = 0;
        if (r < d) {
            for (int w = 0; w < wl; ++w) {
                const int c = Lcol[(size_t)r * wl + w];
                if (c >= 0) last = std::max(last, c * n);
            }
            if (semilinear) last = std::max(last, ((ncomp - 1) * npts + r % npts) * n);
            last = std::max(last, r * n + 1);
            if (latent) last = std::max(last, (d + r) * n);
        } else {
            for (int w = 0; w < wb; ++w) {
                const int c = Bcol[(size_t)(r - d) * wb + w];
                if (c >= 0) last = std::max(last, c * n);
            }
        }
        run = std::max(run, std::max(last, r));
        te_u[r] = std::min(D - 1, run);
        if (!latent) brun = std::max(brun, D + (r < d ? r : m - 1));
        be_u[r] = brun;
    }
    for (int k = 0; k < D; ++k) {
        te_u[m + k] = std::min(D - 1, std::max(te_u[m - 1], m + k));
        be_u[m + k] = latent ? D - 1 : D + m - 1;
    }
}

// Largest compact row list of any kNB-wide panel (mirrors panel_rows in qr_blocked.cuh; ldr > 0: tile-aligned lists).
int max_panel_len(int nt, int nbot, int ncols, const int32_t* te, const int32_t* be, int ldr = 0) {
    const int nrows = nt + nbot, nref = std::min(nrows, ncols);
    int best = 1;
    for (int j0 = 0; j0 < nref; j0 += kNB) {
        const int jl = std::min(j0 + kNB, nref) - 1;
        auto top = [&](int j) { return std::min(te ? te[j] : nt - 1, nt - 1); };
        auto bot = [&](int j) { return std::min(be ? be[j] : nrows - 1, nrows - 1); };
        RowMap rm;
        if (j0 < nt) {
            const int jt = std::min(jl, nt - 1);
            rm = make_row_map(nt, ldr, j0, std::max(top(jt), jt), bot(jl));
        } else {
            rm = make_row_map(nt, ldr, j0, 0, std::max(bot(jl), jl));
        }
        best = std::max(best, rm.len);
    }
    return best;
}

}  // namespace

struct pnmol_b200_handle {
    int kind, device, grid, num_sms;
    unsigned* gsync = nullptr;      // ring of pace-keeping counters (launch_run)
    unsigned gsync_next = 0;
    Problem P;
    std::vector<void*> allocs;
    size_t smem_bytes, smem_optin;
    bool have_op = false, have_prior = false;
    double* steparr = nullptr;  // device: dts | tnew | pv | pinv
    double* step_host = nullptr;  // pinned staging of the same
    cudaEvent_t step_ev = nullptr;
    int steparr_cap = 0;
    // host-path state (pnmol_b200_simulate_final_state_host)
    double *hs_y0 = nullptr, *hs_mean_a = nullptr, *hs_mean_b = nullptr, *hs_chol_a = nullptr, *hs_chol_b = nullptr,
           *hs_diffsum = nullptr, *hs_diffcal = nullptr;
    int32_t* hs_status = nullptr;
    cudaStream_t copy_stream = nullptr;               // device-to-host copies of finished chunks (host path)
    cudaEvent_t chunk_ev = nullptr, copy_ev = nullptr;
    double *hs_err = nullptr, *hs_ref = nullptr;  // scratch of pnmol_b200_run_adaptive
    // multi-CTA path for large state dimension (ek1_large.cuh): chosen when the single-CTA kernels' shared memory
    // does not fit (or PNMOL_B200_FORCE_LARGE=1, used by the parity tests to run both paths on the same inputs)
    bool large = false;
    LargeQR q{};
    size_t smem_large = 0;
    int cluster = 1;  // thread-block cluster size of the multi-CTA kernels (panel factorisation on cluster 0)
    // small-state path (ek1_small.cuh): one warp per member, workspace in the warp's shared-memory slice
    bool small = false;
    SmallGeom sgeo{};
    size_t smem_small = 0;
};

namespace {

template <typename T>
int dev_alloc(pnmol_b200_handle* h, T** out, size_t count) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    h->allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
}

template <typename T>
int dev_upload(pnmol_b200_handle* h, const T** out, const T* host, size_t count) {
    T* p = nullptr;
    int rc = dev_alloc(h, &p, count);
    if (rc) return rc;
    CU(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}

// 2-D tensor map of the reflector buffer of the multi-CTA QR (qr_large.cuh): dim0 = compact row index (contiguous, `lv`
// doubles per reflector), dim1 = reflector; one box = `box_rows` rows x kNB reflectors.  The driver entry point is
// resolved at run time (no link-time dependency on libcuda).
int make_reflector_tensor_map(CUtensorMap* out, double* base, int lv, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(-2, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)lv, (cuuint64_t)kNB};
    const cuuint64_t strides[1] = {(cuuint64_t)lv * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)kNB};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-2, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

int prop_smem_per_sm(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    return v;
}

int ensure_ready(pnmol_b200_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (!h->have_op) return fail(-1, "pnmol_b200_set_operator has not been called");
    if (!h->have_prior) return fail(-1, "pnmol_b200_set_prior has not been called");
    if (h->P.semilinear && !h->P.rparams) return fail(-1, "semi-linear solver needs reaction parameters (pnmol_b200_set_members)");
    CU(cudaSetDevice(h->device));
    return 0;
}

int upload_steps(pnmol_b200_handle* h, int nsteps, double t0, const double* dts, const double* pv, const double* pinv,
                 cudaStream_t st, RunArgs* a) {
    const int n = h->P.n;
    const size_t per = (size_t)nsteps * (2 + 2 * n);
    // Device array + a pinned staging buffer owned by the handle; the copy is asynchronous, and an event recorded after it
    // tells the next call when the staging buffer may be overwritten (no synchronisation of the caller's stream).
    if (h->steparr_cap < (int)per) {
        if (h->step_ev) CU(cudaEventSynchronize(h->step_ev));
        CU(cudaStreamSynchronize(st));
        if (h->steparr) CU(cudaFree(h->steparr));
        if (h->step_host) CU(cudaFreeHost(h->step_host));
        CU(cudaMalloc((void**)&h->steparr, per * sizeof(double)));
        CU(cudaMallocHost((void**)&h->step_host, per * sizeof(double)));
        h->steparr_cap = (int)per;
    }
    if (!h->step_ev) CU(cudaEventCreateWithFlags(&h->step_ev, cudaEventDisableTiming));
    else CU(cudaEventSynchronize(h->step_ev));  // the previous upload has left the staging buffer
    double* buf = h->step_host;
    double t = t0;
    for (int s = 0; s < nsteps; ++s) {
        buf[s] = dts[s];
        t = t + dts[s];  // same floating-point accumulation as pdefilter.py:140 / white.py:139
        buf[nsteps + s] = t;
    }
    std::memcpy(buf + 2 * (size_t)nsteps, pv, sizeof(double) * nsteps * n);
    std::memcpy(buf + 2 * (size_t)nsteps + (size_t)nsteps * n, pinv, sizeof(double) * nsteps * n);
    CU(cudaMemcpyAsync(h->steparr, buf, per * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(h->step_ev, st));
    a->dts = h->steparr;
    a->tnew = h->steparr + nsteps;
    a->pv = h->steparr + 2 * (size_t)nsteps;
    a->pinv = h->steparr + 2 * (size_t)nsteps + (size_t)nsteps * n;
    return 0;
}

// Cooperative launch of a multi-CTA kernel, as thread-block clusters when h->cluster > 1.
template <typename Kernel, typename Args>
int launch_large(pnmol_b200_handle* h, Kernel kernel, Args& a, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = h->smem_large;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = h->cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = h->cluster > 1 ? 2 : 1;
    CU(cudaLaunchKernelEx(&cfg, kernel, h->P, a, h->q));
    return 0;
}

int launch_run(pnmol_b200_handle* h, RunArgs& a, cudaStream_t st) {
    if (h->small) {
        k_run_small<<<h->grid, 32 * h->sgeo.nwarps, h->smem_small, st>>>(h->P, a, h->sgeo);
    } else if (h->large) {
        int rc = launch_large(h, k_run_large, a, st);
        if (rc) return rc;
    } else {
        CU(cudaMemsetAsync(h->P.smslot, 0, sizeof(int) * h->num_sms, st));
        // pace keeping (ek1_kernels.cuh): one counter per launch out of a ring, so that launches of the same handle on
        // different streams (the chunked host route) never share one; PNMOL_B200_PACE=0 switches it off
        const char* pace_env = std::getenv("PNMOL_B200_PACE");
        const bool pace = !pace_env || std::atoi(pace_env) != 0;
        Problem P = h->P;
        P.gsync = nullptr;
        if (pace && h->gsync && a.nsteps > 1 && h->grid > 1) {
            P.gsync = h->gsync + (h->gsync_next++ % kPaceRing);
            CU(cudaMemsetAsync(P.gsync, 0, sizeof(unsigned), st));
        }
        if (P.gsync) {
            // CTAs that wait for each other must all be resident: a cooperative launch guarantees it (two such launches
            // on different streams -- two handles on one device -- are then run one after the other instead of each
            // holding a part of the SMs and waiting for the rest forever)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(h->grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->smem_bytes; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative;
            at[0].val.cooperative = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, k_run, P, a));
        } else
        k_run<<<h->grid, kThreads, h->smem_bytes, st>>>(P, a);
    }
    ++g_launches;
    return 0;
}

}  // namespace

extern "C" {

const char* pnmol_b200_last_error(void) { return g_err.c_str(); }
int pnmol_b200_version(void) { return 100; }
int64_t pnmol_b200_launch_count(void) { return g_launches.load(); }

int pnmol_b200_structure(int kind, int d, int num_derivatives, int nb, const int32_t* L_col, int wl,
                         const int32_t* B_col, int wb, int ncomp, int dense_factor, int32_t* te_p, int32_t* be_p,
                         int32_t* te_u, int32_t* be_u) {
    if (kind < 0 || kind > 3 || d <= 0 || num_derivatives < 1 || nb < 0 || ncomp <= 0 || d % ncomp)
        return fail(-1, "pnmol_b200_structure: invalid shape arguments");
    compute_structure(kind >= 2, kind & 1, d, num_derivatives + 1, nb, ncomp, L_col, wl, B_col, wb, dense_factor, te_p,
                      be_p, te_u, be_u);
    return 0;
}

int pnmol_b200_create(pnmol_b200_handle** out, int kind, int d, int num_derivatives, int nb, int ncomp, int batch,
                      int reaction_id, int device) {
    if (!out) return fail(-1, "null output handle");
    if (kind < 0 || kind > 3) return fail(-1, "unknown solver kind");
    const int n = num_derivatives + 1;
    if (d <= 0 || batch <= 0 || nb < 0 || ncomp <= 0 || ncomp > kMaxComp || d % ncomp) return fail(-1, "invalid d / batch / nb / ncomp");
    if (n < 2 || n > kMaxN) return fail(-1, "num_derivatives must be in 1..7");
    const bool semil = kind & 1;
    if (semil && (reaction_id < 1 || reaction_id > 3)) return fail(-1, "semi-linear solver needs a device reaction id (1..3)");
    if (reaction_id == PNMOL_B200_REACTION_SPRUCE && ncomp != 1) return fail(-1, "spruce reaction has 1 component");
    if (reaction_id == PNMOL_B200_REACTION_SIR && ncomp != 3) return fail(-1, "SIR reaction has 3 components");
    if (reaction_id == PNMOL_B200_REACTION_LOTKA_VOLTERRA && ncomp != 2) return fail(-1, "Lotka-Volterra reaction has 2 components");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    auto* h = new pnmol_b200_handle();
    std::memset(&h->P, 0, sizeof(Problem));
    h->kind = kind;
    h->device = device;
    Problem& P = h->P;
    P.latent = kind >= 2;
    P.semilinear = semil;
    P.reaction = semil ? reaction_id : 0;
    P.d = d; P.n = n; P.nb = nb; P.ncomp = ncomp; P.npts = d / ncomp;
    P.dd = P.latent ? 2 * d : d;
    P.D = n * P.dd;
    P.m = d + nb;
    P.batch = batch;
    P.ld = 2 * P.D + 8;  // 8 spare rows per column: panel row lists are padded to whole 8-row tiles
    if (P.D < P.m) { delete h; return fail(-1, "update_sqrt needs D >= m (src/pnmol/base/sqrt.py:55-57)"); }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    {
        int rc = dev_alloc(h, &P.smslot, (size_t)h->num_sms);
        if (rc) { delete h; return rc; }
        rc = dev_alloc(h, &h->gsync, (size_t)kPaceRing);
        if (rc) { delete h; return rc; }
        P.gsync = nullptr;
    }
    *out = h;
    return 0;
}

int pnmol_b200_destroy(pnmol_b200_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (void* p : h->allocs) cudaFree(p);
    if (h->step_ev) { cudaEventSynchronize(h->step_ev); cudaEventDestroy(h->step_ev); }
    if (h->steparr) cudaFree(h->steparr);
    if (h->step_host) cudaFreeHost(h->step_host);
    for (void* p : {(void*)h->hs_y0, (void*)h->hs_mean_a, (void*)h->hs_mean_b, (void*)h->hs_chol_a, (void*)h->hs_chol_b,
                    (void*)h->hs_diffsum, (void*)h->hs_diffcal, (void*)h->hs_status, (void*)h->hs_err, (void*)h->hs_ref})
        if (p) cudaFree(p);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->chunk_ev) cudaEventDestroy(h->chunk_ev);
    if (h->copy_ev) cudaEventDestroy(h->copy_ev);
    delete h;
    return 0;
}

int pnmol_b200_set_operator(pnmol_b200_handle* h, const int32_t* L_col, const double* L_val, int wl, const double* E_diag,
                            const int32_t* B_col, const double* B_val, int wb, const double* R_sqrtm) {
    if (!h) return fail(-1, "null handle");
    if (h->have_op) return fail(-1, "operator already set (create a new handle)");
    if (!L_col || !L_val || !E_diag || wl <= 0) return fail(-1, "null operator arrays");
    Problem& P = h->P;
    if (P.nb > 0 && (!B_col || !B_val || !R_sqrtm || wb <= 0)) return fail(-1, "null boundary arrays");
    CU(cudaSetDevice(h->device));
    for (size_t k = 0; k < (size_t)P.d * wl; ++k)
        if (L_col[k] >= P.d) return fail(-1, "L_col out of range");
    for (size_t k = 0; k < (size_t)P.nb * wb; ++k)
        if (B_col[k] >= P.d) return fail(-1, "B_col out of range");
    P.wl = wl;
    P.wb = P.nb > 0 ? wb : 0;
    P.wh = std::max(wl + (P.semilinear ? P.ncomp : 0) + 1 + (P.latent ? 1 : 0), P.wb);
    int rc;
    if ((rc = dev_upload(h, &P.Lcol, L_col, (size_t)P.d * wl))) return rc;
    if ((rc = dev_upload(h, &P.Lval, L_val, (size_t)P.d * wl))) return rc;
    if ((rc = dev_upload(h, &P.Ediag, E_diag, (size_t)P.d))) return rc;
    if (P.nb > 0) {
        if ((rc = dev_upload(h, &P.Bcol, B_col, (size_t)P.nb * wb))) return rc;
        if ((rc = dev_upload(h, &P.Bval, B_val, (size_t)P.nb * wb))) return rc;
        if ((rc = dev_upload(h, &P.Rsq, R_sqrtm, (size_t)P.nb * P.nb))) return rc;
    }
    std::vector<int32_t> te_p(P.D), be_p(P.D), te_pd(P.D), te_u(P.m + P.D), be_u(P.m + P.D);
    compute_structure(P.latent, P.semilinear, P.d, P.n, P.nb, P.ncomp, L_col, wl, B_col, P.wb, 0, te_p.data(), be_p.data(),
                      te_u.data(), be_u.data());
    for (int i = 0; i < P.D; ++i) te_pd[i] = P.D - 1;
    if ((rc = dev_upload(h, &P.te_p, te_p.data(), te_p.size()))) return rc;
    if ((rc = dev_upload(h, &P.be_p, be_p.data(), be_p.size()))) return rc;
    if ((rc = dev_upload(h, &P.te_pd, te_pd.data(), te_pd.size()))) return rc;
    if ((rc = dev_upload(h, &P.te_u, te_u.data(), te_u.size()))) return rc;
    if ((rc = dev_upload(h, &P.be_u, be_u.data(), be_u.size()))) return rc;
    // panel geometry of the blocked QR: V buffer rows = 16 G for the largest row list (<= 512, else fallback)
    {
        int maxlen = max_panel_len(P.D, P.D, P.D, te_p.data(), be_p.data(), P.ld);
        maxlen = std::max(maxlen, max_panel_len(P.D, P.D, P.D, te_pd.data(), be_p.data(), P.ld));
        maxlen = std::max(maxlen, max_panel_len(P.D, P.latent ? 0 : P.m, P.m + P.D, te_u.data(), be_u.data(), P.ld));
        maxlen = std::max(maxlen, max_panel_len(P.D, P.d, P.d + P.D, nullptr, nullptr, P.ld));   // initialisation updates
        maxlen = std::max(maxlen, max_panel_len(P.D, P.m, P.m + P.D, nullptr, nullptr, P.ld));
        const int G = maxlen <= 4 * kRPL ? 4 : maxlen <= 8 * kRPL ? 8 : maxlen <= 16 * kRPL ? 16 : 32;
        P.vld = std::max(64, kRPL * G);  // rows of the panel buffers of the blocked QR (LP: 64, 128 or 256)
        P.ldm = P.m <= 96 ? (P.m | 1) : 0;  // odd leading dimension: conflict-free rows and columns
        // the sparse rows of H go to shared memory when two CTAs per SM still fit (else: global scratch, L2-resident)
        P.whs = P.wh;
        h->smem_bytes = smem_doubles(P.D, P.m, P.dd, P.vld, P.ldm, P.whs) * sizeof(double);
        if (2 * (h->smem_bytes + 1024) > (size_t)prop_smem_per_sm(h->device) && P.m * P.wh > 256) {
            const size_t without = smem_doubles(P.D, P.m, P.dd, P.vld, P.ldm, 0) * sizeof(double);
            if (2 * (without + 1024) <= (size_t)prop_smem_per_sm(h->device) || h->smem_bytes > h->smem_optin) {
                P.whs = 0;
                h->smem_bytes = without;
            }
        }
        const char* force = std::getenv("PNMOL_B200_FORCE_LARGE");
        // auto: the single-CTA kernels need their shared memory to fit and are only fast while every panel's row list
        // fits the register-resident panels (<= 32 * kRPL rows); beyond that the whole grid works on one member
        h->large = h->smem_bytes > h->smem_optin || maxlen > 32 * kRPL;
        if (force && h->smem_bytes <= h->smem_optin) h->large = std::atoi(force) != 0;
        const char* pathenv = std::getenv("PNMOL_B200_PATH");
        const std::string want_path = pathenv ? pathenv : "";
        if (want_path == "large") h->large = true;
        if (want_path == "cta" && h->smem_bytes <= h->smem_optin) h->large = false;
        // Small state dimension: one warp per member with the whole workspace in shared memory (ek1_small.cuh).
        // Auto-selected when at least 4 members fit an SM's shared memory side by side (D <~ 48).
        if (want_path == "small" || (want_path.empty() && !force)) {
            const int ld_small = (2 * P.D) | 1;  // odd leading dimension: conflict-free lane-per-column accesses
            const int ldm_small = P.m | 1;
            size_t per_warp = small_smem_doubles(P.D, P.m, P.dd, ld_small, ldm_small, P.wh);
            per_warp += per_warp & 1;
            const int fit = (int)(h->smem_optin / sizeof(double) / per_warp);
            if (P.m <= 96 && (fit >= 4 || (want_path == "small" && fit >= 1))) {
                h->small = true;
                h->large = false;
                P.ld = ld_small;
                P.ldm = ldm_small;
                SmallGeom& geo = h->sgeo;
                geo.per_warp = (int)per_warp;
                geo.nwarps = std::min(PNMOL_SMALL_THREADS / 32, fit);
                if (const char* e = std::getenv("PNMOL_B200_WARPS")) geo.nwarps = std::max(1, std::min(geo.nwarps, std::atoi(e)));  // tuning
                h->smem_small = (size_t)geo.per_warp * geo.nwarps * sizeof(double);
                CU(cudaFuncSetAttribute(k_run_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_small));
                CU(cudaFuncSetAttribute(k_init_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_small));
                int occ = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_run_small, 32 * geo.nwarps, h->smem_small));
                occ = std::max(1, occ);
                h->grid = std::min((P.batch + geo.nwarps - 1) / geo.nwarps, occ * h->num_sms);
                if (const char* e = std::getenv("PNMOL_B200_GRID")) h->grid = std::max(1, std::min(h->grid, std::atoi(e)));
                h->have_op = true;
                return 0;
            }
            if (want_path == "small") return fail(-1, "PNMOL_B200_PATH=small: the member's workspace does not fit in shared memory");
        }
        if (h->large) {
            // one member at a time on the whole grid: one workspace, vectors in global scratch
            LargeQR& q = h->q;
            // Panel buffer: half of the SM's shared memory when the problem allows it (two CTAs per SM: the trailing
            // phases are latency-bound and the cluster panel only needs a row slice per CTA), else all of it.
            const int lp = (maxlen + 7) & ~7;
            const int cap_full = ((int)(h->smem_optin / sizeof(double)) - kLargeFixed - 64) & ~15;
            const int cap_half = ((int)((h->smem_optin + 1024) / 2 / sizeof(double)) - 128 - kLargeFixed - 64) & ~15;
            auto cb_for = [&](int c) { return (size_t)P.m * 17 <= (size_t)c ? 16 : (size_t)P.m * 9 <= (size_t)c ? 8 : (size_t)P.m * 5 <= (size_t)c ? 4 : 0; };
            // (measured: two CTAs per SM speed the trailing phases up by 1.6x at C4 but slow the panel team down, whose
            // SMs then also host a CTA spinning in the grid barrier: a gain only where the trailing update dominates)
            bool two = PNMOL_LARGE_CTAS >= 2 && maxlen >= 2048 && lp <= cap_half && cb_for(cap_half) > 0;
            if (const char* e = std::getenv("PNMOL_B200_LARGE_CTAS")) two = std::atoi(e) >= 2 && lp <= cap_half && cb_for(cap_half) > 0;
            int cap = two ? cap_half : cap_full;
            if (const char* e = std::getenv("PNMOL_B200_LARGE_CAP")) cap = std::max(1280, std::min(cap, std::atoi(e)));  // >= one 64-row V chunk (64 x kLdr)
            cap &= ~15;
            if (lp > cap || cb_for(cap) == 0)
                return fail(-1, "state dimension too large: one panel column does not fit in shared memory");
            q.cap = cap;
            q.cb = cb_for(cap);
            q.lv = std::max(lp + 8, 256);
            q.ycols = P.m + P.D;
            // (the two TMA stages of the trailing phases live in the panel buffer: 2 x 16 x 242 doubles, 128-byte aligned)
            h->smem_large = (size_t)(kLargeFixed + std::max(cap, 2 * kTmaStage + 16)) * sizeof(double);
            P.ldm = 0;
            CU(cudaFuncSetAttribute(k_run_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_large));
            CU(cudaFuncSetAttribute(k_init_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_large));
            int occ = 0, occ2 = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_run_large, kThreads, h->smem_large));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_init_large, kThreads, h->smem_large));
            if (std::min(occ, occ2) < 1) return fail(-1, "multi-CTA kernels do not fit on an SM");
            h->grid = h->num_sms * std::min(std::min(occ, occ2), PNMOL_LARGE_CTAS);
            if (const char* e = std::getenv("PNMOL_B200_GRID")) h->grid = std::max(1, std::min(h->grid, std::atoi(e)));
            // thread-block clusters: the panel factorisation runs on cluster 0 (rows split over its CTAs)
            h->cluster = 1;
            // (measured on B200: clusters of 8 make the C4 panel 2.5x and the C2/C3 panels ~1.15x faster; the trailing
            // phases then run on the 120 CTAs that fit whole clusters instead of 148)
            int want_cluster = 8;
            if (const char* e = std::getenv("PNMOL_B200_CLUSTER")) want_cluster = std::atoi(e);
            if (want_cluster > 1) {
                for (auto fn : {(const void*)k_run_large, (const void*)k_init_large})
                    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(want_cluster); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->smem_large;
                cudaLaunchAttribute at[2];
                at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
                at[1].id = cudaLaunchAttributeClusterDimension;
                at[1].val.clusterDim.x = want_cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 2;
                int ncl = 0, ncl2 = 0;
                CU(cudaOccupancyMaxActiveClusters(&ncl, k_run_large, &cfg));
                CU(cudaOccupancyMaxActiveClusters(&ncl2, k_init_large, &cfg));
                ncl = std::min(ncl, ncl2);
                if (ncl >= 1) { h->cluster = want_cluster; h->grid = ncl * want_cluster; }
            }
            const size_t wsz = (size_t)P.ld * (P.m + P.D);
            if ((rc = dev_alloc(h, &P.W, wsz))) return rc;
            if ((rc = dev_alloc(h, &P.Hcol, (size_t)P.m * P.wh))) return rc;
            if ((rc = dev_alloc(h, &P.Hval, (size_t)P.m * P.wh))) return rc;
            if ((rc = dev_alloc(h, &P.F, (size_t)P.m * P.d))) return rc;
            if ((rc = dev_alloc(h, &P.S, (size_t)P.m * P.m))) return rc;
            if ((rc = dev_alloc(h, &q.Vg, (size_t)kNB * q.lv))) return rc;
            if ((rc = dev_alloc(h, &q.Tg, (size_t)kNB * kLdr))) return rc;
            for (int i = 0; i < kTmaBoxes; ++i)
                if ((rc = make_reflector_tensor_map(&q.tmapV[i], q.Vg, q.lv, large_box_pitch(i)))) return rc;
            if ((rc = dev_alloc(h, &q.Yp, (size_t)(maxlen / 64 + 2) * q.ycols * kNB))) return rc;
            if ((rc = dev_alloc(h, &q.vec, (size_t)P.D + 3 * (size_t)P.m + P.dd + 8))) return rc;
            if ((rc = dev_alloc(h, &q.Ld, (size_t)P.m))) return rc;
            if ((rc = dev_alloc(h, &q.nf, 4))) return rc;   // (two words: the adaptive loop alternates by attempt parity)
            CU(cudaMemset(P.W, 0, wsz * sizeof(double)));
            CU(cudaMemset(q.Vg, 0, (size_t)kNB * q.lv * sizeof(double)));
            h->have_op = true;
            return 0;
        }
    }
    // launch geometry + per-CTA scratch
    CU(cudaFuncSetAttribute(k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    CU(cudaFuncSetAttribute(k_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_run, kThreads, h->smem_bytes));
    int want = PNMOL_MIN_CTAS;
    if (const char* e = std::getenv("PNMOL_B200_CTAS_PER_SM")) want = std::max(1, std::atoi(e));
    occ = std::max(1, std::min(occ, want));
    if (std::getenv("PNMOL_B200_VERBOSE"))
        fprintf(stderr, "pnmol_b200: CTA-per-member path: D=%d m=%d wh=%d (H rows in %s), %zu bytes of dynamic shared memory per CTA, %d CTA(s) per SM\n",
                P.D, P.m, P.wh, P.whs ? "shared memory" : "global scratch", h->smem_bytes, occ);
    h->grid = std::min(P.batch, occ * h->num_sms);
    if (const char* e = std::getenv("PNMOL_B200_GRID")) h->grid = std::max(1, std::min(h->grid, std::atoi(e)));  // tuning
    const size_t wsz = (size_t)P.ld * (P.m + P.D);
    if ((rc = dev_alloc(h, &P.W, wsz * h->grid))) return rc;
    if ((rc = dev_alloc(h, &P.Hcol, (size_t)h->grid * P.m * P.wh))) return rc;
    if ((rc = dev_alloc(h, &P.Hval, (size_t)h->grid * P.m * P.wh))) return rc;
    if ((rc = dev_alloc(h, &P.F, (size_t)h->grid * P.m * P.d))) return rc;
    if ((rc = dev_alloc(h, &P.S, (size_t)h->grid * P.m * P.m))) return rc;
    CU(cudaMemset(P.W, 0, wsz * h->grid * sizeof(double)));
    h->have_op = true;
    return 0;
}

int pnmol_b200_set_prior(pnmol_b200_handle* h, const double* A1d, const double* LQ1d, const double* Lk) {
    if (!h || !A1d || !LQ1d || !Lk) return fail(-1, "null argument");
    if (h->have_prior) return fail(-1, "prior already set (create a new handle)");
    Problem& P = h->P;
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = dev_upload(h, &P.A1d, A1d, (size_t)P.n * P.n))) return rc;
    if ((rc = dev_upload(h, &P.LQ1d, LQ1d, (size_t)P.n * P.n))) return rc;
    if ((rc = dev_upload(h, &P.Lk, Lk, (size_t)P.d * P.d))) return rc;
    double* Kg = nullptr;
    if ((rc = dev_alloc(h, &Kg, (size_t)P.d * P.d))) return rc;
    k_gram<<<P.d, 128>>>(P.Lk, Kg, P.d);
    ++g_launches;
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    P.Kg = Kg;
    h->have_prior = true;
    return 0;
}

int pnmol_b200_set_members(pnmol_b200_handle* h, const double* diff_scale, const double* prior_scale,
                           const double* reaction_params, int nparams) {
    if (!h) return fail(-1, "null handle");
    Problem& P = h->P;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    int rc;
    if (diff_scale && (rc = dev_upload(h, &P.diffscale, diff_scale, (size_t)P.batch * P.ncomp))) return rc;
    if (prior_scale && (rc = dev_upload(h, &P.priorscale, prior_scale, (size_t)P.batch))) return rc;
    if (reaction_params) {
        const int need = P.reaction == 1 ? 1 : P.reaction == 2 ? 2 : P.reaction == 3 ? 4 : 0;
        if (nparams < need || nparams > PNMOL_B200_MAX_REACTION_PARAMS) return fail(-1, "wrong number of reaction parameters");
        if ((rc = dev_upload(h, &P.rparams, reaction_params, (size_t)P.batch * nparams))) return rc;
        P.nparams = nparams;
    }
    return 0;
}

int pnmol_b200_initialize(pnmol_b200_handle* h, const double* y0, double t0, double diffuse_prior_scale, double* mean_out,
                          double* chol_out, int32_t* status, void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (!y0 || !mean_out || !chol_out) return fail(-1, "null state pointer");
    InitArgs a;
    a.y0 = y0; a.t0 = t0; a.prior_scale0 = diffuse_prior_scale;
    a.nugget = h->P.latent ? 1e-6 : 1e-10;  // latent.py:71,98 / white.py:33,51
    a.mean_out = mean_out; a.chol_out = chol_out; a.status = status;
    if (h->small) {
        k_init_small<<<h->grid, 32 * h->sgeo.nwarps, h->smem_small, (cudaStream_t)stream>>>(h->P, a, h->sgeo);
    } else if (h->large) {
        int rc2 = launch_large(h, k_init_large, a, (cudaStream_t)stream);
        if (rc2) return rc2;
    } else {
        CU(cudaMemsetAsync(h->P.smslot, 0, sizeof(int) * h->num_sms, (cudaStream_t)stream));
        k_init<<<h->grid, kThreads, h->smem_bytes, (cudaStream_t)stream>>>(h->P, a);
    }
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_step(pnmol_b200_handle* h, double t_new, double dt, const double* precond, const double* precond_inv,
                    const double* mean_in, const double* chol_in, double* mean_out, double* chol_out, double* err_out,
                    double* ref_out, double* diff_out, int32_t* status, int flags, void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (!precond || !precond_inv || !mean_in || !chol_in || !mean_out || !chol_out) return fail(-1, "null argument");
    if (mean_in == mean_out || chol_in == chol_out) return fail(-1, "pnmol_b200_step is out of place");
    RunArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nsteps = 1; a.flags = flags; a.final_in_b = 1;
    for (int i = 0; i < h->P.n; ++i) { a.pv0[i] = precond[i]; a.pinv0[i] = precond_inv[i]; }
    a.dt0 = dt; a.tnew0 = t_new;
    a.mean_a = const_cast<double*>(mean_in); a.chol_a = const_cast<double*>(chol_in);
    a.mean_b = mean_out; a.chol_b = chol_out;
    a.err_out = err_out; a.ref_out = ref_out; a.diff_last = diff_out; a.status = status;
    if ((rc = launch_run(h, a, (cudaStream_t)stream))) return rc;
    CU(cudaGetLastError());
    return 0;
}

namespace {
int run_common(pnmol_b200_handle* h, double t0, const double* dts, const double* precond, const double* precond_inv,
               int nsteps, double* mean, double* chol, double* mean_tmp, double* chol_tmp, double* err_out, double* ref_out,
               double* diff_last, double* diff_sum, double* mean_traj, double* chol_traj, double* std_traj,
               int32_t* status, int flags, void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (nsteps <= 0) return fail(-1, "nsteps must be positive");
    if (!dts || !precond || !precond_inv || !mean || !chol || !mean_tmp || !chol_tmp) return fail(-1, "null argument");
    RunArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nsteps = nsteps; a.flags = flags; a.final_in_b = 0;
    if ((rc = upload_steps(h, nsteps, t0, dts, precond, precond_inv, (cudaStream_t)stream, &a))) return rc;
    a.mean_a = mean; a.chol_a = chol; a.mean_b = mean_tmp; a.chol_b = chol_tmp;
    a.err_out = err_out; a.ref_out = ref_out; a.diff_last = diff_last; a.diff_sum = diff_sum;
    a.mean_traj = mean_traj; a.chol_traj = chol_traj; a.std_traj = std_traj; a.status = status;
    if ((rc = launch_run(h, a, (cudaStream_t)stream))) return rc;
    CU(cudaGetLastError());
    return 0;
}
}  // namespace

int pnmol_b200_run(pnmol_b200_handle* h, double t0, const double* dts, const double* precond, const double* precond_inv,
                   int nsteps, double* mean, double* chol, double* mean_tmp, double* chol_tmp, double* err_out,
                   double* ref_out, double* diff_last, double* diff_sum, double* mean_traj, double* chol_traj,
                   int32_t* status, int flags, void* stream) {
    return run_common(h, t0, dts, precond, precond_inv, nsteps, mean, chol, mean_tmp, chol_tmp, err_out, ref_out, diff_last,
                      diff_sum, mean_traj, chol_traj, nullptr, status, flags, stream);
}

int pnmol_b200_run_marginals(pnmol_b200_handle* h, double t0, const double* dts, const double* precond,
                             const double* precond_inv, int nsteps, double* mean, double* chol, double* mean_tmp,
                             double* chol_tmp, double* diff_last, double* diff_sum, double* mean_traj, double* std_traj,
                             int32_t* status, int flags, void* stream) {
    if (!mean_traj || !std_traj) return fail(-1, "null trajectory buffer");
    return run_common(h, t0, dts, precond, precond_inv, nsteps, mean, chol, mean_tmp, chol_tmp, nullptr, nullptr, diff_last,
                      diff_sum, mean_traj, nullptr, std_traj, status, flags, stream);
}

int pnmol_b200_run_adaptive(pnmol_b200_handle* h, double t0, double tmax, const double* dt0, double abstol, double reltol,
                            double change_min, double change_max, double safety_scale, int max_attempts, double* mean,
                            double* chol, double* mean_tmp, double* chol_tmp, double* t_out, double* dt_out,
                            double* diff_sum, double* diff_last, int32_t* num_steps, int32_t* num_attempts,
                            int32_t* status, int flags, void* stream) {
    return pnmol_b200_run_adaptive_trajectory(h, t0, tmax, dt0, abstol, reltol, change_min, change_max, safety_scale,
                                              max_attempts, mean, chol, mean_tmp, chol_tmp, t_out, dt_out, diff_sum, diff_last,
                                              num_steps, num_attempts, status, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                              flags, stream);
}

int pnmol_b200_run_adaptive_trajectory(pnmol_b200_handle* h, double t0, double tmax, const double* dt0, double abstol,
                                       double reltol, double change_min, double change_max, double safety_scale,
                                       int max_attempts, double* mean, double* chol, double* mean_tmp, double* chol_tmp,
                                       double* t_out, double* dt_out, double* diff_sum, double* diff_last, int32_t* num_steps,
                                       int32_t* num_attempts, int32_t* status, double* err_last, double* ref_last,
                                       double* t_traj, double* mean_traj, double* chol_traj, int max_traj, int flags,
                                       void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (h->P.latent) return fail(-1, "adaptive steps need an error estimate: white-noise solvers only (src/pnmol/latent.py:217-223)");
    if (!dt0 || !mean || !chol || !mean_tmp || !chol_tmp || !t_out || !dt_out || !diff_sum || !diff_last || !num_steps ||
        !num_attempts || !status)
        return fail(-1, "null argument");
    if (!(tmax > t0) || max_attempts <= 0) return fail(-1, "invalid time span or attempt limit");
    if ((mean_traj || chol_traj || t_traj) && (!mean_traj || !chol_traj || !t_traj || max_traj <= 0))
        return fail(-1, "the trajectory needs t_traj, mean_traj, chol_traj and max_traj > 0");
    if ((err_last == nullptr) != (ref_last == nullptr)) return fail(-1, "err_last and ref_last go together");
    const Problem& P = h->P;
    if (!h->hs_err) {
        CU(cudaMalloc((void**)&h->hs_err, sizeof(double) * P.batch * P.d));
        CU(cudaMalloc((void**)&h->hs_ref, sizeof(double) * P.batch * P.d));
    }
    AdaptiveArgs a;
    std::memset(&a, 0, sizeof(a));
    a.t0 = t0; a.tmax = tmax; a.abstol = abstol; a.reltol = reltol; a.change_min = change_min; a.change_max = change_max;
    a.safety = safety_scale; a.inv_rate = 1.0 / (double)P.n;  // local convergence rate = num_derivatives + 1 (pdefilter.py:215)
    a.dt0 = dt0; a.mean_a = mean; a.chol_a = chol; a.mean_b = mean_tmp; a.chol_b = chol_tmp;
    a.err = err_last ? err_last : h->hs_err; a.ref = ref_last ? ref_last : h->hs_ref;
    a.t_traj = t_traj; a.mean_traj = mean_traj; a.chol_traj = chol_traj; a.max_traj = max_traj;
    a.t_out = t_out; a.dt_out = dt_out; a.diff_sum = diff_sum; a.diff_last = diff_last;
    a.nsteps = num_steps; a.nattempts = num_attempts; a.status = status; a.max_attempts = max_attempts; a.flags = flags;
    if (h->large) {
        CU(cudaFuncSetAttribute(k_run_adaptive_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_large));
        if (h->cluster > 1) CU(cudaFuncSetAttribute(k_run_adaptive_large, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int rc2 = launch_large(h, k_run_adaptive_large, a, (cudaStream_t)stream);
        if (rc2) return rc2;
    } else if (h->small) {
        CU(cudaFuncSetAttribute(k_run_adaptive_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_small));
        k_run_adaptive_small<<<h->grid, 32 * h->sgeo.nwarps, h->smem_small, (cudaStream_t)stream>>>(h->P, a, h->sgeo);
    } else {
        CU(cudaFuncSetAttribute(k_run_adaptive, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        CU(cudaMemsetAsync(h->P.smslot, 0, sizeof(int) * h->num_sms, (cudaStream_t)stream));
        k_run_adaptive<<<h->grid, kThreads, h->smem_bytes, (cudaStream_t)stream>>>(h->P, a);
    }
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_marginal_std(const double* chol, double* std_out, int D, int num_derivatives, int count, int device, void* stream) {
    const int n = num_derivatives + 1;
    if (!chol || !std_out || D <= 0 || n < 1 || D % n || count <= 0) return fail(-1, "invalid argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    k_marginal_std<<<std::min(count, 4 * 148), 256, 0, (cudaStream_t)stream>>>(chol, std_out, D, n, D / n, count);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_profile(pnmol_b200_handle* h, int enable, uint64_t* cycles_out) {
    if (!h) return fail(-1, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    if (cycles_out && h->P.prof) CU(cudaMemcpy(cycles_out, h->P.prof, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (enable && !h->P.prof) {
        int rc = dev_alloc(h, &h->P.prof, 24);
        if (rc) return rc;
    }
    if (h->P.prof) CU(cudaMemset(h->P.prof, 0, 24 * sizeof(uint64_t)));
    if (!enable) h->P.prof = nullptr;
    return 0;
}

int pnmol_b200_cluster_size(pnmol_b200_handle* h, int* requested, int* observed) {
    if (!h || !requested || !observed) return fail(-1, "null argument");
    *requested = h->large ? h->cluster : 0;
    *observed = 0;
    if (h->large && h->q.nf) {
        int32_t v = 0;
        CU(cudaSetDevice(h->device));
        CU(cudaMemcpy(&v, h->q.nf + 2, sizeof(int32_t), cudaMemcpyDeviceToHost));
        *observed = v;
    }
    return 0;
}

int pnmol_b200_path(pnmol_b200_handle* h) {
    if (!h) return fail(-1, "null handle");
    if (!h->have_op) return fail(-1, "pnmol_b200_set_operator has not been called");
    return h->small ? 2 : (h->large ? 1 : 0);
}

int pnmol_b200_rescale(pnmol_b200_handle* h, double* chol, const double* diff_sum, int nsteps, double* diff_cal_out,
                       void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (!chol || !diff_sum || nsteps <= 0) return fail(-1, "invalid argument");
    const Problem& P = h->P;
    dim3 g(std::max(1, std::min(64, (int)((size_t)P.D * P.D / 1024))), std::min(P.batch, 32768));
    k_rescale<<<g, 256, 0, (cudaStream_t)stream>>>(chol, diff_sum, diff_cal_out, nsteps, (size_t)P.D * P.D, P.batch);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_simulate_final_state_host(pnmol_b200_handle* h, const double* y0_host, double t0, double diffuse_prior_scale,
                                         const double* dts, const double* precond, const double* precond_inv, int nsteps,
                                         double* mean_host, double* chol_host, double* diff_cal_host, int32_t* status_host,
                                         int flags, void* stream) {
    int rc = ensure_ready(h);
    if (rc) return rc;
    if (!y0_host || !mean_host || !chol_host) return fail(-1, "null host buffer");
    const Problem& P = h->P;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t msz = (size_t)P.batch * P.D, csz = (size_t)P.batch * P.D * P.D;
    if (!h->hs_y0) {
        CU(cudaMalloc((void**)&h->hs_y0, sizeof(double) * P.batch * P.d));
        CU(cudaMalloc((void**)&h->hs_mean_a, sizeof(double) * msz));
        CU(cudaMalloc((void**)&h->hs_mean_b, sizeof(double) * msz));
        CU(cudaMalloc((void**)&h->hs_chol_a, sizeof(double) * csz));
        CU(cudaMalloc((void**)&h->hs_chol_b, sizeof(double) * csz));
        CU(cudaMalloc((void**)&h->hs_diffsum, sizeof(double) * P.batch));
        CU(cudaMalloc((void**)&h->hs_diffcal, sizeof(double) * P.batch));
        CU(cudaMalloc((void**)&h->hs_status, sizeof(int32_t) * P.batch));
    }
    // Members are independent: the ensemble runs in chunks of whole waves of CTAs, and the device-to-host copy of a
    // finished chunk (the D x D factors dominate: 180 KB per member at D = 150) overlaps the next chunk's kernels on a
    // second stream.  One chunk = the whole ensemble when it is small or runs on the multi-CTA path.
    if (!h->copy_stream) {
        CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->chunk_ev, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->copy_ev, cudaEventDisableTiming));
    }
    int chunk = P.batch;
    if (!h->large && h->grid > 0 && P.batch >= 8 * h->grid) chunk = 4 * h->grid;
    if (const char* e = std::getenv("PNMOL_B200_HOST_CHUNK")) chunk = std::max(1, std::atoi(e));
    const Problem saved = h->P;
    const size_t D = P.D, DD = (size_t)P.D * P.D;
    for (int off = 0; off < saved.batch; off += chunk) {
        const int nb = std::min(chunk, saved.batch - off);
        h->P.batch = nb;
        if (saved.diffscale) h->P.diffscale = saved.diffscale + (size_t)off * saved.ncomp;
        if (saved.priorscale) h->P.priorscale = saved.priorscale + off;
        if (saved.rparams) h->P.rparams = saved.rparams + (size_t)off * saved.nparams;
        double* y0d = h->hs_y0 + (size_t)off * saved.d;
        double *ma = h->hs_mean_a + off * D, *mb = h->hs_mean_b + off * D, *ca = h->hs_chol_a + off * DD, *cb = h->hs_chol_b + off * DD;
        rc = 0;
        cudaError_t ce = cudaMemcpyAsync(y0d, y0_host + (size_t)off * saved.d, sizeof(double) * nb * saved.d, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess) rc = pnmol_b200_initialize(h, y0d, t0, diffuse_prior_scale, ma, ca, h->hs_status + off, stream);
        if (ce == cudaSuccess && !rc)
            rc = pnmol_b200_run(h, t0, dts, precond, precond_inv, nsteps, ma, ca, mb, cb, nullptr, nullptr, nullptr,
                                h->hs_diffsum + off, nullptr, nullptr, h->hs_status + off, flags, stream);
        if (ce == cudaSuccess && !rc) rc = pnmol_b200_rescale(h, ca, h->hs_diffsum + off, nsteps, h->hs_diffcal + off, stream);
        h->P = saved;
        if (ce != cudaSuccess) return fail(-2, std::string("cudaMemcpyAsync (y0): ") + cudaGetErrorString(ce));
        if (rc) return rc;
        CU(cudaEventRecord(h->chunk_ev, st));
        CU(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev, 0));
        CU(cudaMemcpyAsync(mean_host + off * D, ma, sizeof(double) * nb * D, cudaMemcpyDeviceToHost, h->copy_stream));
        CU(cudaMemcpyAsync(chol_host + off * DD, ca, sizeof(double) * nb * DD, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    if (diff_cal_host) CU(cudaMemcpyAsync(diff_cal_host, h->hs_diffcal, sizeof(double) * P.batch, cudaMemcpyDeviceToHost, st));
    if (status_host) CU(cudaMemcpyAsync(status_host, h->hs_status, sizeof(int32_t) * P.batch, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(h->copy_ev, h->copy_stream));
    CU(cudaStreamWaitEvent(st, h->copy_ev, 0));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------ dense sqrt functions
namespace {
// Workspace of the handle-less dense entry points: allocated and freed in stream order on the CALLER's stream
// (cudaMallocAsync / cudaFreeAsync), so concurrent calls on different streams or from different threads never share it.
struct StreamScratch {
    double* p = nullptr;
    cudaStream_t st = nullptr;
    int acquire(size_t count, cudaStream_t stream) {
        st = stream;
        cudaError_t e = cudaMallocAsync((void**)&p, std::max<size_t>(count, 1) * sizeof(double), st);
        if (e != cudaSuccess) return fail(-2, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        return 0;
    }
    ~StreamScratch() { if (p) cudaFreeAsync(p, st); }
};
}  // namespace

int pnmol_b200_sqrt_propagate(const double* S1, const double* S2, double* out, int r, int c1, int c2, int batch, int device,
                              void* stream) {
    if (!S1 || !out || r <= 0 || c1 <= 0 || c2 < 0 || batch <= 0 || (c2 > 0 && !S2)) return fail(-1, "invalid argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const int rows = c1 + c2;
    const int grid = std::min(batch, 296);
    StreamScratch ws;
    int rc = ws.acquire((size_t)grid * rows * r, (cudaStream_t)stream);
    if (rc) return rc;
    double* W = ws.p;
    const size_t smem = sizeof(double) * (rows + 4 + 2 * kWarps);
    CU(cudaFuncSetAttribute(k_sqrt_propagate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_sqrt_propagate<<<grid, kThreads, smem, (cudaStream_t)stream>>>(S1, S2, out, r, c1, c2, batch, W);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_sqrt_update(const double* H, const double* C, const double* meascov, double* C_out, double* K_out, double* S_out,
                           int m, int D, int batch, int device, void* stream) {
    if (!H || !C || !C_out || !K_out || !S_out || m <= 0 || D <= 0 || batch <= 0) return fail(-1, "invalid argument");
    if (meascov && D < m) return fail(-1, "update_sqrt needs D >= m (src/pnmol/base/sqrt.py:55-57)");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const int grid = std::min(batch, 296);
    StreamScratch ws;
    int rc = ws.acquire((size_t)grid * (D + m) * (m + D), (cudaStream_t)stream);
    if (rc) return rc;
    double* W = ws.p;
    const size_t smem = sizeof(double) * (D + m + 4 + 2 * kWarps);
    CU(cudaFuncSetAttribute(k_sqrt_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_sqrt_update<<<grid, kThreads, smem, (cudaStream_t)stream>>>(H, C, meascov, C_out, K_out, S_out, m, D, batch, W);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_smoother_step(const double* m, const double* sc, const double* m_fut, const double* sc_fut, const double* sgain,
                             const double* sq, const double* mp, const double* x, double* mean_out, double* chol_out, int d,
                             int batch, int device, void* stream) {
    if (!m || !sc || !m_fut || !sc_fut || !sgain || !sq || !mp || !x || !mean_out || !chol_out || d <= 0 || batch <= 0)
        return fail(-1, "invalid argument");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const int grid = std::min(batch, 296);
    StreamScratch ws;
    int rc = ws.acquire((size_t)grid * 3 * d * 2 * d, (cudaStream_t)stream);
    if (rc) return rc;
    double* W = ws.p;
    const size_t smem = sizeof(double) * (3 * d + 4 + 2 * kWarps);
    CU(cudaFuncSetAttribute(k_smoother_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_smoother_step<<<grid, kThreads, smem, (cudaStream_t)stream>>>(m, sc, m_fut, sc_fut, sgain, sq, mp, x, mean_out, chol_out, d,
                                                                    batch, W);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_fd_coefficients(int kernel_kind, const double* kernel_params, int nbatch, const double* x, const double* neighbors,
                               int npoints, int stencil, int diffop, double nugget_gram_matrix, double* weights,
                               double* uncertainties, int device, void* stream) {
    if (!kernel_params || !x || !neighbors || !weights || !uncertainties || nbatch <= 0 || npoints <= 0)
        return fail(-1, "invalid argument");
    if (kernel_kind < 0 || kernel_kind > 2) return fail(-1, "kernel_kind: 0 SquareExponential, 1 Matern52, 2 Polynomial");
    if (diffop < 0 || diffop > 1) return fail(-1, "diffop: 0 gradient, 1 laplace");
    if (stencil < 1 || stencil > kMaxStencil) return fail(-1, "stencil size must be 1..8");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t total = (size_t)nbatch * npoints;
    const unsigned grid = (unsigned)((total + 127) / 128);
    k_fd_coefficients<<<grid, 128, 0, (cudaStream_t)stream>>>(kernel_kind, kernel_params, nbatch, x, neighbors, npoints, stencil,
                                                              diffop, nugget_gram_matrix, weights, uncertainties);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

int pnmol_b200_gram_cholesky(int kernel_kind, const double* kernel_params, int nbatch, const double* points, int d,
                             double diagonal_add, const double* data, double* chol_out, double* loglik_out, int32_t* status_out,
                             int device, void* stream) {
    if (!kernel_params || !points || !chol_out || nbatch <= 0 || d <= 0) return fail(-1, "invalid argument");
    if (kernel_kind < 0 || kernel_kind > 2) return fail(-1, "kernel_kind: 0 SquareExponential, 1 Matern52, 2 Polynomial");
    if (loglik_out && !data) return fail(-1, "the log-likelihood needs data");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-3, "no such CUDA device (libpnmol_b200 has no CPU fallback)");
    CU(cudaSetDevice(device));
    const size_t smem = sizeof(double) * ((size_t)d * (d + 1) + d);
    int optin = 0;
    CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const bool in_smem = smem + 256 <= (size_t)optin;
    if (loglik_out && !in_smem) return fail(-4, "log-likelihood: d too large for the shared-memory factorisation");
    if (in_smem) CU(cudaFuncSetAttribute(k_gram_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gram_cholesky<<<std::min(nbatch, 8 * 148), 256, in_smem ? smem : 0, (cudaStream_t)stream>>>(
        kernel_kind, kernel_params, nbatch, points, d, diagonal_add, data, chol_out, loglik_out, status_out, in_smem ? 1 : 0);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"
