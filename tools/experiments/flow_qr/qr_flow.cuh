// Multi-CTA blocked Householder QR as a data-flow pipeline (north-star kernel 2, large state dimension), sm_100a.
//
// Same mathematics, LAPACK dlarfg / dlarft conventions, support envelopes and tile-aligned compact row lists as the other
// QR routines; what changes against the barrier-synchronised version (qr_large.cuh) is the schedule:
//
//   * The matrix (workspace W, column-major, L2 / HBM) is cut into column blocks of kNB = 16 columns.  Block b belongs
//     to CTA b mod G for good: only its owner ever reads or writes it, so successive updates of a block are ordered by
//     program order and no grid barrier is needed inside the factorisation.
//   * Block k is also panel k.  Its owner factors it as soon as it has applied panel k-1 to it, publishes the
//     reflectors V_k (global, compact row list, ring of `nslot` panels) and the compact-WY factor T_k, and raises the
//     `published` counter.  Every CTA applies V_k to its own blocks beyond k (the block that is the next panel first)
//     and acknowledges in fin[k]; a ring slot is rewritten only when all CTAs have acknowledged its previous panel.
//     The dependent per-column chain of panel k+1 thus overlaps the tensor-core updates of panels <= k on the other SMs.
//   * Panel factorisation on ONE CTA, four columns at a time held in REGISTERS across the whole CTA (thread t keeps
//     compact rows t, t + 256, ...): one block barrier per column (norm and the dot products with the later columns of
//     the sub-panel reduced together; all threads derive the dlarfg scalars redundantly), then the sub-panel's 4 x 4
//     block reflector is applied to the rest of the panel on the tensor pipe.
//   * Block-reflector application  C <- C - V T^T (V^T C)  with mma.sync.m8n8k4.f64: the V operand of a trailing update
//     is staged in shared memory by TMA (cp.async.bulk.tensor.2d + mbarrier, two stages of 240 rows x 16 reflectors)
//     and shared by the CTA's eight warps; the C tiles are private to a lane (no reuse), so they go global -> registers
//     as 16-byte loads.  Partial Y^T = C^T V of the warps are summed in shared memory in a fixed order (bitwise
//     reproducible, no floating-point atomics).
#pragma once
#include <cuda.h>
#include <cooperative_groups.h>

namespace pnmol {

namespace cg = cooperative_groups;

constexpr int kFlowChunkTiles = 30;                      // 8-row tiles per TMA stage
constexpr int kFlowChunkRows = 8 * kFlowChunkTiles;      // 240
constexpr int kFlowPitch = kFlowChunkRows + 2;           // 242 = 2 (mod 16): conflict-free second-product operand loads
constexpr int kFlowStage = kNB * kFlowPitch;             // doubles per stage (30976 bytes, a multiple of 128)
constexpr int kFlowExtra = 512;                          // red | bc | Tf (factor-side T)  in doubles
constexpr int kFlowMaxStages = 8;
constexpr int kFlowFin = 16;                             // flags[0] = published panels, flags[kFlowFin + k] = fin[k]

// ---------------------------------------------------------------- flags
__device__ __forceinline__ int flow_ld_acquire(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flow_st_release(int32_t* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// Whole CTA waits until *flag >= target (thread 0 polls).  Only the CTA on the critical path polls eagerly: 147 CTAs
// hammering one L2 line every few dozen nanoseconds slow every L2 access of the CTA that factors the panel down.
__device__ __forceinline__ void flow_wait_ge(const int32_t* flag, int target, bool eager) {
    if (threadIdx.x == 0) {
        const unsigned ns = eager ? 32u : 1024u;
        while (flow_ld_acquire(flag) < target) __nanosleep(ns);
        asm volatile("fence.proxy.async;\n" ::: "memory");  // the published reflectors are read through the async proxy (TMA)
    }
    __syncthreads();
}

// ---------------------------------------------------------------- TMA / mbarrier
__device__ __forceinline__ unsigned flow_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void flow_mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(flow_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void flow_tma_load(double* dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar, unsigned bytes) {
    const unsigned b = flow_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(flow_smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(b) : "memory");
}
__device__ __forceinline__ void flow_mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned b = flow_smem_u32(bar);
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

struct FlowSmem {
    double* A;        // region A: reflectors of the panel being factored ([16][vp] or [4][vp]) | two TMA stages
    double* red;      // [2][8][8] per-warp partial sums of the column reductions
    double* bc;       // [2][8]   diagonal-row entries of the sub-panel columns
    double* Tf;       // [16][kLdr] T factor under construction (T4 of the sub-panel in its top-left corner)
    double* ysh;      // [8][128] per-warp partial Y^T
    uint64_t* bars;   // one mbarrier per TMA stage
    int nstage;       // TMA stages that fit region A (2 .. kFlowMaxStages)
};

// ---------------------------------------------------------------- block-reflector application (one CTA)
// Columns [cbeg, cend) of the workspace (at most 16: one or two groups of 8), rows = the compact list `rm` of the panel
// the reflectors belong to.  SUB: 4 reflectors resident in shared memory (Vs = [4][vp], T4 in the corner of Ts);
// otherwise 16 reflectors streamed from the ring slot `vy` of the tensor map in chunks of 240 rows.
// VEC: 16-byte tile accesses (even segment offsets).  Requires rm.aligned.
template <bool SUB, bool VEC>
static __device__ __noinline__ void flow_apply(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap rm, const double* __restrict__ Vs,
                                  int vp, const double* __restrict__ Ts, const FlowSmem& fs, const CUtensorMap* tmap, int vy,
                                  unsigned& nload, PhaseClock& pc) {
    const int nsub = SUB ? vy : kNB;  // SUB: `vy` carries the number of resident reflectors (2 or 4)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ncols = cend - cbeg;
    const int ngroups = ncols > 8 ? 2 : 1;
    const int grp = ngroups == 2 ? (warp & 1) : 0, cls = ngroups == 2 ? (warp >> 1) : warp, ncls = kWarps / ngroups;
    const int ntile = (rm.len + 7) >> 3;
    const int nt1 = (rm.len1 + 7) >> 3;  // tiles [0, nt1) lie in the first segment
    const int off1 = rm.j0 + 2 * t, off2 = rm.a2 - rm.len1 + 2 * t;
    const int col = cbeg + 8 * grp + g;
    const bool have = col < cend;
    double* cp = W + (size_t)(have ? col : cbeg) * ld;
    const int nch = SUB ? (ntile + kFlowChunkTiles - 1) / kFlowChunkTiles : (ntile + kFlowChunkTiles - 1) / kFlowChunkTiles;
    const int NS = fs.nstage;
    const bool reload = !SUB && nch > NS;   // all chunks resident: the second product needs no further loads
    constexpr unsigned kBytes = (unsigned)(kFlowStage * sizeof(double));
    double* const stage0 = fs.A;
    uint64_t* const bars = fs.bars;
    double* const ysh = fs.ysh;
    const int total = reload ? 2 * nch : nch;   // TMA loads of this call
    if (!SUB) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // region A was last written through the generic proxy
            for (int j = 0; j < NS && j < total; ++j) {
                const unsigned sj = (nload + j) % NS;
                flow_tma_load(stage0 + (size_t)sj * kFlowStage, tmap, (j % nch) * kFlowChunkRows, vy, &bars[sj], kBytes);
            }
        }
    }
    constexpr int kTB = 8;  // tiles of a warp per chunk: at most ceil(30 / 4)
    double y[2][2][2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int n = 0; n < 2; ++n) { y[e][n][0] = 0.0; y[e][n][1] = 0.0; }
    double z[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    unsigned seq = nload;  // running index of the TMA loads of this CTA (stage = seq & 1, parity = (seq >> 1) & 1)
    // C tiles of the warp for chunk `ch`: a = first + i ncls.  The loads of the next (pass, chunk) are issued before the
    // tensor work of the current one (software pipeline over the flat sequence pass 0 chunks, then pass 1 chunks).
    auto first_tile = [&](int ch) { const int tb = ch * kFlowChunkTiles; return tb + ((cls - tb) % ncls + ncls) % ncls; };
    auto chunk_end = [&](int ch) { const int te = (ch + 1) * kFlowChunkTiles; return te < ntile ? te : ntile; };
    double xn[kTB][2];
    {
        const int f0 = first_tile(0), e0 = chunk_end(0);
#pragma unroll
        for (int i = 0; i < kTB; ++i) {
            const int a = f0 + i * ncls;
            xn[i][0] = 0.0; xn[i][1] = 0.0;
            if (a < e0) {
                const double* p = cp + (a < nt1 ? off1 : off2) + 8 * a;
                if (VEC) { const double2 v = *reinterpret_cast<const double2*>(p); xn[i][0] = v.x; xn[i][1] = v.y; }
                else { xn[i][0] = p[0]; xn[i][1] = p[1]; }
            }
        }
    }
    pc.mark(SUB ? 17 : 20);
#pragma unroll 1
    for (int it = 0; it < 2 * nch; ++it) {
        const int pass = it >= nch ? 1 : 0, ch = it - pass * nch;
        const double* vb;
        int vpitch, row0;
        if (SUB) {
            vb = Vs; vpitch = vp; row0 = 0;
        } else {
            const unsigned sq = (pass == 1 && !reload) ? nload + ch : seq;
            vb = stage0 + (size_t)(sq % NS) * kFlowStage; vpitch = kFlowPitch; row0 = ch * kFlowChunkRows;
            if (pass == 0 || reload) flow_mbar_wait(&bars[sq % NS], (sq / NS) & 1);
        }
        const int first = first_tile(ch), tend = chunk_end(ch);
        double xa[kTB][2];
#pragma unroll
        for (int i = 0; i < kTB; ++i) { xa[i][0] = xn[i][0]; xa[i][1] = xn[i][1]; }
        // prefetch the tiles of the next step -- except across the pass boundary when a chunk is revisited at once
        // (one chunk: its second-pass tiles are the first-pass tiles still in xa) 
        const bool pre = it + 1 < 2 * nch && !(nch == 1);
        if (pre) {
            const int chn = (it + 1) >= nch ? it + 1 - nch : it + 1;
            const int fn = first_tile(chn), en = chunk_end(chn);
#pragma unroll
            for (int i = 0; i < kTB; ++i) {
                const int a = fn + i * ncls;
                xn[i][0] = 0.0; xn[i][1] = 0.0;
                if (a < en) {
                    const double* p = cp + (a < nt1 ? off1 : off2) + 8 * a;
                    if (VEC) { const double2 v = *reinterpret_cast<const double2*>(p); xn[i][0] = v.x; xn[i][1] = v.y; }
                    else { xn[i][0] = p[0]; xn[i][1] = p[1]; }
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int i = 0; i < kTB; ++i) {
                const int a = first + i * ncls;
                if (a < tend) {
                    const double* v1 = vb + (size_t)g * vpitch + (8 * a - row0) + 2 * t;
                    if (SUB) {
                        double2 lo = make_double2(0.0, 0.0);
                        if (g < nsub) lo = *reinterpret_cast<const double2*>(v1);
                        dmma884(y[0][0][0], y[0][0][1], xa[i][0], lo.x);
                        dmma884(y[1][0][0], y[1][0][1], xa[i][1], lo.y);
                    } else {
                        const double2 lo = *reinterpret_cast<const double2*>(v1);
                        const double2 hi = *reinterpret_cast<const double2*>(v1 + (size_t)8 * vpitch);
                        dmma884(y[0][0][0], y[0][0][1], xa[i][0], lo.x);
                        dmma884(y[0][1][0], y[0][1][1], xa[i][0], hi.x);
                        dmma884(y[1][0][0], y[1][0][1], xa[i][1], lo.y);
                        dmma884(y[1][1][0], y[1][1][1], xa[i][1], hi.y);
                    }
                }
            }
            if (nch == 1) {  // the only chunk: keep its tiles for the second product
#pragma unroll
                for (int i = 0; i < kTB; ++i) { xn[i][0] = xa[i][0]; xn[i][1] = xa[i][1]; }
            }
        } else {
#pragma unroll
            for (int i = 0; i < kTB; ++i) {
                const int a = first + i * ncls;
                if (a < tend) {
                    const double* v2 = vb + (size_t)(2 * t) * vpitch + (8 * a - row0) + g;
                    if (SUB) {
                        double a0 = 0.0, b0 = 0.0;
                        if (2 * t < nsub) { a0 = v2[0]; b0 = v2[vpitch]; }
                        dmma884(xa[i][0], xa[i][1], z[0][0], a0);
                        dmma884(xa[i][0], xa[i][1], z[0][1], b0);
                    } else {
                        const double a0 = v2[0], b0 = v2[vpitch], a1 = v2[(size_t)8 * vpitch], b1 = v2[(size_t)9 * vpitch];
                        dmma884(xa[i][0], xa[i][1], z[0][0], a0);
                        dmma884(xa[i][0], xa[i][1], z[0][1], b0);
                        dmma884(xa[i][0], xa[i][1], z[1][0], a1);
                        dmma884(xa[i][0], xa[i][1], z[1][1], b1);
                    }
                }
            }
            if (have) {
#pragma unroll
                for (int i = 0; i < kTB; ++i) {
                    const int a = first + i * ncls;
                    if (a < tend) {
                        double* p = cp + (a < nt1 ? off1 : off2) + 8 * a;
                        if (VEC) *reinterpret_cast<double2*>(p) = make_double2(xa[i][0], xa[i][1]);
                        else { p[0] = xa[i][0]; p[1] = xa[i][1]; }
                    }
                }
            }
        }
        if (!SUB && (pass == 0 || reload)) {
            if (reload) __syncthreads();  // every warp is done with this stage
            // next load into the stage just released: sequence position + NS
            const int pos = (int)(seq - nload) + NS;
            if (tid == 0 && pos < total) flow_tma_load(stage0 + (size_t)(seq % NS) * kFlowStage, tmap, (pos % nch) * kFlowChunkRows, vy, &bars[seq % NS], kBytes);
            ++seq;
        }
        if (it == nch - 1) {
            pc.mark(SUB ? 18 : 21);
            // partial Y^T of this warp -> shared memory; totals over the warps of the column group in a fixed order
            double* mine = ysh + warp * 128 + lane;
            mine[0] = have ? y[0][0][0] + y[1][0][0] : 0.0;
            mine[32] = have ? y[0][0][1] + y[1][0][1] : 0.0;
            mine[64] = have ? y[0][1][0] + y[1][1][0] : 0.0;
            mine[96] = have ? y[0][1][1] + y[1][1][1] : 0.0;
            __syncthreads();
            double yt[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int c = 0; c < ncls; ++c) {
                const double* p = ysh + (ngroups == 2 ? 2 * c + grp : c) * 128 + lane;
                yt[0][0] += p[0]; yt[0][1] += p[32]; yt[1][0] += p[64]; yt[1][1] += p[96];
            }
            // Y'^T = -(Y^T T): z[n][q] = Y'^T[col g][reflector 8 n + 2 t + q]
#pragma unroll
            for (int h = 0; h < (SUB ? 1 : 2); ++h)
#pragma unroll
                for (int sx = 0; sx < 2; ++sx) {
                    const double* tp = Ts + (8 * h + 2 * t + sx) * kLdr + g;
                    dmma884(z[0][0], z[0][1], yt[h][sx], tp[0]);
                    if (!SUB) dmma884(z[1][0], z[1][1], yt[h][sx], tp[8]);
                }
#pragma unroll
            for (int n = 0; n < 2; ++n) { z[n][0] = -z[n][0]; z[n][1] = -z[n][1]; }
            pc.mark(SUB ? 19 : 22);
        }
    }
    nload = seq;
    __syncthreads();  // ysh and the stages are reused by the next call
}

template <bool SUB>
__device__ __forceinline__ void flow_apply_dispatch(double* __restrict__ W, int ld, int cbeg, int cend, const RowMap& rm,
                                                    const double* Vs, int vp, const double* Ts, const FlowSmem& fs,
                                                    const CUtensorMap* tmap, int vy, unsigned& nload, PhaseClock& pc) {
    if (cbeg >= cend) return;
    const bool vec = (((rm.j0 | (rm.a2 - rm.len1) | ld) & 1) == 0) && ((reinterpret_cast<size_t>(W) & 15) == 0);
    if (vec) flow_apply<SUB, true>(W, ld, cbeg, cend, rm, Vs, vp, Ts, fs, tmap, vy, nload, pc);
    else flow_apply<SUB, false>(W, ld, cbeg, cend, rm, Vs, vp, Ts, fs, tmap, vy, nload, pc);
}

// ---------------------------------------------------------------- sub-panel factorisation (whole CTA, registers)
// Panel columns c0 .. c0 + nc - 1 (nc <= 4) of the panel that starts at workspace column j0; the diagonal of panel
// column p is compact row p (thread p, slot 0).  Writes R (and zeros below the diagonal inside the envelope) to the
// workspace, the reflectors (unit diagonal, zeros above) to Vsub[q][vp] (shared) and Vg[(c0 + q) lv + c] (global ring
// slot), tau to sc[3 (c0 + q)], the sub-panel's 4 x 4 T factor to the top-left corner of Tf.
template <int RPT, int SC, bool SM>
static __device__ __noinline__ void flow_subpanel(double* __restrict__ W, int ld, const Shape& s, int j0, int c0, int nc, const RowMap rm,
                                     double* __restrict__ Vsub, int vp, double* __restrict__ Vg, int lv, double* __restrict__ sc,
                                     const FlowSmem& fs, PhaseClock& pc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = rm.len, lpad = (L + 7) & ~7, nt = s.nt;
    double* const red_base = fs.red;
    double* const bc_base = fs.bc;
    double* const Tf = fs.Tf;
    double x[SC][RPT];
    int et[SC], eb[SC];
#pragma unroll
    for (int q = 0; q < SC; ++q) {
        const bool valid = q < nc;
        et[q] = valid ? env_top(s, j0 + c0 + q) : -1;
        eb[q] = valid ? env_bot(s, j0 + c0 + q) : -1;
        const double* col = W + (size_t)(j0 + c0 + (valid ? q : 0)) * ld;
        const double* scol = Vsub + (size_t)q * vp;   // SM: the panel sits in shared memory (zero outside the envelopes)
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int c = tid + kThreads * r;
            if (SM) {
                x[q][r] = (valid && c < lpad) ? scol[c] : 0.0;
            } else {
                const int row = rm.row(c);
                const bool ok = valid && c < L && (row < nt ? row <= et[q] : row <= eb[q]);
                x[q][r] = ok ? col[row] : 0.0;
            }
        }
    }
    double tauv[SC];
    pc.mark(23);
#pragma unroll
    for (int i = 0; i < SC; ++i) {
        const int p = c0 + i;
        double* wcol = W + (size_t)(j0 + p) * ld;
        double* bc = bc_base + (i & 1) * 8;
        double* red = red_base + (i & 1) * 64;
        // rows above the diagonal hold finished R entries; the diagonal entry and the same row of the later columns are
        // published; both leave the register copy of column i, which then is the sub-diagonal part only
        if (i < nc && tid < p) wcol[rm.row(tid)] = x[i][0];
        if (tid == p) {
#pragma unroll
            for (int k = 0; k < SC; ++k)
                if (k >= i) bc[k] = x[k][0];
        }
        if (tid <= p) x[i][0] = 0.0;
        double part[SC];
#pragma unroll
        for (int k = 0; k < SC; ++k) {
            if (k >= i) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; r += 2) {
                    a0 = fma(x[i][r], x[k][r], a0);
                    if (r + 1 < RPT) a1 = fma(x[i][r + 1], x[k][r + 1], a1);
                }
                part[k] = a0 + a1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double tmp[SC];
#pragma unroll
            for (int k = 0; k < SC; ++k)
                if (k >= i) tmp[k] = __shfl_xor_sync(0xffffffffu, part[k], o);
#pragma unroll
            for (int k = 0; k < SC; ++k)
                if (k >= i) part[k] += tmp[k];
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < SC; ++k)
                if (k >= i) red[warp * 8 + k] = part[k];
        }
        __syncthreads();
        // totals over the 8 warps: lane l fetches the partial of warp (l & 7) for column (l >> 3), three butterfly stages
        // (a fixed summation tree), then the totals are broadcast from lanes 0, 8, 16, 24
        double tot[SC], e[SC];
        {
            const int kk = lane >> 3;
            double v = (kk < SC && kk >= i) ? red[(lane & 7) * 8 + kk] : 0.0;
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
#pragma unroll
            for (int k = 0; k < SC; ++k) {
                if (k >= i) {
                    tot[k] = __shfl_sync(0xffffffffu, v, 8 * k);
                    e[k] = bc[k];
                }
            }
        }
        const double ss = tot[i], al = e[i];
        double tau = 0.0, beta = al, scale = 0.0;
        if (i < nc && ss != 0.0) {  // dlarfg: xnorm == 0 -> H = I
            const double s2 = fma(al, al, ss);
            const double rn = rsqrt(s2);
            const double nrm = s2 * rn;
            beta = -copysign(nrm, al);  // Fortran SIGN semantics of dlarfg
            tau = (beta - al) * -copysign(rn, al);
            scale = __drcp_rn(al - beta);
        }
        tauv[i] = tau;
#pragma unroll
        for (int k = 0; k < SC; ++k) {
            if (k > i) {
                const double f = -tau * fma(scale, tot[k], e[k]);
                const double gk = f * scale;
#pragma unroll
                for (int r = 0; r < RPT; ++r) x[k][r] = fma(gk, x[i][r], x[k][r]);
                if (tid == p) x[k][0] += f;
            }
        }
        // column i becomes the reflector: scale * x below the diagonal, one on it (zero when H = I), zeros above
#pragma unroll
        for (int r = 0; r < RPT; ++r) x[i][r] *= scale;
        if (tid == p) {
            x[i][0] = tau != 0.0 ? 1.0 : 0.0;
            if (i < nc) wcol[rm.row(p)] = beta;
        }
    }
    pc.mark(12);
    // zeros below the diagonal inside the column's own envelope; reflectors to shared memory and to the ring slot
#pragma unroll
    for (int q = 0; q < SC; ++q) {
        const int p = c0 + q;
        double* wcol = W + (size_t)(j0 + p) * ld;
        double* vs = Vsub + (size_t)q * vp;
        double* vg = Vg + (size_t)p * lv;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int c = tid + kThreads * r;
            if (c < lpad) {
                if (q < nc && c > p && c < L) {
                    const int row = rm.row(c);
                    if (row < nt ? row <= et[q] : row <= eb[q]) wcol[row] = 0.0;
                }
                vs[c] = x[q][r];
                vg[c] = x[q][r];
            }
        }
    }
    // T factor of the sub-panel (dlarft, forward / columnwise) from the Gram entries v_i . v_k
    constexpr int NG = SC * (SC - 1) / 2;
    double gq[NG];
    {
        int idx = 0;
#pragma unroll
        for (int i = 0; i < SC - 1; ++i)
#pragma unroll
            for (int k = i + 1; k < SC; ++k) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; r += 2) {
                    a0 = fma(x[i][r], x[k][r], a0);
                    if (r + 1 < RPT) a1 = fma(x[i][r + 1], x[k][r + 1], a1);
                }
                gq[idx++] = a0 + a1;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double tmp[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) tmp[k] = __shfl_xor_sync(0xffffffffu, gq[k], o);
#pragma unroll
        for (int k = 0; k < NG; ++k) gq[k] += tmp[k];
    }
    double* red = red_base;  // (buffer 0 was last read before the barrier of the last column)
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NG; ++k) red[warp * 8 + k] = gq[k];
    }
    __syncthreads();
    if (tid == 0) {
        double gs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // v_i . v_k for (i, k) = 01 02 03 12 13 23  (SC = 2: only 01)
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) sum += red[w * 8 + k];
            gs[k] = sum;
        }
        const double t00 = tauv[0], t11 = tauv[1], t22 = SC > 2 ? tauv[SC > 2 ? 2 : 0] : 0.0, t33 = SC > 2 ? tauv[SC > 2 ? 3 : 0] : 0.0;
        const double t01 = -t11 * (t00 * gs[0]);
        const double t02 = -t22 * fma(t01, gs[3], t00 * gs[1]);
        const double t12 = -t22 * (t11 * gs[3]);
        const double t03 = -t33 * fma(t02, gs[5], fma(t01, gs[4], t00 * gs[2]));
        const double t13 = -t33 * fma(t12, gs[5], t11 * gs[4]);
        const double t23 = -t33 * (t22 * gs[5]);
        double* T = Tf;
        T[0] = t00; T[1] = t01; T[2] = t02; T[3] = t03;
        T[kLdr] = 0.0; T[kLdr + 1] = t11; T[kLdr + 2] = t12; T[kLdr + 3] = t13;
        T[2 * kLdr] = 0.0; T[2 * kLdr + 1] = 0.0; T[2 * kLdr + 2] = t22; T[2 * kLdr + 3] = t23;
        T[3 * kLdr] = 0.0; T[3 * kLdr + 1] = 0.0; T[3 * kLdr + 2] = 0.0; T[3 * kLdr + 3] = t33;
#pragma unroll
        for (int i = 0; i < SC; ++i) sc[3 * (c0 + i)] = tauv[i];
    }
    __syncthreads();
}

// Gram matrix of the panel's reflectors on the tensor pipe, operands prefetched four tiles ahead (the reflectors sit in
// shared memory when the whole panel fits, else in the ring slot in global memory / L2).
static __device__ __noinline__ void flow_gram(const double* __restrict__ V, int ldv, int len, double* __restrict__ Gs, double* __restrict__ scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntile = (len + 7) >> 3;
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    double d00[2] = {0.0, 0.0}, d01[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
    const double* v0 = V + (size_t)g * ldv + 2 * t;
    const double* v1 = v0 + (size_t)8 * ldv;
    for (int i0 = warp; i0 < ntile; i0 += 4 * kWarps) {
        double2 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kWarps;
            a[u] = make_double2(0.0, 0.0); b[u] = make_double2(0.0, 0.0);
            if (i < ntile) {
                a[u] = *reinterpret_cast<const double2*>(v0 + 8 * i);
                b[u] = *reinterpret_cast<const double2*>(v1 + 8 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dmma884(c00[0], c00[1], a[u].x, a[u].x);
            dmma884(c01[0], c01[1], a[u].x, b[u].x);
            dmma884(c11[0], c11[1], b[u].x, b[u].x);
            dmma884(d00[0], d00[1], a[u].y, a[u].y);
            dmma884(d01[0], d01[1], a[u].y, b[u].y);
            dmma884(d11[0], d11[1], b[u].y, b[u].y);
        }
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

// Factor panel k (columns j0 .. j0 + nbk - 1) on this CTA and write V_k / T_k into ring slot `slot`.
template <int RPT, int SC>
static __device__ __noinline__ void flow_panel_factor(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk, const RowMap rm,
                                         const LargeQR& q, const LargeSmem& ls, const FlowSmem& fs, int slot, unsigned& nload,
                                         PhaseClock& pc) {
    const int tid = threadIdx.x;
    const int L = rm.len, lpad = (L + 7) & ~7;
    const bool v16 = q.v16 != 0;
    double* Vg = q.Vg + (size_t)slot * kNB * q.lv;
    for (int idx = tid; idx < kNB * kLdr; idx += kThreads) fs.Tf[idx] = 0.0;
    for (int idx = tid; idx < 3 * kNB; idx += kThreads) ls.sc[idx] = 0.0;
    // reflector slots beyond the panel are zero
    for (int idx = tid; idx < (kNB - nbk) * lpad; idx += kThreads) {
        const int r = nbk + idx / lpad, c = idx % lpad;
        Vg[(size_t)r * q.lv + c] = 0.0;
        if (v16) fs.A[(size_t)r * q.vp + c] = 0.0;
    }
    __syncthreads();
    if (v16) {
        // the whole panel in shared memory (entries outside a column's own envelope are zero); finished columns become
        // the reflectors in place, the in-panel block-reflector applications never leave the SM
        __shared__ int penv[2 * kNB];
        if (tid < nbk) { penv[tid] = env_top(s, j0 + tid); penv[kNB + tid] = env_bot(s, j0 + tid); }
        __syncthreads();
        const int total = nbk * lpad;
        for (int e0 = tid; e0 < total; e0 += 8 * kThreads) {  // eight loads in flight per thread
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * kThreads;
                v[u] = 0.0;
                if (e < total) {
                    const int qc = e / lpad, c = e - qc * lpad;
                    const int row = rm.row(c);
                    if (c < L && (row < s.nt ? row <= penv[qc] : row <= penv[kNB + qc])) v[u] = W[(size_t)(j0 + qc) * ld + row];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * kThreads;
                if (e < total) {
                    const int qc = e / lpad, c = e - qc * lpad;
                    fs.A[(size_t)qc * q.vp + c] = v[u];
                }
            }
        }
        __syncthreads();
        pc.mark(13);
        RowMap rs;  // compact index space of the shared-memory panel: one contiguous segment starting at 0
        rs.j0 = 0; rs.len1 = lpad; rs.a2 = lpad; rs.len = lpad; rs.aligned = true;
        for (int c0 = 0; c0 < nbk; c0 += SC) {
            const int nc = nbk - c0 < SC ? nbk - c0 : SC;
            double* Vsub = fs.A + (size_t)c0 * q.vp;
            flow_subpanel<RPT, SC, true>(W, ld, s, j0, c0, nc, rm, Vsub, q.vp, Vg, q.lv, ls.sc, fs, pc);
            pc.mark(14);
            if (c0 + SC < nbk) flow_apply_dispatch<true>(fs.A, q.vp, c0 + SC, nbk, rs, Vsub, q.vp, fs.Tf, fs, nullptr, SC, nload, pc);
            pc.mark(15);
        }
    } else {
        for (int c0 = 0; c0 < nbk; c0 += SC) {
            const int nc = nbk - c0 < SC ? nbk - c0 : SC;
            flow_subpanel<RPT, SC, false>(W, ld, s, j0, c0, nc, rm, fs.A, q.vp, Vg, q.lv, ls.sc, fs, pc);
            pc.mark(14);
            if (c0 + SC < nbk) flow_apply_dispatch<true>(W, ld, j0 + c0 + SC, j0 + nbk, rm, fs.A, q.vp, fs.Tf, fs, nullptr, SC, nload, pc);
            pc.mark(15);
        }
    }
    // T factor of the whole panel
    flow_gram(v16 ? fs.A : Vg, v16 ? q.vp : q.lv, L, ls.Gs, ls.scratch);
    if (tid < 32) panel_t_factor(ls.Gs, ls.sc, nbk, fs.Tf);
    __syncthreads();
    double* Tg = q.Tg + (size_t)slot * kNB * kLdr;
    for (int idx = tid; idx < kNB * kLdr; idx += kThreads) Tg[idx] = fs.Tf[idx];
    pc.mark(16);
}

__device__ __forceinline__ void flow_panel_factor_dispatch(double* __restrict__ W, int ld, const Shape& s, int j0, int nbk,
                                                           const RowMap& rm, const LargeQR& q, const LargeSmem& ls, const FlowSmem& fs,
                                                           int slot, unsigned& nload, PhaseClock& pc) {
    const int lpad = (rm.len + 7) & ~7;
    if (lpad <= 2 * kThreads) flow_panel_factor<2, 4>(W, ld, s, j0, nbk, rm, q, ls, fs, slot, nload, pc);
    else if (lpad <= 6 * kThreads) flow_panel_factor<6, 4>(W, ld, s, j0, nbk, rm, q, ls, fs, slot, nload, pc);
    else flow_panel_factor<17, 2>(W, ld, s, j0, nbk, rm, q, ls, fs, slot, nload, pc);  // (two columns at a time: 34 values per thread)
}

// ---------------------------------------------------------------- driver
// Grid-wide blocked QR.  On return (after a grid barrier) the upper triangle holds R.
static __device__ void householder_qr_large(cg::grid_group& grid, double* __restrict__ W, int ld, const Shape s, const LargeQR& q,
                                            const LargeSmem& ls, PhaseClock& pc) {
    const int tid = threadIdx.x;
    const int nrows = s.nt + s.nbot;
    const int nref = nrows < s.ncols ? nrows : s.ncols;
    const int np = (nref + kNB - 1) / kNB, nblk = (s.ncols + kNB - 1) / kNB;
    const int G = (int)gridDim.x, me = (int)blockIdx.x, S = q.nslot;
    int32_t* flags = q.flags;
    int32_t* fin = flags + kFlowFin;
    FlowSmem fs;
    {
        double* base = reinterpret_cast<double*>((reinterpret_cast<size_t>(ls.PB) + 127) & ~(size_t)127);
        fs.A = base;
        double* extra = base + q.regionA;
        fs.red = extra; fs.bc = extra + 128; fs.Tf = extra + 160;
        fs.ysh = ls.scratch;
    }
    fs.bars = reinterpret_cast<uint64_t*>(ls.bars);
    fs.nstage = q.regionA / kFlowStage < kFlowMaxStages ? q.regionA / kFlowStage : kFlowMaxStages;
    for (int k = me * kThreads + tid; k < np + kFlowFin; k += G * kThreads) flags[k] = 0;
    __threadfence();
    grid.sync();
    unsigned nload = *ls.nload;   // (uniform over the CTA; written back at the end)
    auto factor_and_publish = [&](int k) {
        const int j0 = k * kNB;
        const int nbk = nref - j0 < kNB ? nref - j0 : kNB;
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        if (k >= S) flow_wait_ge(&fin[k - S], G, true);  // nobody reads the previous panel of this ring slot any more
        pc.mark(12);
        flow_panel_factor_dispatch(W, ld, s, j0, nbk, rm, q, ls, fs, k % S, nload, pc);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("fence.proxy.async;\n" ::: "memory");
            flow_st_release(&flags[0], k + 1);
        }
    };
    if (me == 0) factor_and_publish(0);
    pc.mark(8);
#pragma unroll 1
    for (int k = 0; k < np; ++k) {
        int b = k + 1 + (((me - (k + 1)) % G) + G) % G;  // this CTA's first block beyond k
        const int j0 = k * kNB;
        const int nbk = nref - j0 < kNB ? nref - j0 : kNB;
        // a last panel of fewer than kNB columns (nref < ncols) leaves trailing columns in its own block
        const int own_end = j0 + kNB < s.ncols ? j0 + kNB : s.ncols;
        const bool own_rest = k % G == me && j0 + nbk < own_end;
        if (b >= nblk && !own_rest) {  // nothing left for this CTA: acknowledge every remaining panel and leave
            for (int kk = k + tid; kk < np; kk += kThreads) atomicAdd(&fin[kk], 1);
            break;
        }
        const RowMap rm = panel_rows(s, j0, j0 + nbk - 1);
        const int slot = k % S;
        flow_wait_ge(&flags[0], k + 1, b == k + 1);   // (the owner of the next panel is on the critical path)
        pc.mark(9);
        const double* Tk = q.Tg + (size_t)slot * kNB * kLdr;
        for (int idx = tid; idx < kNB * kLdr; idx += kThreads) ls.Ts[idx] = __ldcg(Tk + idx);
        __syncthreads();
        if (own_rest) flow_apply_dispatch<false>(W, ld, j0 + nbk, own_end, rm, nullptr, 0, ls.Ts, fs, &q.tmapV, slot * kNB, nload, pc);
        for (; b < nblk; b += G) {
            const int c0 = b * kNB, cend = c0 + kNB < s.ncols ? c0 + kNB : s.ncols;
            flow_apply_dispatch<false>(W, ld, c0, cend, rm, nullptr, 0, ls.Ts, fs, &q.tmapV, slot * kNB, nload, pc);
            pc.mark(10);
            if (b == k + 1 && b < np) {
                factor_and_publish(b);
                pc.mark(8);
            }
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(&fin[k], 1); }
    }
    __syncthreads();
    if (tid == 0) *ls.nload = nload;
    grid.sync();
    pc.mark(11);
}

// Once per kernel: the mbarriers of the TMA stages and the load counter.
__device__ __forceinline__ void flow_init_barriers(LargeSmem& ls, unsigned long long* bars, unsigned* nload) {
    ls.bars = bars;
    ls.nload = nload;
    if (threadIdx.x == 0) {
        for (int j = 0; j < kFlowMaxStages; ++j) flow_mbar_init(reinterpret_cast<uint64_t*>(&bars[j]));
        *nload = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
}

}  // namespace pnmol
