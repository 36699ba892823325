// Probe: cooperative launch + thread-block clusters + DSMEM on this GPU (used to design the cluster panel of qr_large).
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void probe(double* out, long long* cyc, int iters) {
    extern __shared__ double sm[];
    cg::grid_group grid = cg::this_grid();
    cg::cluster_group cl = cg::this_cluster();
    const unsigned r = cl.block_rank(), cs = cl.num_blocks();
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double* buf = sm + (it & 1) * 64 * 16;
        // all-gather: every CTA writes 16 doubles into slot r of every CTA of the cluster
        if (threadIdx.x < 16 * cs) {
            const unsigned dst = threadIdx.x / 16, k = threadIdx.x % 16;
            double* remote = cl.map_shared_rank(buf, dst);
            remote[r * 16 + k] = (double)(r + 1) * (k + 1) + it;
        }
        cl.sync();
        if (threadIdx.x < 16) {
            double s = 0.0;
            for (unsigned q = 0; q < cs; ++q) s += buf[q * 16 + threadIdx.x];
            acc += s;
        }
    }
    long long t1 = clock64();
    grid.sync();
    if (threadIdx.x < 16) out[blockIdx.x * 16 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    for (int cs : {2, 4, 8, 16}) {
        cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        const size_t smem = 200 * 1024;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        cfg.gridDim = dim3(cs);
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, probe, &cfg);
        printf("cluster %2d: max active clusters %d (%s) -> %d CTAs\n", cs, ncl, cudaGetErrorString(e), ncl * cs);
        if (ncl <= 0) continue;
        cfg.gridDim = dim3(ncl * cs);
        double* out; long long* cyc;
        cudaMalloc(&out, sizeof(double) * 16 * ncl * cs); cudaMalloc(&cyc, sizeof(long long) * ncl * cs);
        const int iters = 1000;
        e = cudaLaunchKernelEx(&cfg, probe, out, cyc, iters);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long c0 = 0; double o0 = 0;
        cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost); cudaMemcpy(&o0, out, sizeof(o0), cudaMemcpyDeviceToHost);
        printf("   launch %s / %s: %.0f cycles per all-gather+cluster.sync round, out[0]=%g\n", cudaGetErrorString(e), cudaGetErrorString(e2), (double)c0 / iters, o0);
        cudaFree(out); cudaFree(cyc);
    }
    return 0;
}
