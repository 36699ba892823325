// Dependent-issue latencies on sm_100a (single warp, clock64): DFMA, DADD, DMUL, SHFL(64-bit), LDS.64, rsqrt, rcp, sqrt, div.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (double)((i * 7) % 1024);
    __syncwarp();
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    int idx = 0;
#define TIME(slot, body) t0 = clock64(); _Pragma("unroll 1") for (int i = 0; i < N; ++i) { body; } t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0);
    TIME(0, x = fma(x, y, z))
    TIME(1, x = x + z)
    TIME(2, x = x * y)
    TIME(3, x = __shfl_xor_sync(0xffffffffu, x, 1))
    TIME(4, x = x + __shfl_xor_sync(0xffffffffu, x, 1))
    TIME(5, { idx = (int)sm[idx & 1023]; })
    TIME(6, x = rsqrt(x + 2.0))
    TIME(7, x = __drcp_rn(x + 2.0))
    TIME(8, x = sqrt(x + 2.0))
    TIME(9, x = 1.0 / (x + 2.0))
    // independent DFMA throughput for one warp: 8 chains
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    TIME(10, { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); a4 = fma(a4, y, z); a5 = fma(a5, y, z); a6 = fma(a6, y, z); a7 = fma(a7, y, z); })
    double v = sm[threadIdx.x];
    TIME(11, { v = sm[(threadIdx.x + i) & 1023] + v; })
    out[threadIdx.x] = x + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + idx + v;
}
__global__ void bar_probe(long long* cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[12 + blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 32 * 8); cudaMemset(cyc, 0, 32 * 8);
    probe<<<1, 32>>>(out, cyc, 1.5); probe<<<1, 32>>>(out, cyc, 1.5);
    bar_probe<<<1, 256>>>(cyc);
    long long h[32]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA dependent", "DADD dependent", "DMUL dependent", "SHFL.64 dependent", "SHFL.64 + DADD", "LDS pointer chase (+cvt)", "rsqrt(double)", "__drcp_rn", "sqrt(double)", "1.0/x (double)", "8 independent DFMA (per group of 8)", "LDS.64 + DADD dependent", "__syncthreads (256 thr, alone)"};
    for (int i = 0; i < 13; ++i) printf("%-38s %8.1f cycles\n", names[i], (double)h[i] / N);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
