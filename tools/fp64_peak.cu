// FP64 pipe micro-benchmark for sm_100a: DFMA (CUDA-core) vs DMMA (mma.sync m8n8k4 f64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shuffle-reduction mixed with DFMA: models the warp-per-column Householder apply
__global__ void shfl_kernel(double* out, int iters) {
    double v = threadIdx.x * 1e-3;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        double w = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        acc = fma(w, 1e-9, acc);
        v = fma(v, 0.999, 1e-3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2, 4}) {
            if (threads * bps > 2048) continue;
            int blocks = p.multiProcessorCount * bps;
            int iters = 20000;
            float ms;
            dfma_kernel<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 8 * iters * (double)blocks * threads;
            printf("DFMA threads=%4d blocks/SM=%d : %8.3f ms  %8.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            dmma_kernel<<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            // per warp per mma: 8*8*4 FMAs = 512 flop
            fl = 512.0 * 8 * iters * (double)blocks * (threads / 32);
            printf("DMMA threads=%4d blocks/SM=%d : %8.3f ms  %8.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
        }
    }
    {
        int blocks = p.multiProcessorCount * 2, threads = 512, iters = 20000; float ms;
        shfl_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        shfl_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double red = (double)iters * blocks * (threads / 32);
        printf("SHFL-reduce(64b,5 stage): %8.3f ms  %8.2f G warp-reductions/s (%.2f clk/SM each @1.9GHz)\n", ms, red / ms * 1e-6,
               1.9e9 * (ms * 1e-3) / (red / p.multiProcessorCount));
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
