// Verify mma.sync.m8n8k4.f64 fragment layouts on sm_100a and measure dependent / independent issue cost.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void check(const double* A, const double* B, double* C, long long* cyc) {  // A 8x4 row-major, B 4x8 row-major, C 8x8
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, A[g * 4 + t], B[t * 8 + g]);
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1;
    double a = A[g * 4 + t] * 1e-3, b = B[t * 8 + g] * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 512; ++i) dmma(c0, c1, a, b);
    long long t1 = clock64();
    double d[8][2];
    for (int k = 0; k < 8; ++k) { d[k][0] = k; d[k][1] = -k; }
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < 512; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma(d[k][0], d[k][1], a, b);
    }
    long long t3 = clock64();
    double s = c0 + c1;
    for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1];
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
    C[64 + lane] = s;
}
int main() {
    double hA[32], hB[32], hC[96], ref[64];
    for (int i = 0; i < 32; ++i) { hA[i] = 1 + i * 0.5; hB[i] = 2 - i * 0.25; }
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += hA[i * 4 + k] * hB[k * 8 + j]; ref[i * 8 + j] = s; }
    double *A, *B, *C; long long* cyc; cudaMalloc(&A, 256); cudaMalloc(&B, 256); cudaMalloc(&C, 96 * 8); cudaMalloc(&cyc, 16);
    cudaMemcpy(A, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, 256, cudaMemcpyHostToDevice);
    check<<<1, 32>>>(A, B, C, cyc);
    long long h[2]; cudaMemcpy(hC, C, 96 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 64; ++i) err = fmax(err, fabs(hC[i] - ref[i]));
    printf("layout check max err %.3e (A[g][t], B[t][g], C[g][2t..2t+1])\n", err);
    printf("dependent DMMA %.1f cycles, 8 independent DMMA %.1f cycles per group (%.1f each)\n", h[0] / 512.0, h[1] / 512.0, h[1] / 4096.0);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
