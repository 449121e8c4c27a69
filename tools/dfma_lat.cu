// tools/dfma_lat.cu — development microbenchmark: latency of dependent FP64 instructions (one warp per SM)
#include <cuda_runtime.h>
#include <cstdio>
__global__ void chain(double* out, long long* cyc, int iters, double m) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a = fma(a, m, b);
    }
    long long t1 = clock64();
    double c = threadIdx.x * 2e-3, d = threadIdx.x * 3e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { c = fma(c, m, b); d = fma(d, m, b); }
    }
    long long t2 = clock64();
    double e[4] = {1e-3, 2e-3, 3e-3, 4e-3};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) { e[0] = fma(e[0], m, b); e[1] = fma(e[1], m, b); e[2] = fma(e[2], m, b); e[3] = fma(e[3], m, b); }
    }
    long long t3 = clock64();
    double s = a;
    long long t4 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) s = s + b;
    }
    long long t5 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c + d + e[0] + e[1] + e[2] + e[3] + s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t5 - t4; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 64);
    const int iters = 1000;
    for (int threads : {32, 64, 128, 256}) {
        chain<<<148, threads>>>(out, cyc, iters, 1.0000001); cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads/SM=%3d: cycles per DFMA: 1 chain %.1f, 2 chains %.1f, 4 chains %.1f; per dependent DADD %.1f\n", threads, h[0] / (16.0 * iters), h[1] / (16.0 * iters), h[2] / (16.0 * iters), h[3] / (16.0 * iters));
    }
    return 0;
}
