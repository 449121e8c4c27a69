// tools/f2i_micro.cu — development microbenchmark: throughput of F2I.S64.F64 (double → int64) against DFMA, 8 warps per SM
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(long long* out, int iters, double m) {
    double a[8]; long long acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { acc += __double2ll_rn(a[i]); a[i] += 1.0; }
            else { a[i] = fma(a[i], m, 1.0); a[i] += 1.0; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (long long)s;
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 256>>>(out, iters, 1.0000001); else k<1><<<148, 256>>>(out, iters, 1.0000001);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        const double per_sm_clk = (double)iters * 8 * 256 / (best * 1e-3 * 1.965e9);
        printf("%s: %.3f ms  -> %.1f (F2I+DADD | DFMA+DADD) pairs per clock per SM\n", mode == 0 ? "F2I.S64.F64 + DADD" : "DFMA + DADD", best, per_sm_clk);
    }
    return 0;
}
