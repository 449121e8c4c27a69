// tools/tmem_micro.cu — development microbenchmark: tcgen05.st / tcgen05.ld (32x32b.x64) round-trip throughput when
// tensor memory is used as a spill space for per-thread accumulators (8 warps, 64 columns each pair of warps).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,"
                 "%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};"
                 :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),
                    "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]),
                    "r"(r[32]),"r"(r[33]),"r"(r[34]),"r"(r[35]),"r"(r[36]),"r"(r[37]),"r"(r[38]),"r"(r[39]),"r"(r[40]),"r"(r[41]),"r"(r[42]),"r"(r[43]),"r"(r[44]),"r"(r[45]),"r"(r[46]),"r"(r[47]),
                    "r"(r[48]),"r"(r[49]),"r"(r[50]),"r"(r[51]),"r"(r[52]),"r"(r[53]),"r"(r[54]),"r"(r[55]),"r"(r[56]),"r"(r[57]),"r"(r[58]),"r"(r[59]),"r"(r[60]),"r"(r[61]),"r"(r[62]),"r"(r[63]) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                   "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
                   "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
                   "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
                 : "r"(taddr));
}

template <int MODE>   // 0: st+ld round trip, 1: ld only, 2: st only
__global__ void __launch_bounds__(256) tmem_loop(int reps, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; i++) r[i] = threadIdx.x * 64 + i;
    tmem_st64(taddr, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t bad = 0;
    for (int it = 0; it < reps; it++) {
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 64; i++) r[i] += 1;
            tmem_st64(taddr, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (MODE != 2) {
            uint32_t q[64];
            tmem_ld64(taddr, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 64; i++) { bad |= (q[i] ^ r[i]); r[i] = q[i]; }
            } else {
#pragma unroll
                for (int i = 0; i < 64; i++) bad += q[i];
            }
        }
    }
    sink[blockIdx.x * 256 + threadIdx.x] = bad + r[5];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory");
}

template <int MODE> void run(const char* name, int reps, uint32_t* sink) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    tmem_loop<MODE><<<148, 256>>>(reps, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); tmem_loop<MODE><<<148, 256>>>(reps, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = ms * 1e-3 * 1.965e9 / reps;
    const double bytes = 256.0 * 64 * 4 * (MODE == 0 ? 2 : 1);
    uint32_t h[256]; CK(cudaMemcpy(h, sink, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-10s %8.3f ms  %7.1f cycles per iteration (64 KB per direction per CTA)  %6.1f B/cycle/SM   check=%u\n", name, ms, cyc, bytes / cyc, h[3]);
}
int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 20000;
    uint32_t* sink; CK(cudaMalloc(&sink, 148 * 256 * 4));
    run<0>("st+ld", reps, sink);
    run<1>("ld", reps, sink);
    run<2>("st", reps, sink);
    return 0;
}
