// tools/umma_gemm_test.cu — development test + timing of lwe_gemm_umma_kernel against a CPU integer GEMM.
//   ./umma_gemm_test [nct] [Kd] [W] [nkeys] [nlimb] [check 0/1]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "kernels_gemm_umma.cuh"
using namespace tac;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// direct tile builder from explicit digits (bypasses the decomposition): d [nct][Kd] values < 2^(8·nlimb)
__global__ void tiles_from_digits(const uint32_t* d, int nct, int mpad, int Kd, int nkb, int nlimb, uint8_t* DA) {
    const size_t total = (size_t)mpad * nkb * 2;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ct = (int)(idx % mpad); const size_t q = idx / mpad; const int khalf = (int)(q & 1), kb = (int)(q >> 1);
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
        if (ct < nct) for (int kq = 0; kq < 16; kq++) {
            const int k = kb * UG_KB + khalf * 16 + kq; if (k >= Kd) continue;
            const uint32_t dp = d[(size_t)ct * Kd + k];
            lo[kq >> 2] |= (dp & 0xFFu) << (8 * (kq & 3)); hi[kq >> 2] |= ((dp >> 8) & 0xFFu) << (8 * (kq & 3));
        }
        const int mt = ct / UG_MT, row = ct - mt * UG_MT;
        uint8_t* tile = DA + ((size_t)mt * nkb + kb) * (size_t)(nlimb * 2 * UG_MT * 16);
        *reinterpret_cast<uint4*>(tile + ((size_t)(0 * 2 + khalf) * UG_MT + row) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (nlimb == 2) *reinterpret_cast<uint4*>(tile + ((size_t)(1 * 2 + khalf) * UG_MT + row) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
}

int main(int argc, char** argv) {
    const int nct = argc > 1 ? atoi(argv[1]) : 200, Kd = argc > 2 ? atoi(argv[2]) : 100, W = argc > 3 ? atoi(argv[3]) : 128, nkeys = argc > 4 ? atoi(argv[4]) : 2;
    const int nlimb = argc > 5 ? atoi(argv[5]) : 2, check = argc > 6 ? atoi(argv[6]) : 1;
    const int nkb = (Kd + UG_KB - 1) / UG_KB, mtiles = (nct + UG_MT - 1) / UG_MT, mpad = mtiles * UG_MT, ntiles = (W + UG_NT - 1) / UG_NT;
    std::mt19937_64 rng(11);
    std::vector<uint32_t> d((size_t)nct * Kd);
    for (auto& v : d) v = (uint32_t)(rng() & (nlimb == 2 ? 0xFFFFu : 0xFFu));
    std::vector<uint64_t> key((size_t)nkeys * Kd * W), corr((size_t)nkeys * W);
    for (auto& v : key) v = rng();
    for (auto& v : corr) v = rng();
    uint32_t* d_d; uint64_t *d_key, *d_corr, *d_out; uint8_t *DA, *KP;
    CK(cudaMalloc(&d_d, d.size() * 4)); CK(cudaMemcpy(d_d, d.data(), d.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_key, key.size() * 8)); CK(cudaMemcpy(d_key, key.data(), key.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_corr, corr.size() * 8)); CK(cudaMemcpy(d_corr, corr.data(), corr.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)nct * nkeys * W * 8)); CK(cudaMemset(d_out, 0xEE, (size_t)nct * nkeys * W * 8));
    const size_t da_bytes = (size_t)mtiles * nkb * nlimb * 2 * UG_MT * 16, kp_bytes = (size_t)nkeys * ntiles * nkb * UG_B_BYTES;
    CK(cudaMalloc(&DA, da_bytes)); CK(cudaMalloc(&KP, kp_bytes));
    tiles_from_digits<<<1024, 256>>>(d_d, nct, mpad, Kd, nkb, nlimb, DA);
    umma_key_tiles_kernel<<<2048, 256>>>(d_key, nkeys, Kd, W, nkb, KP);
    CK(cudaDeviceSynchronize());
    const int grid = mtiles * ntiles * nkeys;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        if (nlimb == 2) {
            CK(cudaFuncSetAttribute(lwe_gemm_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UgCfg<2>::SMEM));
            lwe_gemm_umma_kernel<2><<<grid, UG_THREADS, UgCfg<2>::SMEM>>>(DA, nct, mtiles, KP, W, nkeys, nkb, d_corr, nullptr, 0, d_out);
        } else {
            CK(cudaFuncSetAttribute(lwe_gemm_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UgCfg<1>::SMEM));
            lwe_gemm_umma_kernel<1><<<grid, UG_THREADS, UgCfg<1>::SMEM>>>(DA, nct, mtiles, KP, W, nkeys, nkb, d_corr, nullptr, 0, d_out);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
    }
    const double macs = (double)nct * Kd * W * nkeys * (nlimb == 2 ? 15 : 8);
    printf("nct=%d Kd=%d W=%d nkeys=%d nlimb=%d grid=%d: %.3f ms, %.1f u8 TOP/s\n", nct, Kd, W, nkeys, nlimb, grid, best, 2 * macs / (best * 1e-3) / 1e12);
    if (check) {
        std::vector<uint64_t> out((size_t)nct * nkeys * W);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        const size_t total_e = (size_t)nct * nkeys * W, nsample = check == 2 ? 4000 : total_e;
        for (size_t e = 0; e < nsample; e++) {
            const size_t id = check == 2 ? (size_t)(rng() % total_e) : e;
            const int c = (int)(id % W), j = (int)((id / W) % nkeys), ct = (int)(id / ((size_t)W * nkeys));
            uint64_t s = 0;
            for (int k = 0; k < Kd; k++) s += (uint64_t)d[(size_t)ct * Kd + k] * key[((size_t)j * Kd + k) * W + c];
            const uint64_t want = corr[(size_t)j * W + c] - s, got = out[((size_t)ct * nkeys + j) * W + c];
            if (want != got) { if (bad < 5) printf("mismatch ct=%d j=%d col=%d want=%016llx got=%016llx diff=%016llx\n", ct, j, c, (unsigned long long)want, (unsigned long long)got, (unsigned long long)(want - got)); bad++; }
        }
        printf("%s: %zu mismatches of %zu\n", bad ? "FAIL" : "OK", bad, out.size());
        return bad ? 1 : 0;
    }
    return 0;
}
