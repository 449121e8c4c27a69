// kernels_shape.inl — instantiates the EP kernels for one shape.  The including .cu defines TAC_N, TAC_K, TAC_SHAPE_FN and
// the list of PBS level counts via TAC_PBS_LEVELS(X).
#include "kernels_ep.cuh"
#include "shape_launch.h"

#include <algorithm>
#include <cstdlib>

namespace tac {
namespace {

constexpr int SN = TAC_N, SK = TAC_K;

#define TAC_SET_SMEM(kern, bytes)                                                                      \
    do {                                                                                               \
        cudaError_t e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
        if (e__ != cudaSuccess) return e__;                                                            \
    } while (0)

cudaError_t s_poly_fft(const KLaunch& k, const uint64_t* polys, size_t npoly, double scale, double2* out) {
    constexpr int M = SN / 2;
    const size_t smem = (size_t)(16 * M + tab_len(SN)) * sizeof(cplx);
    TAC_SET_SMEM(poly_fft_kernel<SN>, smem);
    poly_fft_kernel<SN><<<(unsigned)((npoly + 15) / 16), 256, smem, k.stream>>>(polys, npoly, scale, k.wT, out);
    return cudaGetLastError();
}

template <int L, int B, int NT, int MINB = 1, int DEPTH = 5, int NS = 0>
cudaError_t launch_pbs(const KLaunch& k, const uint64_t* small, int nct, int n, const double2* bsk, int base_log, uint64_t alpha, uint64_t* out) {
    typedef EpCfg<SN, SK, L, B> C;
    const size_t smem = PbsSmem<C, NS>::bytes;
    TAC_SET_SMEM((pbs_kernel<SN, SK, L, B, NT, MINB, DEPTH, NS>), smem);
    pbs_kernel<SN, SK, L, B, NT, MINB, DEPTH, NS><<<(unsigned)((nct + B - 1) / B), NT, smem, k.stream>>>(small, nct, n, bsk, base_log, alpha, k.wT, out);
    return cudaGetLastError();
}
template <int L, int DEPTH, int NS = 0>
cudaError_t launch_pbs_wide(const KLaunch& k, const uint64_t* small, int nct, int n, const double2* bsk, int base_log, uint64_t alpha, uint64_t* out) {
    typedef EpCfg<SN, SK, L, 1> C;
    const size_t smem = WideSmem<C, NS>::bytes;
    TAC_SET_SMEM((pbs_wide_kernel<SN, SK, L, 1, 256, DEPTH, NS>), smem);
    pbs_wide_kernel<SN, SK, L, 1, 256, DEPTH, NS><<<(unsigned)nct, 256, smem, k.stream>>>(small, nct, n, bsk, base_log, alpha, k.wT, out);
    return cudaGetLastError();
}
#if TAC_N == 512
template <int L, int DEPTH, int BLOG>
cudaError_t launch_pbs_merged(const KLaunch& k, const uint64_t* small, int nct, int n, const double2* bsk, int base_log, uint64_t alpha, uint64_t* out) {
    typedef EpCfg<SN, SK, L, 3> C;
    const size_t smem = MergedSmem<C>::bytes;
    TAC_SET_SMEM((pbs_merged_kernel<SN, SK, L, 3, 256, DEPTH, BLOG>), smem);
    pbs_merged_kernel<SN, SK, L, 3, 256, DEPTH, BLOG><<<(unsigned)((nct + 2) / 3), 256, smem, k.stream>>>(small, nct, n, bsk, base_log, alpha, k.wT, out);
    return cudaGetLastError();
}
#endif
template <int L>
cudaError_t pbs_levels(const KLaunch& k, const uint64_t* small, int nct, int n, const double2* bsk, int base_log, uint64_t alpha, uint64_t* out) {
#if TAC_N == 512
    // Two kernels, chosen by a wave-count cost model (times of one wave measured on B200, tools/pbs_bench.cu):
    //  * pbs_merged_kernel, 3 ciphertexts per CTA sharing every BSK load, the L levels of a step in one barrier interval
    //    (255 registers, no spills): 7.1 ms per wave of 3·SMs ciphertexts — the throughput configuration.  (pbs_kernel, the
    //    level-by-level schedule it replaces, needs 8.0 ms; it remains for level counts whose L buffers do not fit.)
    //  * pbs_wide_kernel, 1 ciphertext per CTA with the levels transformed in parallel: 3.9 ms per wave of SMs ciphertexts
    //    — the latency configuration for small batches (one AES block = 128 ciphertexts).
    const long waves3 = ((nct + 2) / 3 + k.sm_count - 1) / k.sm_count, waves1 = (nct + k.sm_count - 1) / k.sm_count;
    if (L >= 2 && waves1 * 39 <= waves3 * 71) return launch_pbs_wide<(L >= 2 ? L : 2), 3>(k, small, nct, n, bsk, base_log, alpha, out);
    if constexpr (L == 2 || L == 3) {
        // the shipped base log as a compile-time constant (its shifts and masks fold), any other at run time
        // (TAC_PBS_GENERIC_BASE_LOG in the environment forces the run-time form: the parity test compares the two)
        if (base_log == 12 && !getenv("TAC_PBS_GENERIC_BASE_LOG")) return launch_pbs_merged<L, 4, 12>(k, small, nct, n, bsk, base_log, alpha, out);
        return launch_pbs_merged<L, 4, 0>(k, small, nct, n, bsk, base_log, alpha, out);
    }
    return launch_pbs<L, 3, 256, 1, 4, 1>(k, small, nct, n, bsk, base_log, alpha, out);      // ring of 4 rows + 1 row staged by cp.async.bulk
#else
    return launch_pbs<L, 1, 128>(k, small, nct, n, bsk, base_log, alpha, out);      // test-only parameter sets: one instantiation
#endif
}
cudaError_t s_pbs(const KLaunch& k, int levels, const uint64_t* small, int nct, int n, const double2* bsk, int base_log, uint64_t alpha, uint64_t* out) {
#define X(LV) if (levels == LV) return pbs_levels<LV>(k, small, nct, n, bsk, base_log, alpha, out);
    TAC_PBS_LEVELS(X)
#undef X
    return cudaErrorInvalidValue;
}

#ifndef TAC_CBS_LEVELS
#define TAC_CBS_LEVELS(X) X(1) X(2)          // circuit-bootstrap level counts the vertical-packing kernels are instantiated for
#endif
template <int L, int B, int NT>
cudaError_t launch_vp(const KLaunch& k, const double2* ggsw_f, int nbox, int n_in, int first, const uint64_t* lut, size_t lut_stride,
                      const uint64_t* init_glwe, int n_out, int base_log, uint64_t* out) {
    typedef EpCfg<SN, SK, L, B> C;
    const size_t smem = EpSmem<C>::bytes;
    TAC_SET_SMEM((vp_kernel<SN, SK, L, B, NT>), smem);
    dim3 grid((n_out + B - 1) / B, nbox);
    vp_kernel<SN, SK, L, B, NT><<<grid, NT, smem, k.stream>>>(ggsw_f, n_in, first, lut, lut_stride, init_glwe, n_out, base_log, k.wT, out);
    return cudaGetLastError();
}
template <int L>
cudaError_t vp_levels(const KLaunch& k, const double2* ggsw_f, int nbox, int n_in, int first, const uint64_t* lut, size_t lut_stride,
                      const uint64_t* init_glwe, int n_out, int base_log, uint64_t* out) {
#if TAC_N == 512
    // 3 outputs per CTA with 256 threads (222 registers, no spills) measured 20 % faster on B200 than 4 outputs with 320
    // threads (168-register cap, spills) and 28 % faster than 2 outputs with 160 threads
    if (n_out >= 3) return launch_vp<L, 3, 256>(k, ggsw_f, nbox, n_in, first, lut, lut_stride, init_glwe, n_out, base_log, out);
    if (n_out == 2) return launch_vp<L, 2, 160>(k, ggsw_f, nbox, n_in, first, lut, lut_stride, init_glwe, n_out, base_log, out);
#endif
    return launch_vp<L, 1, 128>(k, ggsw_f, nbox, n_in, first, lut, lut_stride, init_glwe, n_out, base_log, out);
}
cudaError_t s_vp(const KLaunch& k, int levels, const double2* ggsw_f, int nbox, int n_in, int first, const uint64_t* lut, size_t lut_stride,
                 const uint64_t* init_glwe, int n_out, int base_log, uint64_t* out) {
#define X(LV) if (levels == LV) return vp_levels<LV>(k, ggsw_f, nbox, n_in, first, lut, lut_stride, init_glwe, n_out, base_log, out);
    TAC_CBS_LEVELS(X)
#undef X
    return cudaErrorInvalidValue;
}

template <int L>
cudaError_t tree_levels(const KLaunch& k, const double2* ggsw_f, int nbox, int n_in, int ggsw_idx, const uint64_t* lut, size_t lut_stride,
                        const uint64_t* node_in, int n_nodes_in, int n_out, int base_log, uint64_t* node_out) {
    typedef EpCfg<SN, SK, L, 1> C;
    const size_t smem = EpSmem<C>::bytes + (size_t)C::G * SN * sizeof(uint64_t);
    TAC_SET_SMEM((cmux_tree_kernel<SN, SK, L, 128>), smem);
    dim3 grid(n_nodes_in / 2, n_out, nbox);
    cmux_tree_kernel<SN, SK, L, 128><<<grid, 128, smem, k.stream>>>(ggsw_f, n_in, ggsw_idx, lut, lut_stride, node_in, n_nodes_in, n_out, base_log, k.wT, node_out);
    return cudaGetLastError();
}
cudaError_t s_tree(const KLaunch& k, int levels, const double2* ggsw_f, int nbox, int n_in, int ggsw_idx, const uint64_t* lut, size_t lut_stride,
                   const uint64_t* node_in, int n_nodes_in, int n_out, int base_log, uint64_t* node_out) {
#define X(LV) if (levels == LV) return tree_levels<LV>(k, ggsw_f, nbox, n_in, ggsw_idx, lut, lut_stride, node_in, n_nodes_in, n_out, base_log, node_out);
    TAC_CBS_LEVELS(X)
#undef X
    return cudaErrorInvalidValue;
}

// one CMux-with-rotation step per accumulator (B = 1 CTA each) — test entry point for the FFT / external-product core
template <int L, int NT>
__global__ void __launch_bounds__(NT, 1)
cmux_rotate_test_kernel(const cplx* __restrict__ ggsw_f, const int* __restrict__ rot, int base_log, const cplx* __restrict__ g_wT,
                        uint64_t* __restrict__ acc_io) {
    typedef EpCfg<SN, SK, L, 1> C;
    typedef MacCfg<C, NT> MC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EpSmem<C> sm(smem_raw);
    const int tid = threadIdx.x;
    uint64_t* g = acc_io + (size_t)blockIdx.x * C::G * SN;
    for (int i = tid; i < tab_len(C::N); i += NT) sm.wT[i] = g_wT[i];
    for (int i = tid; i < C::G * SN; i += NT) sm.acc[i] = g[i];
    __syncthreads();
    cplx outr[MC::SPT][1][C::G];
#pragma unroll
    for (int a = 0; a < MC::SPT; a++)
#pragma unroll
        for (int c = 0; c < C::G; c++) outr[a][0][c] = mk(0.0, 0.0);
    const int r = rot[blockIdx.x];
    ep_step_device<C, NT>(tid, sm, ggsw_f, [&](int job, int jj, uint64_t& x0, uint64_t& x1) { rot_diff_pair<SN>(sm.acc + (size_t)job * SN, jj, r, x0, x1); }, base_log, outr);
    __syncthreads();
    for (int i = tid; i < C::G * SN; i += NT) g[i] = sm.acc[i];
}
template <int L>
cudaError_t launch_cmux_test(const KLaunch& k, const double2* gf, const int* rot, int base_log, int n_acc, uint64_t* acc) {
    typedef EpCfg<SN, SK, L, 1> C;
    const size_t smem = EpSmem<C>::bytes;
    TAC_SET_SMEM((cmux_rotate_test_kernel<L, 128>), smem);
    cmux_rotate_test_kernel<L, 128><<<n_acc, 128, smem, k.stream>>>(gf, rot, base_log, k.wT, acc);
    return cudaGetLastError();
}
cudaError_t s_cmux_test(const KLaunch& k, int levels, const double2* gf, const int* rot, int base_log, int n_acc, uint64_t* acc) {
    if (levels == 1) return launch_cmux_test<1>(k, gf, rot, base_log, n_acc, acc);
#define X(LV) if (levels == LV) return launch_cmux_test<LV>(k, gf, rot, base_log, n_acc, acc);
    TAC_PBS_LEVELS(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t s_sample_extract(const KLaunch& k, const uint64_t* glwe, size_t n, uint64_t* out) {
    const size_t total = n * ((size_t)SK * SN + 1);
    const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)k.sm_count * 16);
    sample_extract_kernel<SN, SK><<<grid ? grid : 1, 256, 0, k.stream>>>(glwe, n, out);
    return cudaGetLastError();
}

const ShapeOps kOps = {SN, SK, s_poly_fft, s_pbs, s_vp, s_tree, s_cmux_test, s_sample_extract};

}  // namespace

const ShapeOps* TAC_SHAPE_FN() { return &kOps; }

}  // namespace tac
