// tac_common.h — parameter block and small host/device integer helpers shared by every kernel.
//
// Torus = uint64_t with wrapping arithmetic (q = 2^64).  Layouts are documented in include/tfhe_aes_cuda.h.
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define TAC_HD __host__ __device__ __forceinline__
#else
#define TAC_HD inline
#endif

// Mirrors WopbsParameters + max_noise_level_squared (reference src/tfhe/shortint_woppbs_1bit/parameters.rs:9-13).
// Same field order as `tac_params` in include/tfhe_aes_cuda.h.
struct TacParams {
    int32_t n, k, N;
    int32_t pbs_l, pbs_b;
    int32_t ks_l, ks_b;
    int32_t cbs_l, cbs_b;
    int32_t pfks_l, pfks_b;
    int32_t max_noise_sq;
    double s_lwe, s_glwe, s_pfks;
};

namespace tac {

// ---- signed decomposition, bit-exact with tfhe 0.11.2 SignedDecomposer (see oracle/oracle.cpp for the restatement notes)
TAC_HD uint64_t closest_representable(uint64_t x, int b, int l) {
    const int non_rep = 64 - b * l;
    const uint64_t msb = (x >> (non_rep - 1)) & 1ull;
    return ((x >> non_rep) + msb) << non_rep;
}
TAC_HD uint64_t decomp_init_state(uint64_t x, int b, int l) {
    const int rep = b * l;
    uint64_t res = x >> (64 - rep - 1);
    const uint64_t rounding_bit = res & 1ull;
    res = ((res + 1ull) >> 1) & ((~0ull) >> (64 - rep));
    const uint64_t half = 1ull << (rep - 1);
    const uint64_t need_balance = (res > half || (res == half && rounding_bit == 1ull)) ? 1ull : 0ull;
    return res - (need_balance << rep);
}
// returns the signed digit of the current lowest level and advances the state
TAC_HD int64_t decomp_next(uint64_t& state, int b) {
    const uint64_t mask = (1ull << b) - 1ull;
    const uint64_t res = state & mask;
    state >>= b;
    uint64_t carry = ((res - 1ull) | state) & res;
    carry >>= (b - 1);
    state += carry;
    return (int64_t)(res - (carry << b));
}
// modulus switch to Z_{2N} for blind rotation ([U] pbs_modulus_switch)
TAC_HD int modswitch(uint64_t a, int logN) {
    const int lg = logN + 1;
    return (int)((a + (1ull << (64 - lg - 1))) >> (64 - lg));
}
TAC_HD int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }

}  // namespace tac
