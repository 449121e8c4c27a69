// kernels.cuh — sm_100a kernels of the WoP-PBS hot path (SURVEY.md §8a rows a3, a6, a9–a15, a17).
//
//   lwe_gemm_kernel        K1/K5  batched LWE keyswitch and private functional packing keyswitch as an exact integer
//                                 GEMM mod 2^64:  out[ct][col] = corr[col] − Σ_k digit'[ct][k]·key[k][col]
//   pbs_kernel             K4     persistent blind rotation + sample extract (homomorphic_shift_boolean), B ciphertexts per
//                                 CTA share every BSK load
//   poly_fft_kernel        K2/K6  torus polynomial → Fourier slots (BSK conversion, GGSW fill_with_forward_fourier)
//   vp_kernel              K7     vertical packing: blind rotation by the circuit-bootstrapped GGSWs + sample extract
//   cmux_tree_kernel       K7     one CMux-tree layer (only when n_in > log2 N)
//   aes_*_kernel           K8     AddRoundKey / ShiftRows+MixColumns / final round as gather-adds on the flat state
#pragma once
#include <cuda_runtime.h>
#include "ep_step.cuh"

namespace tac {

template <int N> struct LogN { static constexpr int v = (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : (N == 2048) ? 11 : -1; };

// ================================================================================================ digits
// bit-exact signed decomposition (tfhe SignedDecomposer), stored with the offset B/2 so that digits are non-negative:
// digit' = digit + B/2 ∈ [0, B].  Key order of the keyswitch key: block i holds level l first ([U] lwe_keyswitch.rs).
__global__ void ks_digits_kernel(const uint64_t* __restrict__ in, int nct, int big, int b, int l, uint32_t* __restrict__ dig) {
    const size_t total = (size_t)nct * big;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t ct = idx / big; const int i = (int)(idx - ct * big);
        uint64_t st = decomp_init_state(in[ct * (big + 1) + i], b, l);
        uint32_t* o = dig + (ct * big + i) * l;
        for (int s = 0; s < l; s++) o[s] = (uint32_t)(decomp_next(st, b) + (int64_t)(1u << (b - 1)));
    }
}
// PFKS: closest_representable first, all big+1 elements (mask and body), key block stores level 1 first and is iterated
// reversed ([U] lwe_private_functional_packing_keyswitch.rs).
__global__ void pfks_digits_kernel(const uint64_t* __restrict__ in, int nct, int big1, int b, int l, uint32_t* __restrict__ dig) {
    const size_t total = (size_t)nct * big1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        uint64_t st = decomp_init_state(closest_representable(in[idx], b, l), b, l);
        uint32_t* o = dig + idx * l;
        for (int lev = l; lev >= 1; lev--) o[lev - 1] = (uint32_t)(decomp_next(st, b) + (int64_t)(1u << (b - 1)));
    }
}

// ================================================================================================ integer GEMM mod 2^64
// out[ct][j][col] = corr[j][col] − Σ_k dig[ct][k] · key[j][k][col]   (+ last_col_add[ct·stride] on col == W-1)
// CTA tile: (4·RB ciphertexts) × 128 columns; thread: RB ciphertexts × 2 adjacent columns; K chunked by 32 through smem.
// The product u32 × u64 → low 64 bits is one IMAD.WIDE.U32 (low limb, 64-bit accumulate) plus one IMAD (high limb).
template <int RB>
__global__ void __launch_bounds__(256)
lwe_gemm_kernel(const uint32_t* __restrict__ dig, int nct, int Kd, const uint64_t* __restrict__ key, int W, int nkeys,
                const uint64_t* __restrict__ corr, const uint64_t* __restrict__ last_col_add, size_t add_stride,
                uint64_t* __restrict__ out) {
    constexpr int TB = 4 * RB, KC = 32, TN = 128;
    __shared__ __align__(16) uint32_t dsm[TB][KC];
    const int tiles_per_key = (W + TN - 1) / TN;
    const int j = blockIdx.x / tiles_per_key;
    const int col = (blockIdx.x - j * tiles_per_key) * TN + 2 * (threadIdx.x & 63);
    const int ty = threadIdx.x >> 6;
    const int ct0 = blockIdx.y * TB;
    const bool col_ok = col < W;           // W is even, so col+1 < W as well
    uint64_t acc[RB][2];
#pragma unroll
    for (int r = 0; r < RB; r++) { acc[r][0] = 0; acc[r][1] = 0; }
    const uint64_t* kbase = key + (size_t)j * Kd * W + (col_ok ? col : 0);
    for (int k0 = 0; k0 < Kd; k0 += KC) {
        for (int e = threadIdx.x; e < TB * KC; e += 256) {
            const int r = e / KC, kk = e - r * KC;
            const int ct = ct0 + r, k = k0 + kk;
            dsm[r][kk] = (ct < nct && k < Kd) ? dig[(size_t)ct * Kd + k] : 0u;
        }
        __syncthreads();
        const int kmax = min(KC, Kd - k0);
        if (col_ok) {
#pragma unroll 4
            for (int kk = 0; kk < kmax; kk++) {
                const ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2*>(kbase + (size_t)(k0 + kk) * W));
                const uint32_t k0lo = (uint32_t)kv.x, k0hi = (uint32_t)(kv.x >> 32), k1lo = (uint32_t)kv.y, k1hi = (uint32_t)(kv.y >> 32);
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    const uint32_t d = dsm[ty * RB + r][kk];
                    acc[r][0] += (uint64_t)d * k0lo; acc[r][0] += (uint64_t)(d * k0hi) << 32;
                    acc[r][1] += (uint64_t)d * k1lo; acc[r][1] += (uint64_t)(d * k1hi) << 32;
                }
            }
        }
        __syncthreads();
    }
    if (!col_ok) return;
    const uint64_t c0 = corr ? corr[(size_t)j * W + col] : 0ull, c1 = corr ? corr[(size_t)j * W + col + 1] : 0ull;
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const int ct = ct0 + ty * RB + r;
        if (ct >= nct) continue;
        uint64_t v0 = c0 - acc[r][0], v1 = c1 - acc[r][1];
        if (last_col_add && col + 1 == W - 1) v1 += last_col_add[(size_t)ct * add_stride];   // W is even: the last column is a v1
        uint64_t* o = out + ((size_t)ct * nkeys + j) * W + col;
        *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(v0, v1);
    }
}

// ================================================================================================ Fourier transform of torus polynomials
// 16 polynomials per CTA (one per 16-thread group).  out[poly][M] in slot order, scaled by `scale`·2^-64.
template <int N>
__global__ void __launch_bounds__(256)
poly_fft_kernel(const uint64_t* __restrict__ polys, size_t npoly, double scale, const cplx* __restrict__ g_twist,
                const cplx* __restrict__ g_wM, cplx* __restrict__ out) {
    constexpr int M = N / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* S = reinterpret_cast<cplx*>(smem_raw);
    cplx* twist = S + 16 * M;
    cplx* wM = twist + M;
    for (int i = threadIdx.x; i < M; i += 256) { twist[i] = g_twist[i]; wM[i] = g_wM[i]; }
    __syncthreads();
    const int grp = threadIdx.x >> 4, t = threadIdx.x & 15;
    const size_t poly = (size_t)blockIdx.x * 16 + grp;
    if (poly < npoly) key_fft_pass1<N>(t, polys + poly * N, scale, twist, wM, S + grp * M);
    __syncthreads();
    if (poly < npoly) fft_fwd_pass2<N>(t, S + grp * M);
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * 16 * M;
    const size_t lim = npoly * M;
    for (int i = threadIdx.x; i < 16 * M; i += 256)
        if (base + i < lim) out[base + i] = S[i];
}

// ================================================================================================ CMux chain plumbing
template <class C>
struct EpSmem {
    uint64_t* acc; cplx* S; cplx* twist; cplx* wM; unsigned char* extra;
    __device__ explicit EpSmem(unsigned char* raw) {
        acc = reinterpret_cast<uint64_t*>(raw);
        S = reinterpret_cast<cplx*>(acc + C::acc_words);
        twist = S + C::s_cplx;
        wM = twist + C::M;
        extra = reinterpret_cast<unsigned char*>(wM + C::M);
    }
    static constexpr size_t bytes = C::acc_words * 8 + C::s_cplx * 16 + 2 * (size_t)C::M * 16;
};

template <class C, int NT, class RotFn>
__device__ __forceinline__ void ep_step_device(int tid, const EpSmem<C>& sm, const cplx* __restrict__ ggsw, RotFn rotf, const DecompF64& dc,
                                               cplx (&out)[MacCfg<C, NT>::SPT][C::B][C::G]) {
    typedef MacCfg<C, NT> MC;
#pragma unroll
    for (int lev = C::L; lev >= 1; lev--) {
        ph_fwd1<C>(tid, NT, lev, sm.acc, rotf, dc, sm.twist, sm.wM, sm.S);
        __syncthreads();
        ph_fwd2<C>(tid, NT, sm.S);
        __syncthreads();
        ph_mac<C, MC::NT_MAC, MC::SPT>(tid, lev, ggsw, sm.S, out);
        __syncthreads();
    }
    ph_outw<C, MC::NT_MAC, MC::SPT>(tid, sm.S, out);
    __syncthreads();
    ph_inv1<C>(tid, NT, sm.wM, sm.S);
    __syncthreads();
    ph_inv2<C>(tid, NT, sm.twist, sm.S, sm.acc);
    __syncthreads();
}

// [U] glwe_sample_extraction.rs::extract_lwe_sample_from_glwe_ciphertext(.., MonomialDegree(0)); element e of the LWE
template <class C>
__device__ __forceinline__ uint64_t sample_extract_elem(const uint64_t* __restrict__ glwe, int e) {
    if (e == C::K * C::N) return glwe[(size_t)C::K * C::N];
    const int p = e / C::N, j = e - p * C::N;
    const uint64_t* a = glwe + (size_t)p * C::N;
    return (j == 0) ? a[0] : (0ull - a[C::N - j]);
}

// ================================================================================================ PBS (homomorphic_shift_boolean)
// in: small LWE [nct][n+1]; out: big LWE [nct][kN+1] encrypting bit·2·alpha.
// [U] wop_pbs.rs::homomorphic_shift_boolean + bootstrap.rs::{blind_rotate_assign, bootstrap}
template <int N, int K, int L, int B, int NT>
__global__ void __launch_bounds__(NT, 1)
pbs_kernel(const uint64_t* __restrict__ lwe_small, int nct, int n, const cplx* __restrict__ bsk, int base_log, uint64_t alpha,
           const cplx* __restrict__ g_twist, const cplx* __restrict__ g_wM, uint64_t* __restrict__ out_big) {
    typedef EpCfg<N, K, L, B> C;
    typedef MacCfg<C, NT> MC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EpSmem<C> sm(smem_raw);
    uint16_t* rots = reinterpret_cast<uint16_t*>(sm.extra);       // [B][n+1] monomial degrees
    const int tid = threadIdx.x;
    const int ct0 = blockIdx.x * B;
    const int n1 = n + 1;
    for (int i = tid; i < C::M; i += NT) { sm.twist[i] = g_twist[i]; sm.wM[i] = g_wM[i]; }
    for (int idx = tid; idx < B * n1; idx += NT) {
        const int b = idx / n1, i = idx - b * n1, ct = ct0 + b;
        uint64_t a = 0;
        if (ct < nct) {
            a = lwe_small[(size_t)ct * n1 + i];
            if (i == n) a += (1ull << 62);                         // centre the error for the negacyclic LUT
        }
        rots[idx] = (uint16_t)modswitch(a, LogN<N>::v);
    }
    __syncthreads();
    // accumulator = trivial GLWE(-alpha in every coefficient) · X^{-b~}
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int b = idx / (C::G * N), rem = idx - b * C::G * N, p = rem / N, j = rem - p * N;
        uint64_t v = 0;
        if (p == K) {
            const int s = (j + (int)rots[b * n1 + n]) & (2 * N - 1);
            v = (s < N) ? (0ull - alpha) : alpha;
        }
        sm.acc[idx] = v;
    }
    __syncthreads();
    cplx out[MC::SPT][B][C::G];
#pragma unroll
    for (int a = 0; a < MC::SPT; a++)
#pragma unroll
        for (int b = 0; b < B; b++)
#pragma unroll
            for (int c = 0; c < C::G; c++) out[a][b][c] = mk(0.0, 0.0);
    const DecompF64 dc = make_decomp(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    for (int i = 0; i < n; i++) {
        ep_step_device<C, NT>(tid, sm, bsk + ggsw_sz * i, [&](int b) { return (int)rots[b * n1 + i]; }, dc, out);
    }
    constexpr int LW = K * N + 1;
    for (int idx = tid; idx < B * LW; idx += NT) {
        const int b = idx / LW, e = idx - b * LW, ct = ct0 + b;
        if (ct >= nct) continue;
        uint64_t v = sample_extract_elem<C>(sm.acc + (size_t)b * C::G * N, e);
        if (e == K * N) v += alpha;
        out_big[(size_t)ct * LW + e] = v;
    }
}

// ================================================================================================ vertical packing
// One CTA evaluates B outputs of one box (= one circuit_bootstrap call).  ggsw_f: [nbox][n_in][L][G][G][M].
// The accumulator starts from init_glwe (CMux-tree result) when given, else from the trivial GLWE of LUT polynomial o.
// Blind rotation uses GGSWs n_in-1 … first_ggsw with X^{-1}, X^{-2}, X^{-4}, …   ([U] wop_pbs.rs::{vertical_packing, blind_rotate_assign})
template <int N, int K, int L, int B, int NT>
__global__ void __launch_bounds__(NT, 1)
vp_kernel(const cplx* __restrict__ ggsw_f, int n_in, int first_ggsw, const uint64_t* __restrict__ lut, size_t lut_stride,
          const uint64_t* __restrict__ init_glwe, int n_out, int base_log, const cplx* __restrict__ g_twist,
          const cplx* __restrict__ g_wM, uint64_t* __restrict__ out) {
    typedef EpCfg<N, K, L, B> C;
    typedef MacCfg<C, NT> MC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EpSmem<C> sm(smem_raw);
    const int tid = threadIdx.x;
    const int box = blockIdx.y, o0 = blockIdx.x * B;
    for (int i = tid; i < C::M; i += NT) { sm.twist[i] = g_twist[i]; sm.wM[i] = g_wM[i]; }
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int b = idx / (C::G * N), rem = idx - b * C::G * N, p = rem / N, j = rem - p * N;
        const int o = o0 + b;
        uint64_t v = 0;
        if (o < n_out) {
            if (init_glwe) v = init_glwe[((size_t)box * n_out + o) * C::G * N + rem];
            else if (p == K) v = lut[(size_t)o * lut_stride + j];
        }
        sm.acc[idx] = v;
    }
    __syncthreads();
    cplx outr[MC::SPT][B][C::G];
#pragma unroll
    for (int a = 0; a < MC::SPT; a++)
#pragma unroll
        for (int b = 0; b < B; b++)
#pragma unroll
            for (int c = 0; c < C::G; c++) outr[a][b][c] = mk(0.0, 0.0);
    const DecompF64 dc = make_decomp(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    const cplx* gbox = ggsw_f + (size_t)box * n_in * ggsw_sz;
    int deg = 1;
    for (int g = n_in - 1; g >= first_ggsw; g--) {
        const int rot = 2 * N - deg;            // multiply by X^{-deg}
        ep_step_device<C, NT>(tid, sm, gbox + (size_t)g * ggsw_sz, [&](int) { return rot; }, dc, outr);
        deg <<= 1;
    }
    constexpr int LW = K * N + 1;
    for (int idx = tid; idx < B * LW; idx += NT) {
        const int b = idx / LW, e = idx - b * LW, o = o0 + b;
        if (o >= n_out) continue;
        out[((size_t)box * n_out + o) * LW + e] = sample_extract_elem<C>(sm.acc + (size_t)b * C::G * N, e);
    }
}

// One CMux-tree layer ([U] wop_pbs.rs::cmux_tree_memory_optimized, evaluated level by level): for every (box, output,
// pair i): node_out[i] = c0 + G ⊡ (c1 − c0) with c0 = node_in[2i], c1 = node_in[2i+1].  leaf != 0: inputs are LUT
// polynomials (trivial GLWEs).  Implemented with the rotation step on a doubled trick: acc = c0, "rot" disabled — the
// difference c1 − c0 is written to a scratch accumulator instead.  One CTA per node (B = 1).
template <int N, int K, int L, int NT>
__global__ void __launch_bounds__(NT, 1)
cmux_tree_kernel(const cplx* __restrict__ ggsw_f, int n_in, int ggsw_idx, const uint64_t* __restrict__ lut, size_t lut_stride,
                 const uint64_t* __restrict__ node_in, int n_nodes_in, int n_out, int base_log, const cplx* __restrict__ g_twist,
                 const cplx* __restrict__ g_wM, uint64_t* __restrict__ node_out) {
    typedef EpCfg<N, K, L, 1> C;
    typedef MacCfg<C, NT> MC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EpSmem<C> sm(smem_raw);
    uint64_t* diff = reinterpret_cast<uint64_t*>(sm.extra);       // [G][N]  c1 − c0
    const int tid = threadIdx.x;
    const int pair = blockIdx.x, o = blockIdx.y, box = blockIdx.z;
    const int n_pairs = n_nodes_in / 2;
    for (int i = tid; i < C::M; i += NT) { sm.twist[i] = g_twist[i]; sm.wM[i] = g_wM[i]; }
    for (int idx = tid; idx < C::G * N; idx += NT) {
        uint64_t c0, c1;
        if (node_in) {
            const uint64_t* base = node_in + (((size_t)box * n_out + o) * n_nodes_in + 2 * pair) * C::G * N;
            c0 = base[idx]; c1 = base[(size_t)C::G * N + idx];
        } else {
            const int p = idx / N, j = idx - p * N;
            const uint64_t* lp = lut + (size_t)o * lut_stride + (size_t)(2 * pair) * N;
            c0 = (p == K) ? lp[j] : 0ull; c1 = (p == K) ? lp[N + j] : 0ull;
        }
        sm.acc[idx] = c0; diff[idx] = c1 - c0;
    }
    __syncthreads();
    cplx outr[MC::SPT][1][C::G];
#pragma unroll
    for (int a = 0; a < MC::SPT; a++)
#pragma unroll
        for (int c = 0; c < C::G; c++) outr[a][0][c] = mk(0.0, 0.0);
    const DecompF64 dc = make_decomp(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    const cplx* ggsw = ggsw_f + ((size_t)box * n_in + ggsw_idx) * ggsw_sz;
    // same phases as ep_step_device, but the decomposed operand is `diff` instead of a rotation difference
    const int grp = tid >> 4, t = tid & 15, ngrp = NT >> 4;
#pragma unroll
    for (int lev = L; lev >= 1; lev--) {
        for (int job = grp; job < C::G; job += ngrp) {
            const uint64_t* poly = diff + (size_t)job * N;
            fft_fwd_pass1<N>(t, [&](int j) { return digit_f64<L>(poly[j], dc, lev); }, sm.twist, sm.wM, sm.S + (size_t)job * C::M);
        }
        __syncthreads();
        ph_fwd2<C>(tid, NT, sm.S);
        __syncthreads();
        ph_mac<C, MC::NT_MAC, MC::SPT>(tid, lev, ggsw, sm.S, outr);
        __syncthreads();
    }
    ph_outw<C, MC::NT_MAC, MC::SPT>(tid, sm.S, outr);
    __syncthreads();
    ph_inv1<C>(tid, NT, sm.wM, sm.S);
    __syncthreads();
    ph_inv2<C>(tid, NT, sm.twist, sm.S, sm.acc);
    __syncthreads();
    uint64_t* dst = node_out + (((size_t)box * n_out + o) * n_pairs + pair) * C::G * N;
    for (int idx = tid; idx < C::G * N; idx += NT) dst[idx] = sm.acc[idx];
}

// ================================================================================================ AES linear layers
// flat state: [block][byte = 4·col + row][bit, MSB first][L]   (reference data_model.rs:165-188 re-expressed)
// AddRoundKey (data_model.rs:270-274): out = in + rk, rk broadcast over blocks
__global__ void aes_add_round_key_kernel(const uint64_t* __restrict__ in, const uint64_t* __restrict__ rk, size_t blk_words, size_t total,
                                         uint64_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i] + rk[i % blk_words];
}
// ShiftRows on the three SBOX·{1,2,3} states + MixColumns + AddRoundKey (fhe_sbox_gal_mul_pbs.rs:61-82, :106-117)
//   new[r][c] = mul2[r] ^ mul1[r-1] ^ mul1[r-2] ^ mul3[r-3]  (rows mod 4, within shifted column c)
// muls: [block][byte][24 = (S, 2S, 3S) × 8 bits][L]
__global__ void aes_mix_columns_kernel(const uint64_t* __restrict__ muls, const uint64_t* __restrict__ rk, int L, size_t total,
                                       uint64_t* __restrict__ state) {
    const size_t byte_words = (size_t)8 * L, blk_words = 16 * byte_words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t blk = i / blk_words; const size_t rem = i - blk * blk_words;
        const int byte = (int)(rem / byte_words); const size_t w = rem - (size_t)byte * byte_words;   // bit·L + e
        const int c = byte >> 2, r = byte & 3;
        const uint64_t* mb = muls + blk * 16 * 24 * (size_t)L;
        auto src = [&](int which, int row) -> uint64_t {
            const int old_byte = 4 * ((c + row) & 3) + row;        // ShiftRows: new[row][c] = old[row][(c+row)%4]
            return mb[((size_t)old_byte * 24 + (size_t)which * 8) * L + w];
        };
        state[i] = src(1, r) + src(0, (r + 3) & 3) + src(0, (r + 2) & 3) + src(2, (r + 1) & 3) + rk[rem];
    }
}
// last round (fhe_sbox_gal_mul_pbs.rs:119-129): ShiftRows(sub) + rk[40..44]
__global__ void aes_final_round_kernel(const uint64_t* __restrict__ sub, const uint64_t* __restrict__ rk, int L, size_t total,
                                       uint64_t* __restrict__ out) {
    const size_t byte_words = (size_t)8 * L, blk_words = 16 * byte_words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t blk = i / blk_words; const size_t rem = i - blk * blk_words;
        const int byte = (int)(rem / byte_words); const size_t w = rem - (size_t)byte * byte_words;
        const int c = byte >> 2, r = byte & 3;
        const int old_byte = 4 * ((c + r) & 3) + r;
        out[i] = sub[blk * blk_words + (size_t)old_byte * byte_words + w] + rk[rem];
    }
}
// leveled XOR (BitXorAssign, shortint_woppbs_1bit.rs:134-142 → lwe_ciphertext_add_assign): a += b, element-wise
__global__ void lwe_add_kernel(uint64_t* __restrict__ a, const uint64_t* __restrict__ b, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) a[i] += b[i];
}
__global__ void negate_kernel(uint64_t* __restrict__ a, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) a[i] = 0ull - a[i];
}

}  // namespace tac
