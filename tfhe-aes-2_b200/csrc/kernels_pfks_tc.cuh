// kernels_pfks_tc.cuh — the two keyswitches (LWE keyswitch, private functional packing keyswitch) on the tensor cores.
//
// out[ct][j][col] = corr[j][col] − Σ_k d'[ct][k]·key[j][k][col]  (mod 2^64) is an exact integer GEMM with M = ciphertexts,
// K = (kN+1)·l digits, N = (k+1)·(k+1)N columns.  It is mapped to unsigned 8-bit tensor-core MMAs by splitting both operands
// into byte limbs:   d' = a0 + 2^8·a1  (d' ∈ [0, 2^16); the single value 2^16 is patched afterwards, see below),
//                    key = Σ_{b<8} 2^(8b)·key_b.
// Only limb pairs of weight 8(a+b) < 64 matter mod 2^64: (a0,b) for b = 0..7 and (a1,b) for b = 0..6 — 15 MMAs per tile and
// k-block, accumulated by weight into 8 int32 tiles (each sum stays below 2·4128·255² < 2^31) and recombined at the end.
// The LWE keyswitch (digits d' ∈ [0, 2^3]) is the same GEMM with a single digit limb (NLIMB = 1, 8 MMAs).
//
// Layouts (one 32-wide k-block at a time, every fragment read is a conflict-free 32-bit shared-memory load):
//   digit planes  DA[limb 2][kb][khalf 2][ct (padded to 128)][16 B]
//   key planes    KP[j][kb][column tile of 32][byte b 8][khalf 2][n 32][16 B]          (8 KB contiguous per CTA stage)
// mma.sync.m16n8k32.u8.u8.s32 fragments (PTX ISA): A row = lane/4 (+8), k = 4·(lane%4) (+16); B k = 4·(lane%4) (+16), n = lane/4;
// C row = lane/4 (+8), col = 2·(lane%4) (+1).
#pragma once
#include <cuda_runtime.h>
#include "tac_common.h"

namespace tac {

constexpr int TC_KB = 32;          // k per block
constexpr int TC_MT = 128;         // ciphertexts per CTA
constexpr int TC_NT = 32;          // columns per CTA
constexpr int TC_STAGES = 3;

// ---- digits: exact PFKS decomposition (closest_representable + iterator), biased by B/2, split into two byte planes.
// One thread produces 16 consecutive k of one ciphertext.  d' == 2^16 (digit = +B/2, only on exact ties) is stored as 0 and
// recorded in the fix-up list.
// ks_mode = 0: PFKS (closest_representable first; storage index s ↔ level s+1).
// ks_mode = 1: LWE keyswitch (mask elements only, in_stride words per ciphertext; storage index s ↔ level l-s; one limb).
__global__ void pfks_digits_tc_kernel(const uint64_t* __restrict__ in, int nct, int mpad, int in_stride, int b, int l, int Kd, int nkb, int ks_mode,
                                      uint8_t* __restrict__ DA, uint32_t* __restrict__ fix_count, uint2* __restrict__ fix_list, uint32_t fix_cap) {
    const size_t total = (size_t)mpad * nkb * 2;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ct = (int)(idx % mpad);
        const size_t q = idx / mpad;
        const int khalf = (int)(q & 1), kb = (int)(q >> 1);
        uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
        if (ct < nct) {
#pragma unroll
            for (int kq = 0; kq < 16; kq++) {
                const int k = kb * TC_KB + khalf * 16 + kq;
                if (k >= Kd) continue;
                const int i = k / l, lev = ks_mode ? l - (k - i * l) : k - i * l + 1;
                const uint64_t x = in[(size_t)ct * in_stride + i];
                uint64_t st = decomp_init_state(ks_mode ? x : closest_representable(x, b, l), b, l);
                int64_t d = 0;
                for (int q2 = l; q2 >= lev; q2--) d = decomp_next(st, b);
                uint32_t dp = (uint32_t)(d + (int64_t)(1u << (b - 1)));
                if (dp >> 16) {                                   // d' == 2^16: patched by pfks_fixup_kernel
                    const uint32_t slot = atomicAdd(fix_count, 1u);
                    if (slot < fix_cap) fix_list[slot] = make_uint2((uint32_t)ct, (uint32_t)k);
                    dp = 0;
                }
                lo[kq >> 2] |= (dp & 0xFFu) << (8 * (kq & 3));
                hi[kq >> 2] |= (dp >> 8) << (8 * (kq & 3));
            }
        }
        const size_t plane = (size_t)nkb * 2 * mpad * 16;
        const size_t off = (((size_t)kb * 2 + khalf) * mpad + ct) * 16;
        *reinterpret_cast<uint4*>(DA + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (!ks_mode) *reinterpret_cast<uint4*>(DA + plane + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
}

// ---- key planes (once per key upload)
__global__ void pfks_key_planes_kernel(const uint64_t* __restrict__ key, int nkeys, int Kd, int W, int nkb, uint8_t* __restrict__ KP) {
    const int ntiles = (W + TC_NT - 1) / TC_NT;
    const size_t total = (size_t)nkeys * nkb * ntiles * 2 * TC_NT;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(idx % TC_NT);
        size_t q = idx / TC_NT;
        const int khalf = (int)(q & 1); q >>= 1;
        const int tile = (int)(q % ntiles); q /= ntiles;
        const int kb = (int)(q % nkb);
        const int j = (int)(q / nkb);
        uint32_t pl[8][4];
#pragma unroll
        for (int bb = 0; bb < 8; bb++) { pl[bb][0] = pl[bb][1] = pl[bb][2] = pl[bb][3] = 0; }
#pragma unroll
        for (int kq = 0; kq < 16; kq++) {
            const int k = kb * TC_KB + khalf * 16 + kq;
            const uint64_t v = (k < Kd && tile * TC_NT + n < W) ? key[((size_t)j * Kd + k) * W + tile * TC_NT + n] : 0ull;
#pragma unroll
            for (int bb = 0; bb < 8; bb++) pl[bb][kq >> 2] |= (uint32_t)((v >> (8 * bb)) & 0xFFull) << (8 * (kq & 3));
        }
        uint8_t* base = KP + ((((size_t)j * nkb + kb) * ntiles + tile) * 8) * (2 * TC_NT * 16);
#pragma unroll
        for (int bb = 0; bb < 8; bb++)
            *reinterpret_cast<uint4*>(base + (((size_t)bb * 2 + khalf) * TC_NT + n) * 16) = make_uint4(pl[bb][0], pl[bb][1], pl[bb][2], pl[bb][3]);
    }
}

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- the GEMM: CTA = 128 ciphertexts × 32 columns, 8 warps as 4 (M) × 2 (N), warp tile 32 × 16.  NLIMB digit limbs.
// out[ct][j][col] = corr[j][col] − Σ (+ last_col_add[ct·add_stride] on the last column: the LWE body of the keyswitch)
template <int NLIMB>
__global__ void __launch_bounds__(256, 1)
lwe_gemm_tc_kernel(const uint8_t* __restrict__ DA, int nct, int mpad, const uint8_t* __restrict__ KP, int W, int nkeys, int nkb,
                   const uint64_t* __restrict__ corr, const uint64_t* __restrict__ last_col_add, size_t add_stride, uint64_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t As[TC_STAGES][NLIMB * 2 * TC_MT * 4]; // [limb][khalf][row][4 words]
    __shared__ __align__(16) uint32_t Bs[TC_STAGES][8 * 2 * TC_NT * 4];     // [byte][khalf][n][4 words]
    const int ntiles = (W + TC_NT - 1) / TC_NT;
    const int j = blockIdx.x / ntiles, tile = blockIdx.x - j * ntiles;
    const int ct0 = blockIdx.y * TC_MT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
    const size_t a_plane = (size_t)nkb * 2 * mpad * 16;
    const uint8_t* kp_base = KP + (((size_t)j * nkb) * ntiles + tile) * (8 * 2 * TC_NT * 16);
    const size_t kp_stride = (size_t)ntiles * (8 * 2 * TC_NT * 16);

    auto load_stage = [&](int stage, int kb) {
        // A: 4 chunks of 2 KB ([limb][khalf] × 128 rows × 16 B); B: one 8 KB chunk.  256 threads × 4 × 16 B.
#pragma unroll
        for (int r = 0; r < NLIMB; r++) {
            const int e = tid + r * 256;                 // 16-byte unit of the A stage
            const int chunk = e >> 7, row = e & 127;     // chunk = limb*2 + khalf
            const uint8_t* src = DA + (size_t)(chunk >> 1) * a_plane + (((size_t)kb * 2 + (chunk & 1)) * mpad + ct0 + row) * 16;
            cp_async16(&As[stage][e * 4], src);
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int e = tid + r * 256;                 // 0..511: 16-byte unit of the B stage
            cp_async16(&Bs[stage][e * 4], kp_base + (size_t)kb * kp_stride + (size_t)e * 16);
        }
    };

    int acc[2][2][8][4];
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++)
#pragma unroll
            for (int w = 0; w < 8; w++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[mi][ni][w][e] = 0;

#pragma unroll
    for (int s = 0; s < TC_STAGES - 1; s++) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; kb++) {
        cp_async_wait<TC_STAGES - 2>();
        __syncthreads();
        {   // prefetch the stage that was consumed in the previous iteration
            const int nk = kb + TC_STAGES - 1;
            if (nk < nkb) load_stage(nk % TC_STAGES, nk);
            cp_async_commit();
        }
        const uint32_t* as = As[kb % TC_STAGES];
        const uint32_t* bs = Bs[kb % TC_STAGES];
        uint32_t a[NLIMB][2][4];                         // [limb][mi][frag]
#pragma unroll
        for (int limb = 0; limb < NLIMB; limb++)
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int row = wm * 32 + mi * 16 + g;
                a[limb][mi][0] = as[((limb * 2 + 0) * TC_MT + row) * 4 + t];
                a[limb][mi][1] = as[((limb * 2 + 0) * TC_MT + row + 8) * 4 + t];
                a[limb][mi][2] = as[((limb * 2 + 1) * TC_MT + row) * 4 + t];
                a[limb][mi][3] = as[((limb * 2 + 1) * TC_MT + row + 8) * 4 + t];
            }
#pragma unroll
        for (int ni = 0; ni < 2; ni++) {
            const int n = wn * 16 + ni * 8 + g;
#pragma unroll
            for (int bb = 0; bb < 8; bb++) {
                const uint32_t b0 = bs[((bb * 2 + 0) * TC_NT + n) * 4 + t];
                const uint32_t b1 = bs[((bb * 2 + 1) * TC_NT + n) * 4 + t];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    mma_u8(acc[mi][ni][bb], a[0][mi], b0, b1);
                    if (NLIMB > 1 && bb < 7) mma_u8(acc[mi][ni][bb + 1], a[NLIMB - 1][mi], b0, b1);
                }
            }
        }
    }
    cp_async_wait<0>();
    // recombine the weight tiles and write  corr − Σ
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 2; ni++) {
            const int col = tile * TC_NT + wn * 16 + ni * 8 + 2 * t;
            if (col >= W) continue;                      // W is even: col + 1 < W as well
            const uint64_t c0 = corr[(size_t)j * W + col], c1 = corr[(size_t)j * W + col + 1];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ct = ct0 + wm * 32 + mi * 16 + g + 8 * h;
                if (ct >= nct) continue;
                uint64_t s0 = 0, s1 = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    s0 += (uint64_t)(uint32_t)acc[mi][ni][w][2 * h] << (8 * w);
                    s1 += (uint64_t)(uint32_t)acc[mi][ni][w][2 * h + 1] << (8 * w);
                }
                uint64_t v0 = c0 - s0, v1 = c1 - s1;
                if (last_col_add && col + 1 == W - 1) v1 += last_col_add[(size_t)ct * add_stride];
                *reinterpret_cast<ulonglong2*>(out + ((size_t)ct * nkeys + j) * W + col) = make_ulonglong2(v0, v1);
            }
        }
}

// ---- patch for digits d' == 2^16: out[ct][j][:] −= 2^16 · key[j][k][:]
__global__ void pfks_fixup_kernel(const uint32_t* __restrict__ fix_count, const uint2* __restrict__ fix_list, uint32_t fix_cap,
                                  const uint64_t* __restrict__ key, int Kd, int W, int nkeys, uint64_t* __restrict__ out) {
    const uint32_t n = min(*fix_count, fix_cap);
    for (uint32_t e = blockIdx.x; e < n; e += gridDim.x) {
        const uint2 f = fix_list[e];
        for (int idx = threadIdx.x; idx < nkeys * W; idx += blockDim.x) {
            const int j = idx / W, col = idx - j * W;
            const uint64_t v = key[((size_t)j * Kd + f.y) * W + col] << 16;
            atomicAdd(reinterpret_cast<unsigned long long*>(out + ((size_t)f.x * nkeys + j) * W + col), (unsigned long long)(0ull - v));
        }
    }
}

}  // namespace tac
