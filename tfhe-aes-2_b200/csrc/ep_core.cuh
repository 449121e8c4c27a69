// ep_core.cuh — the external-product step (CMux with a monomial rotation) as thread-level phases.
//
//   acc(GLWE) += GGSW ⊡ (acc · X^rot − acc)            reference path: [U] tfhe fft64/crypto/{bootstrap,ggsw,wop_pbs}.rs
//
// Every phase is a __host__ __device__ function of (thread id, shared buffers): the CUDA kernels call them between
// __syncthreads(), and tests/cpu/ep_emul.cpp runs the very same functions thread by thread on the CPU, so the index
// arithmetic (FFT passes, swizzles, slot order, rotation, decomposition) is validated without a GPU.
//
// FFT: a negacyclic size-N real transform is a size-M = N/2 complex FFT of (p[j] + i·p[j+M])·e^{iπj/N}.  M = 16·P and
// one FFT is done by 16 threads in two register passes (DFT-P over the stride-16 samples, twiddle, transpose through
// shared memory, DFT-16).  The Fourier "slot" order produced by the forward transform is arbitrary but fixed; the
// Fourier-domain keys are produced by the same forward routine, all Fourier-domain work is pointwise, and the inverse
// transform consumes the same order.
//
//   slot σ = q·16 + (i ^ (q & 15))  holds frequency  f = q + P·bitrev4(i)        (q < P, i < 16)
#pragma once
#include "tac_common.h"
#include <cstring>

namespace tac {

#if defined(__CUDACC__)
typedef double2 cplx;
#else
struct alignas(16) cplx { double x, y; };
#endif

TAC_HD cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
TAC_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
TAC_HD cplx cmul_conj(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
TAC_HD void cfma(cplx& o, cplx a, cplx b) {
    o.x += a.x * b.x; o.x -= a.y * b.y;
    o.y += a.x * b.y; o.y += a.y * b.x;
}

// 2^52 + f for a small unsigned integer f, without an int→double conversion instruction
TAC_HD double u32_magic(uint32_t f) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)f);
#else
    const uint64_t bits = 0x4330000000000000ull | (uint64_t)f;
    double d; memcpy(&d, &bits, 8); return d;
#endif
}

// cos(2πk/128), k = 0..32 (first quadrant; the rest by symmetry)
TAC_HD double cos128(int k) {
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.998795456205172392714771604759100694;
        case 2: return 0.995184726672196886244836953109479922;
        case 3: return 0.989176509964780973451673738016243064;
        case 4: return 0.980785280403230449126182236134239037;
        case 5: return 0.970031253194543992603984207286100251;
        case 6: return 0.956940335732208864935797886980269969;
        case 7: return 0.941544065183020778412509402599502357;
        case 8: return 0.923879532511286756128183189396788287;
        case 9: return 0.903989293123443331586200297230537049;
        case 10: return 0.881921264348355029712756863660388350;
        case 11: return 0.857728610000272069902269984284770137;
        case 12: return 0.831469612302545237078788377617905757;
        case 13: return 0.803207531480644909806676512963141924;
        case 14: return 0.773010453362736960810906609758469801;
        case 15: return 0.740951125354959091175616897495162730;
        case 16: return 0.707106781186547524400844362104849039;
        case 17: return 0.671558954847018400625376850427421803;
        case 18: return 0.634393284163645498215171613225493371;
        case 19: return 0.595699304492433343467036528829969890;
        case 20: return 0.555570233019602224742830813948532874;
        case 21: return 0.514102744193221726593693838968815773;
        case 22: return 0.471396736825997648556387625905254378;
        case 23: return 0.427555093430282094320966856888798534;
        case 24: return 0.382683432365089771728459984030398867;
        case 25: return 0.336889853392220050689253212619147570;
        case 26: return 0.290284677254462367636192375817395275;
        case 27: return 0.242980179903263889948274162077471118;
        case 28: return 0.195090322016128267848284868477022241;
        case 29: return 0.146730474455361751658850129646717820;
        case 30: return 0.098017140329560601994195563888641846;
        case 31: return 0.049067674327418014254954976942682658;
        default: return 0.0;
    }
}
// multiply by exp(-2πi k/128) (INV = false) or exp(+2πi k/128) (INV = true), k in [0, 64).  k is a compile-time
// constant after unrolling, so the branches and the table fold away.
template <bool INV>
TAC_HD cplx mul_w128(cplx d, int k) {
    const double c = 0.70710678118654752440084436210485;
    if (k == 0) return d;
    if (k == 32) return INV ? mk(-d.y, d.x) : mk(d.y, -d.x);
    if (k == 16) return INV ? mk((d.x - d.y) * c, (d.x + d.y) * c) : mk((d.x + d.y) * c, (d.y - d.x) * c);
    if (k == 48) return INV ? mk(-(d.x + d.y) * c, (d.x - d.y) * c) : mk((d.y - d.x) * c, -(d.x + d.y) * c);
    double wr, ws;   // cos, sin of 2πk/128
    if (k < 32) { wr = cos128(k); ws = cos128(32 - k); } else { wr = -cos128(64 - k); ws = cos128(k - 32); }
    const double wi = INV ? ws : -ws;
    return mk(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
}

// in-register DFT, P ∈ {16, 32}.  Forward: DIF, natural in → bit-reversed out.  Inverse: DIT, bit-reversed in → natural
// out, unnormalised.  All loop bounds are compile-time so the twiddles fold to immediates.
template <int P>
TAC_HD void dft_fwd(cplx* v) {
#pragma unroll
    for (int len = P; len >= 2; len >>= 1) {
        const int half = len >> 1, tstep = 128 / len;
#pragma unroll
        for (int s = 0; s < P; s += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const cplx u = v[s + j], w = v[s + j + half];
                v[s + j] = mk(u.x + w.x, u.y + w.y);
                v[s + j + half] = mul_w128<false>(mk(u.x - w.x, u.y - w.y), j * tstep);
            }
        }
    }
}
template <int P>
TAC_HD void dft_inv(cplx* v) {
#pragma unroll
    for (int len = 2; len <= P; len <<= 1) {
        const int half = len >> 1, tstep = 128 / len;
#pragma unroll
        for (int s = 0; s < P; s += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const cplx u = v[s + j];
                const cplx w = mul_w128<true>(v[s + j + half], j * tstep);
                v[s + j] = mk(u.x + w.x, u.y + w.y);
                v[s + j + half] = mk(u.x - w.x, u.y - w.y);
            }
        }
    }
}
template <int P> TAC_HD int bitrev(int i) {
    int r = 0;
#pragma unroll
    for (int b = 1; b < P; b <<= 1) { r = (r << 1) | (i & 1); i >>= 1; }
    return r;
}
TAC_HD int slot_of(int q, int i) { return q * 16 + (i ^ (q & 15)); }

// The twist e^{iπj/N} of sample j = t + 16m factors as  tw_t · c_m  with  c_m = e^{iπ·16m/N} = e^{2πi·m/(N/8)}  — a
// compile-time constant per register — and tw_t, which commutes with the DFT over m and is folded into the inter-pass
// twiddle.  One table serves both directions:
//     wT[slot_of(q, t)] = e^{iπt/N} · e^{-2πi·tq/M}            (M entries, swizzled like S so that both the pass-1
//                                                                 (t across lanes) and pass-A (q across lanes) reads are conflict-free)
// ------------------------------------------------------------------------------------------------ forward, pass 1
// `src(jj, a, b)` yields the real samples jj and jj + M (0 <= jj < M) as doubles.  Thread t (0..15) of the FFT group.
template <int N, class Src>
TAC_HD void fft_fwd_pass1(int t, Src src, const cplx* __restrict__ wT, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16, CSTEP = 1024 / N;     // c_m = exp(+2πi · m·CSTEP / 128)
    cplx v[P];
#pragma unroll
    for (int m = 0; m < P; m++) {
        double a, b;
        src(t + 16 * m, a, b);
        v[m] = mul_w128<true>(mk(a, b), m * CSTEP);
    }
    dft_fwd<P>(v);
#pragma unroll
    for (int i = 0; i < P; i++) {
        const int sl = slot_of(bitrev<P>(i), t);
        S[sl] = cmul(v[i], wT[sl]);
    }
}
// ------------------------------------------------------------------------------------------------ forward, pass 2 (in place)
template <int N>
TAC_HD void fft_fwd_pass2(int t, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16;
#pragma unroll
    for (int c2 = 0; c2 < P / 16; c2++) {
        const int q = t + 16 * c2;
        cplx v[16];
#pragma unroll
        for (int tt = 0; tt < 16; tt++) v[tt] = S[slot_of(q, tt)];
        dft_fwd<16>(v);
#pragma unroll
        for (int i = 0; i < 16; i++) S[slot_of(q, i)] = v[i];
    }
}
// ------------------------------------------------------------------------------------------------ inverse, pass A (in place)
template <int N>
TAC_HD void fft_inv_passA(int t, const cplx* __restrict__ wT, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16;
#pragma unroll
    for (int c2 = 0; c2 < P / 16; c2++) {
        const int q = t + 16 * c2;
        cplx v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = S[slot_of(q, i)];
        dft_inv<16>(v);
#pragma unroll
        for (int tt = 0; tt < 16; tt++) {
            const int sl = slot_of(q, tt);
            S[sl] = (tt == 0) ? v[tt] : cmul_conj(v[tt], wT[sl]);
        }
    }
}
// ------------------------------------------------------------------------------------------------ inverse, pass B
// `sink(jj, re, im)` receives real samples jj and jj + M (0 <= jj < M) of the inverse transform (unnormalised).
template <int N, class Sink>
TAC_HD void fft_inv_passB(int t, const cplx* __restrict__ S, Sink sink) {
    constexpr int M = N / 2, P = M / 16, CSTEP = 1024 / N;
    cplx v[P];
#pragma unroll
    for (int i = 0; i < P; i++) v[i] = S[slot_of(bitrev<P>(i), t)];
    dft_inv<P>(v);
#pragma unroll
    for (int m = 0; m < P; m++) {
        const cplx z = mul_w128<false>(v[m], m * CSTEP);
        sink(t + 16 * m, z.x, z.y);
    }
}

// ------------------------------------------------------------------------------------------------ torus conversion
// fractional part of x (a real number whose integer part is irrelevant) as a torus element.
// [U] tfhe fft64/math/fft/mod.rs::convert_add_backward_torus
TAC_HD uint64_t f64_to_torus(double x) {
    const double C = 6755399441055744.0;     // 1.5 * 2^52: (x + C) - C == rint(x) for |x| < 2^51
    const double r = (x + C) - C;
    const double t = (x - r) * 18446744073709551616.0;
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ll_rn(t);
#else
    if (t >= 9223372036854775808.0) return 1ull << 63;
    return (uint64_t)(int64_t)__builtin_nearbyint(t);
#endif
}
TAC_HD double torus_to_f64(uint64_t v) { return (double)(int64_t)v * (1.0 / 18446744073709551616.0); }

// ------------------------------------------------------------------------------------------------ decomposition for the f64 path
// All L signed digits of a torus value, bit-identical to tfhe's SignedDecomposer iterator (tac_common.h), computed once
// per coefficient and step, branch-free.  Matching the iterator's tie rule matters even on the f64 path: a top-level tie
// resolved the other way yields a different (equally valid) ciphertext, which would make the ciphertext-level comparison
// of one external product with the oracle impossible.
// Digits are cached in shared memory as biased 16-bit values (digit + 2^15; |digit| <= 2^(b-1) <= 2^14), the samples jj
// and jj + M packed in one 32-bit word — exactly the pair one FFT input register needs.
constexpr uint32_t kDigitBias = 32768u;
// All L digits of x in 32-bit arithmetic, branch-free; lev-1 indexed, biased by kDigitBias.
//   y      = x + 2^(63-rep)                      rounding to the top rep = b·L bits
//   f_l    = bits [64-b·l, 64-b·(l-1)) of y      raw digit fields (f_1 is the most significant)
//   level l (from L down to 1):  r = f_l + carry_in;  carry_out = r > B/2  ||  (r == B/2 && tiebit_l);  digit = r - carry_out·B
//   tiebit_l = msb(f_{l-1}) for l > 1;  tiebit_1 = the iterator's "balance" decision
//            = f_1 > B/2 || (f_1 == B/2 && (lower fields != 0 || rounding bit))           (decomp_init_state)
// Equality with decomp_init_state/decomp_next, ties included, is checked in tests/test_ep_emulation.py.
template <int L>
TAC_HD void decompose_digits(uint64_t x, int b, uint32_t (&dig)[L]) {
    const int rep = b * L;
    const uint32_t B = 1u << b, half = B >> 1, mask = B - 1u;
    const uint64_t y = x + (1ull << (63 - rep));
    const uint32_t rounding_bit = (uint32_t)(x >> (63 - rep)) & 1u;
    uint32_t f[L];
#pragma unroll
    for (int l = 1; l <= L; l++) f[l - 1] = (uint32_t)(y >> (64 - b * l)) & mask;
    uint32_t lower = 0;
#pragma unroll
    for (int l = 2; l <= L; l++) lower |= f[l - 1];
    const uint32_t balance = (f[0] > half) | ((f[0] == half) & ((lower != 0u) | rounding_bit));
    uint32_t carry = 0;
#pragma unroll
    for (int l = L; l >= 1; l--) {
        const uint32_t r = f[l - 1] + carry;
        const uint32_t tiebit = (l > 1) ? (f[(l > 1) ? l - 2 : 0] >> (b - 1)) : balance;
        carry = (r > half) | ((r == half) & tiebit);
        dig[l - 1] = r - (carry << b) + kDigitBias;
    }
}
// Fast path: the closed-form balanced decomposition
//     field_l = ((x + add) >> (64 - b·l)) & (B-1),   digit_l = field_l - B/2          (add = rounding bit + B/2 at every level)
// agrees with the iterator unless some level is an exact tie (raw digit == B/2  ⇔  field_l == 0, probability ≈ L·2^-b per
// coefficient); only then the exact routine above is replayed, out of line.
struct DecompFast {
    uint64_t add;
    uint32_t mask, bias_adj;     // bias_adj = kDigitBias - B/2
    int b;
};
TAC_HD DecompFast make_decomp_fast(int b, int l) {
    DecompFast d;
    uint64_t add = 1ull << (63 - b * l);
    for (int lev = 1; lev <= l; lev++) add += (1ull << (b - 1)) << (64 - b * lev);
    d.add = add; d.mask = (1u << b) - 1u; d.bias_adj = kDigitBias - (1u << (b - 1)); d.b = b;
    return d;
}
#if defined(__CUDACC__)
template <int L> __device__ __noinline__ void decompose_digits_slow(uint64_t x, int b, uint32_t (&dig)[L]) { decompose_digits<L>(x, b, dig); }
#else
template <int L> inline void decompose_digits_slow(uint64_t x, int b, uint32_t (&dig)[L]) { decompose_digits<L>(x, b, dig); }
#endif
template <int L>
TAC_HD void decompose_digits_fast(uint64_t x, const DecompFast& dc, uint32_t (&dig)[L]) {
    const uint64_t y = x + dc.add;
    bool tie = false;
#pragma unroll
    for (int l = 1; l <= L; l++) {
        const uint32_t f = (uint32_t)(y >> (64 - dc.b * l)) & dc.mask;
        tie = tie || (f == 0u);
        dig[l - 1] = f + dc.bias_adj;
    }
    if (tie) decompose_digits_slow<L>(x, dc.b, dig);
}
template <int L>
TAC_HD void decompose_pair(uint64_t x0, uint64_t x1, const DecompFast& dc, uint32_t (&out)[L]) {
    uint32_t d0[L], d1[L];
    decompose_digits_fast<L>(x0, dc, d0);
    decompose_digits_fast<L>(x1, dc, d1);
#pragma unroll
    for (int s = 0; s < L; s++) out[s] = (d0[s] & 0xFFFFu) | (d1[s] << 16);
}
TAC_HD void unpack_digits(uint32_t w, double& a, double& b) {
    const double sub = 4503599627370496.0 + (double)kDigitBias;       // 2^52 + bias
    a = u32_magic(w & 0xFFFFu) - sub;
    b = u32_magic(w >> 16) - sub;
}
// coefficient j of (p · X^rot − p), rot in [0, 2N)
template <int N>
TAC_HD uint64_t rot_diff(const uint64_t* __restrict__ p, int j, int rot) {
    const int s = (j - rot) & (2 * N - 1);
    const uint64_t v = p[s & (N - 1)];
    constexpr int LOGN = (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : (N == 2048) ? 11 : -1;
    static_assert(LOGN > 0, "unsupported polynomial size");
    const uint64_t neg = (uint64_t)(s >> LOGN) & 1ull;
    return ((v ^ (0ull - neg)) + neg) - p[j];
}

}  // namespace tac
