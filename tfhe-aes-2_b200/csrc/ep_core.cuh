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

// cos/sin of 2πk/32, k = 0..8 (other octants by symmetry)
TAC_HD double cos32(int k) {
    switch (k) {
        case 0: return 1.0;
        case 1: return 0.98078528040323044912618223613424;
        case 2: return 0.92387953251128675612818318939679;
        case 3: return 0.83146961230254523707878837761791;
        case 4: return 0.70710678118654752440084436210485;
        case 5: return 0.55557023301960222474283081394853;
        case 6: return 0.38268343236508977172845998403040;
        case 7: return 0.19509032201612826784828486847702;
        default: return 0.0;
    }
}
// multiply by exp(-2πi k/32) (INV = false) or exp(+2πi k/32) (INV = true), k in [0, 16)
template <bool INV>
TAC_HD cplx mul_w32(cplx d, int k) {
    const double c = 0.70710678118654752440084436210485;
    if (k == 0) return d;
    if (k == 8) return INV ? mk(-d.y, d.x) : mk(d.y, -d.x);
    if (k == 4) return INV ? mk((d.x - d.y) * c, (d.x + d.y) * c) : mk((d.x + d.y) * c, (d.y - d.x) * c);
    if (k == 12) return INV ? mk(-(d.x + d.y) * c, (d.x - d.y) * c) : mk((d.y - d.x) * c, -(d.x + d.y) * c);
    double wr, ws;   // cos, sin of 2πk/32
    if (k < 8) { wr = cos32(k); ws = cos32(8 - k); } else { wr = -cos32(16 - k); ws = cos32(k - 8); }
    const double wi = INV ? ws : -ws;
    return mk(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
}

// in-register DFT, P ∈ {16, 32}.  Forward: DIF, natural in → bit-reversed out.  Inverse: DIT, bit-reversed in → natural
// out, unnormalised.  All loop bounds are compile-time so the twiddles fold to immediates.
template <int P>
TAC_HD void dft_fwd(cplx* v) {
#pragma unroll
    for (int len = P; len >= 2; len >>= 1) {
        const int half = len >> 1, tstep = 32 / len;
#pragma unroll
        for (int s = 0; s < P; s += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const cplx u = v[s + j], w = v[s + j + half];
                v[s + j] = mk(u.x + w.x, u.y + w.y);
                v[s + j + half] = mul_w32<false>(mk(u.x - w.x, u.y - w.y), j * tstep);
            }
        }
    }
}
template <int P>
TAC_HD void dft_inv(cplx* v) {
#pragma unroll
    for (int len = 2; len <= P; len <<= 1) {
        const int half = len >> 1, tstep = 32 / len;
#pragma unroll
        for (int s = 0; s < P; s += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const cplx u = v[s + j];
                const cplx w = mul_w32<true>(v[s + j + half], j * tstep);
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

// ------------------------------------------------------------------------------------------------ forward, pass 1
// `src(j)` returns the real sample j (0 <= j < N) as a double.  Thread t (0..15) of the FFT group.
template <int N, class Src>
TAC_HD void fft_fwd_pass1(int t, Src src, const cplx* __restrict__ twist, const cplx* __restrict__ wM, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16;
    cplx v[P];
#pragma unroll
    for (int m = 0; m < P; m++) {
        const int j = t + 16 * m;
        const double a = src(j), b = src(j + M);
        const cplx tw = twist[j];
        v[m] = mk(a * tw.x - b * tw.y, a * tw.y + b * tw.x);
    }
    dft_fwd<P>(v);
#pragma unroll
    for (int i = 0; i < P; i++) {
        const int q = bitrev<P>(i);
        S[slot_of(q, t)] = (q == 0) ? v[i] : cmul(v[i], wM[t * q]);
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
TAC_HD void fft_inv_passA(int t, const cplx* __restrict__ wM, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16;
#pragma unroll
    for (int c2 = 0; c2 < P / 16; c2++) {
        const int q = t + 16 * c2;
        cplx v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = S[slot_of(q, i)];
        dft_inv<16>(v);
#pragma unroll
        for (int tt = 0; tt < 16; tt++) S[slot_of(q, tt)] = (tt == 0) ? v[tt] : cmul_conj(v[tt], wM[tt * q]);
    }
}
// ------------------------------------------------------------------------------------------------ inverse, pass B
// `sink(j, value)` receives real sample j (0 <= j < N) of the inverse transform, already multiplied by `scale`.
template <int N, class Sink>
TAC_HD void fft_inv_passB(int t, const cplx* __restrict__ twist, const cplx* __restrict__ S, double scale, Sink sink) {
    constexpr int M = N / 2, P = M / 16;
    cplx v[P];
#pragma unroll
    for (int i = 0; i < P; i++) v[i] = S[slot_of(bitrev<P>(i), t)];
    dft_inv<P>(v);
#pragma unroll
    for (int m = 0; m < P; m++) {
        const int j = t + 16 * m;
        const cplx z = cmul_conj(v[m], twist[j]);
        sink(j, z.x * scale);
        sink(j + M, z.y * scale);
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
// Signed decomposition digit of level `lev` as a double, bit-identical to tfhe's SignedDecomposer iterator
// (tac_common.h::decomp_init_state / decomp_next).  Fast path: the closed-form balanced decomposition
//     field_l = ((x + add) >> (64 - b·l)) & (B-1),   digit_l = field_l - B/2,
// where `add` carries the rounding bit and B/2 at every level.  It agrees with the iterator unless some level
// l >= lev has an exact tie (field_l == 0, i.e. raw digit == B/2, probability ~2^-b per level); only then the
// iterator is replayed.  Matching the tie rule matters: a top-level tie resolved the other way yields a different
// (equally valid) ciphertext, which would make ciphertext-level comparison with the oracle impossible.
struct DecompF64 {
    uint64_t add;        // rounding bit + B/2 at every level
    uint32_t mask;       // B - 1
    double magic_sub;    // 2^52 + B/2
    int b, l;
};
TAC_HD DecompF64 make_decomp(int b, int l) {
    DecompF64 d;
    const int non_rep = 64 - b * l;
    uint64_t add = 1ull << (non_rep - 1);
    for (int lev = 1; lev <= l; lev++) add += (1ull << (b - 1)) << (64 - b * lev);
    d.add = add; d.mask = (1u << b) - 1u; d.magic_sub = 4503599627370496.0 + (double)(1u << (b - 1)); d.b = b; d.l = l;
    return d;
}
TAC_HD double digit_exact(uint64_t x, int b, int l, int lev) {
    uint64_t st = decomp_init_state(x, b, l);
    int64_t d = 0;
    for (int q = l; q >= lev; q--) d = decomp_next(st, b);
    return (double)d;
}
template <int L>
TAC_HD double digit_f64(uint64_t x, const DecompF64& d, int lev) {
    const uint64_t x2 = x + d.add;
    const uint32_t f = (uint32_t)(x2 >> (64 - d.b * lev)) & d.mask;
    bool tie = (f == 0u);
#pragma unroll
    for (int q = 2; q <= L; q++)
        if (q > lev) tie = tie || (((uint32_t)(x2 >> (64 - d.b * q)) & d.mask) == 0u);
    if (tie) return digit_exact(x, d.b, L, lev);
    return u32_magic(f) - d.magic_sub;
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
