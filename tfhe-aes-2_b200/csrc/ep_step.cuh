// ep_step.cuh — one CMux/external-product step for a CTA that owns B GLWE accumulators, as barrier-separated phases.
//
// Shared-memory objects (per CTA):
//   acc   uint64  [B][G][N]     the B accumulators (G = k+1 polynomials each)
//   S     cplx    [B*G][M]      one FFT buffer per (ciphertext, polynomial); reused for the MAC output
//   twist cplx    [M]           e^{iπj/N}
//   wM    cplx    [M]           e^{-2πie/M}
// Per-thread registers that live across the phases of one step: out[SPT][B][G] (Fourier-domain accumulators of the
// frequency slots this thread owns).
//
// Phase order for one step (a barrier after every phase):
//   for level = L .. 1:   fwd1(level)  fwd2  mac(level)
//   outw  inv1  inv2
// Both the CUDA kernels (kernels.cu) and the CPU emulation (tests/cpu/ep_emul.cpp) follow this order.
#pragma once
#include "ep_core.cuh"

namespace tac {

#if defined(__CUDA_ARCH__)
#define TAC_LDG(p) __ldg(p)
#else
#define TAC_LDG(p) (*(p))
#endif

template <int N_, int K_, int L_, int B_>
struct EpCfg {
    static constexpr int N = N_, K = K_, L = L_, B = B_, G = K_ + 1, M = N_ / 2, JOBS = B_ * (K_ + 1);
    static constexpr size_t acc_words = (size_t)B_ * (K_ + 1) * N_;
    static constexpr size_t s_cplx = (size_t)B_ * (K_ + 1) * (N_ / 2);
};
constexpr int floor_pow2(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }
template <class C, int NT>
struct MacCfg {
    static constexpr int NT_MAC = floor_pow2(NT) < C::M ? floor_pow2(NT) : C::M;
    static constexpr int SPT = C::M / NT_MAC;
};

// forward FFT pass 1 of the level-`lev` digits of (acc·X^rot − acc), one job per (ciphertext b, polynomial p)
// rotf(b) returns the monomial degree (in [0, 2N)) applied to ciphertext b in this step
template <class C, class RotFn>
TAC_HD void ph_fwd1(int tid, int nt, int lev, const uint64_t* __restrict__ acc, RotFn rotf, const DecompF64& dc,
                    const cplx* __restrict__ twist, const cplx* __restrict__ wM, cplx* __restrict__ S) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) {
        const uint64_t* poly = acc + (size_t)job * C::N;
        const int r = rotf(job / C::G);
        fft_fwd_pass1<C::N>(
            t, [&](int j) { return digit_f64<C::L>(rot_diff<C::N>(poly, j, r), dc, lev); }, twist, wM, S + (size_t)job * C::M);
    }
}
template <class C>
TAC_HD void ph_fwd2(int tid, int nt, cplx* __restrict__ S) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) fft_fwd_pass2<C::N>(t, S + (size_t)job * C::M);
}
// out[b][c] += Σ_p fft(digits_{lev,p} of ct b) · GGSW[lev-1][p][c]   at the slots owned by this thread.
// ggsw: Fourier GGSW of this step, [L][G][G][M] slot-ordered (already scaled by 2^-64 / M).
template <class C, int NT_MAC, int SPT>
TAC_HD void ph_mac(int tid, int lev, const cplx* __restrict__ ggsw, const cplx* __restrict__ S, cplx (&out)[SPT][C::B][C::G]) {
    if (tid >= NT_MAC) return;
    const cplx* gl = ggsw + (size_t)(lev - 1) * C::G * C::G * C::M;
#pragma unroll
    for (int it = 0; it < SPT; it++) {
        const int tau = tid + it * NT_MAC;
#pragma unroll
        for (int p = 0; p < C::G; p++) {
            cplx g[C::G];
#pragma unroll
            for (int c = 0; c < C::G; c++) g[c] = TAC_LDG(gl + (size_t)(p * C::G + c) * C::M + tau);
#pragma unroll
            for (int b = 0; b < C::B; b++) {
                const cplx x = S[(size_t)(b * C::G + p) * C::M + tau];
#pragma unroll
                for (int c = 0; c < C::G; c++) cfma(out[it][b][c], x, g[c]);
            }
        }
    }
}
template <class C, int NT_MAC, int SPT>
TAC_HD void ph_outw(int tid, cplx* __restrict__ S, cplx (&out)[SPT][C::B][C::G]) {
    if (tid >= NT_MAC) return;
#pragma unroll
    for (int it = 0; it < SPT; it++) {
        const int tau = tid + it * NT_MAC;
#pragma unroll
        for (int b = 0; b < C::B; b++)
#pragma unroll
            for (int c = 0; c < C::G; c++) {
                S[(size_t)(b * C::G + c) * C::M + tau] = out[it][b][c];
                out[it][b][c] = mk(0.0, 0.0);
            }
    }
}
template <class C>
TAC_HD void ph_inv1(int tid, int nt, const cplx* __restrict__ wM, cplx* __restrict__ S) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) fft_inv_passA<C::N>(t, wM, S + (size_t)job * C::M);
}
template <class C>
TAC_HD void ph_inv2(int tid, int nt, const cplx* __restrict__ twist, const cplx* __restrict__ S, uint64_t* __restrict__ acc) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) {
        uint64_t* poly = acc + (size_t)job * C::N;
        fft_inv_passB<C::N>(t, twist, S + (size_t)job * C::M, 1.0, [&](int j, double v) { poly[j] += f64_to_torus(v); });
    }
}

// Fourier transform of a torus polynomial (keys): 16 threads, buffer S[M]; result left in S in slot order, scaled by `scale`.
template <int N>
TAC_HD void key_fft_pass1(int t, const uint64_t* __restrict__ poly, double scale, const cplx* __restrict__ twist,
                          const cplx* __restrict__ wM, cplx* __restrict__ S) {
    fft_fwd_pass1<N>(t, [&](int j) { return torus_to_f64(poly[j]) * scale; }, twist, wM, S);
}

}  // namespace tac
