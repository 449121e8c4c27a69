// ep_step.cuh — one CMux/external-product step for a CTA that owns B GLWE accumulators, as barrier-separated phases.
//
// Shared-memory objects (per CTA):
//   acc   uint64  [B][G][N]       the B accumulators (G = k+1 polynomials each)
//   S     cplx    [B*G][M]        one FFT buffer per (ciphertext, polynomial); reused for the MAC output
//   dig   uint32  [B*G][L][M]     all decomposition digits of this step, samples (jj, jj+M) packed as biased u16 pairs
//   wT    cplx    [M]             combined twist/twiddle table (ep_core.cuh)
// Per-thread registers that live across the phases of one step: out[SPT][B][G] (Fourier-domain accumulators of the
// frequency slots this thread owns).
//
// Phase order for one step (a barrier after every phase):
//   decomp
//   for level = L .. 1:   fwd1(level)  fwd2  mac(level)         (outw follows the last mac without a barrier)
//   inv1  inv2
// Both the CUDA kernels (kernels.cuh) and the CPU emulation (tests/cpu/ep_emul.cpp) follow this order.
#pragma once
#include "ep_core.cuh"

#include <cmath>

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
    static constexpr size_t dig_words = (size_t)B_ * (K_ + 1) * L_ * (N_ / 2);
};
constexpr int floor_pow2(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }
template <class C, int NT>
struct MacCfg {
    static constexpr int NT_MAC = floor_pow2(NT) < C::M ? floor_pow2(NT) : C::M;
    static constexpr int SPT = C::M / NT_MAC;
};

// Decompose every coefficient of the B·G operand polynomials into its L digits.  coef(job, j) returns coefficient j of
// operand polynomial `job` (for a CMux with rotation: (acc·X^rot − acc)[j]).
template <class C, class CoefFn>
TAC_HD void ph_decomp(int tid, int nt, CoefFn coef, int base_log, uint32_t* __restrict__ dig) {
    for (int idx = tid; idx < C::JOBS * C::M; idx += nt) {
        const int job = idx / C::M, jj = idx - job * C::M;
        uint32_t w[C::L];
        decompose_pair<C::L>(coef(job, jj), coef(job, jj + C::M), base_log, w);
#pragma unroll
        for (int s = 0; s < C::L; s++) dig[((size_t)job * C::L + s) * C::M + jj] = w[s];
    }
}
// forward FFT pass 1 of the level-`lev` digits, one job per (ciphertext b, polynomial p)
template <class C>
TAC_HD void ph_fwd1(int tid, int nt, int lev, const uint32_t* __restrict__ dig, const cplx* __restrict__ wT, cplx* __restrict__ S) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) {
        const uint32_t* d = dig + ((size_t)job * C::L + (lev - 1)) * C::M;
        fft_fwd_pass1<C::N>(t, [&](int jj, double& a, double& b) { unpack_digits(d[jj], a, b); }, wT, S + (size_t)job * C::M);
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
        // software pipeline over the G rows: the key loads of row p+1 are in flight while row p is multiplied
        cplx g[2][C::G];
#pragma unroll
        for (int c = 0; c < C::G; c++) g[0][c] = TAC_LDG(gl + (size_t)c * C::M + tau);
#pragma unroll
        for (int p = 0; p < C::G; p++) {
            if (p + 1 < C::G) {
#pragma unroll
                for (int c = 0; c < C::G; c++) g[(p + 1) & 1][c] = TAC_LDG(gl + (size_t)((p + 1) * C::G + c) * C::M + tau);
            }
#pragma unroll
            for (int b = 0; b < C::B; b++) {
                const cplx x = S[(size_t)(b * C::G + p) * C::M + tau];
#pragma unroll
                for (int c = 0; c < C::G; c++) cfma(out[it][b][c], x, g[p & 1][c]);
            }
        }
    }
}
// Written right after the last ph_mac without a barrier: a thread only reads and writes its own slots of S.
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
TAC_HD void ph_inv1(int tid, int nt, const cplx* __restrict__ wT, cplx* __restrict__ S) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) fft_inv_passA<C::N>(t, wT, S + (size_t)job * C::M);
}
template <class C>
TAC_HD void ph_inv2(int tid, int nt, const cplx* __restrict__ S, uint64_t* __restrict__ acc) {
    const int grp = tid >> 4, t = tid & 15, ngrp = nt >> 4;
    for (int job = grp; job < C::JOBS; job += ngrp) {
        uint64_t* poly = acc + (size_t)job * C::N;
        fft_inv_passB<C::N>(t, S + (size_t)job * C::M, [&](int jj, double re, double im) {
            poly[jj] += f64_to_torus(re);
            poly[jj + C::M] += f64_to_torus(im);
        });
    }
}

// Fourier transform of a torus polynomial (keys): 16 threads, buffer S[M]; result left in S in slot order, scaled by `scale`·2^-64.
template <int N>
TAC_HD void key_fft_pass1(int t, const uint64_t* __restrict__ poly, double scale, const cplx* __restrict__ wT, cplx* __restrict__ S) {
    fft_fwd_pass1<N>(t, [&](int jj, double& a, double& b) {
        a = torus_to_f64(poly[jj]) * scale;
        b = torus_to_f64(poly[jj + N / 2]) * scale;
    }, wT, S);
}

// host-side construction of the combined table in extended precision (capi.cu and the CPU emulation)
inline void build_wT(int N, cplx* wT) {
    const int M = N / 2, P = M / 16;
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int q = 0; q < P; q++)
        for (int t = 0; t < 16; t++) {
            const long double ang = pi * t / N - 2.0L * pi * (long double)(t * q) / M;
            cplx w; w.x = (double)cosl(ang); w.y = (double)sinl(ang);
            wT[slot_of(q, t)] = w;
        }
}

}  // namespace tac
