// kernels_merged.cuh — experiment: blind rotation with the L levels of a step MERGED in one barrier interval for B ciphertexts.
//
// pbs_kernel runs a step as L × (forward FFT → barrier → MAC → barrier) because one FFT buffer per (ciphertext, polynomial)
// is all that fits next to the accumulators.  Here the accumulators leave shared memory: the 16-thread group of job
// (b, p) is the only writer of accumulator polynomial (b, p) and — apart from the ROTATED reads of the decomposition —
// its only reader, and each of its threads reads and writes the same 32 coefficients in every step (forward pass 1 consumes
// samples t + 16m and t + 16m + M, inverse pass B produces exactly those).  So every thread keeps its 32 coefficients in
// REGISTERS, and a copy for the rotated reads lives in the rows of the level-1 FFT buffer, which are dead between the
// inverse transform of one step and the level-1 forward pass of the next.  Shared memory then holds L buffers per job
// (184 KB for L = 3, B = 3) and a step needs two CTA barriers:
//
//   group (b,p):  rotated reads + digits of all levels | pass 1 × L | pass 2 × L      ── barrier ──
//   slot thread:  Σ over all L·G key rows (one prefetch ring)  → sums in buffer 1      ── barrier ──
//   group (b,c):  inverse pass A | pass B + accumulate (registers) + copy for the next rotation
//
// Same arithmetic in the same order as pbs_kernel: bit-identical results.
#pragma once
#include "kernels_ep.cuh"
#include "tmem_ops.cuh"

namespace tac {

// forward pass 1 with a source indexed by the compile-time register row m (sample jj = t + 16m)
template <int N, class Src>
TAC_HD void xfft_fwd_pass1_m(int t, Src src, cplx* __restrict__ S) {
    constexpr int M = N / 2, P = M / 16;
    cplx v[P];
    static_for<0, P>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        double a, b;
        src(mc, a, b);
        v[bitrev<P>(m)] = mk(a, b);
    });
    dft_fwd_twisted<N>(v);
    static_for<0, P>([&](auto qc) { constexpr int q = decltype(qc)::value; S[slot_of(q, t)] = v[q]; });
}
// inverse pass B with a sink indexed by the compile-time register row m
template <int N, class Sink>
TAC_HD void xfft_inv_passB_m(int t, const cplx* __restrict__ S, Sink sink) {
    constexpr int M = N / 2, P = M / 16, CSTEP = 1024 / N;
    cplx v[P];
    static_for<0, P>([&](auto ic) { constexpr int i = decltype(ic)::value; v[i] = S[slot_of(bitrev<P>(i), t)]; });
    dft_inv<P>(v);
    static_for<0, P>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        const cplx z = mul_w128<false, m * CSTEP>(v[m]);
        sink(mc, z.x, z.y);
    });
}

// coefficients jj and jj + N/2 of (p · X^rot − own), the unrotated coefficients supplied by the caller
template <int N>
TAC_HD void rot_diff_pair_own(const uint64_t* __restrict__ p, int jj, int rot, uint64_t o0, uint64_t o1, uint64_t& x0, uint64_t& x1) {
    constexpr int LOGN = LogN<N>::v;
    const uint32_t s0 = (uint32_t)(jj - rot) & (uint32_t)(2 * N - 1);
    const uint32_t i0 = s0 & (uint32_t)(N - 1), i1 = i0 ^ (uint32_t)(N / 2);
    const uint32_t n0 = s0 >> LOGN, n1 = n0 ^ (i0 >> (LOGN - 1));
    const uint64_t v0 = p[i0], v1 = p[i1];
    const uint32_t m0 = 0u - n0, m1 = 0u - n1;
    const uint64_t w0 = ((uint64_t)((uint32_t)(v0 >> 32) ^ m0) << 32) | ((uint32_t)v0 ^ m0);
    const uint64_t w1 = ((uint64_t)((uint32_t)(v1 >> 32) ^ m1) << 32) | ((uint32_t)v1 ^ m1);
    x0 = (w0 + n0) - o0;
    x1 = (w1 + n1) - o1;
}

// pass 2 of the forward transform split into its load / arithmetic / store parts (explicit software pipelining)
template <int N> TAC_HD void fwd2_load(int q, const cplx* __restrict__ S, cplx (&v)[16]) {
    static_for<0, 16>([&](auto tc) { constexpr int tt = decltype(tc)::value; v[bitrev<16>(tt)] = S[slot_of(q, tt)]; });
}
template <int N> TAC_HD void fwd2_math(const cplx (&base)[4], cplx (&v)[16]) {
    dit_stages<16, 2>(v, [&](auto lc, auto kc, cplx& u, cplx& w) {
        constexpr int LEN = decltype(lc)::value, k = decltype(kc)::value;
        constexpr int lg = (LEN == 16) ? 3 : (LEN == 8) ? 2 : (LEN == 4) ? 1 : 0;
        bfly_r(u, w, rot128<-128 * k / LEN>(base[lg]));
    });
}
template <int N> TAC_HD void fwd2_store(int q, cplx* __restrict__ S, const cplx (&v)[16]) {
    static_for<0, 16>([&](auto rc) { constexpr int r = decltype(rc)::value; S[slot_of(q, r)] = v[r]; });
}
template <int N> TAC_HD void fwd2_base(int q, const cplx* __restrict__ wT, cplx (&base)[4]) {
    constexpr int M = N / 2, P = M / 16;
    static_for<0, 4>([&](auto ic) { constexpr int i = decltype(ic)::value; base[i] = wT[M + ((1 << i) - 1) * P + q]; });
}
// pass 1 split: arithmetic into v, then the stores
template <int N, class Src> TAC_HD void fwd1_math(Src src, cplx (&v)[N / 32]) {
    constexpr int P = N / 32;
    static_for<0, P>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        double a, b;
        src(mc, a, b);
        v[bitrev<P>(m)] = mk(a, b);
    });
    dft_fwd_twisted<N>(v);
}
template <int N> TAC_HD void fwd1_store(int t, cplx* __restrict__ S, const cplx (&v)[N / 32]) {
    static_for<0, N / 32>([&](auto qc) { constexpr int q = decltype(qc)::value; S[slot_of(q, t)] = v[q]; });
}

template <class C, int NS = 0> struct MergedXSmem {
    static constexpr size_t bytes = (size_t)C::L * C::s_cplx * 16 + (size_t)tab_len(C::N) * 16 + 64 + (size_t)NS * C::G * C::M * sizeof(cplx);
};

#ifndef TAC_MG_PARK
#define TAC_MG_PARK 0
#endif
#ifndef TAC_MG_RING_EARLY
#define TAC_MG_RING_EARLY 0
#endif
#ifndef TAC_MG_ORDER
#define TAC_MG_ORDER 0
#endif
#ifndef TAC_MG_BLOG
#define TAC_MG_BLOG 0
#endif
#ifndef TAC_MG_I2F
#define TAC_MG_I2F 0
#endif
#ifndef TAC_MG_PF
#define TAC_MG_PF 0
#endif
#ifndef TAC_MG_DIGITS
#define TAC_MG_DIGITS 0
#endif
#ifndef TAC_MG_PF_MODE
#define TAC_MG_PF_MODE 0
#endif
#ifndef TAC_MG_PIPE
#define TAC_MG_PIPE 0
#endif
#ifdef TAC_MG_TIMING
__device__ long long tac_mg_times[16];
#define MG_T(k) do { if (tid == 0 && blockIdx.x == 0) { const long long now_ = clock64(); tac_mg_times[k] += now_ - mg_tprev; mg_tprev = now_; } } while (0)
#else
#define MG_T(k) do { } while (0)
#endif
#ifndef TAC_MERGED_CH
#define TAC_MERGED_CH 4
#endif

template <int N, int K, int L, int B, int NT, int MAC_DEPTH = 4, int NS = 0>
__global__ void __launch_bounds__(NT, 1)
pbs_mergedx_kernel(const uint64_t* __restrict__ lwe_small, int nct, int n, const cplx* __restrict__ bsk, int base_log, uint64_t alpha,
                  const cplx* __restrict__ g_wT, uint64_t* __restrict__ out_big) {
    typedef EpCfg<N, K, L, B> C;
    constexpr int JOBS = C::JOBS, ROWS = L * C::G, NMAC = C::M, M = C::M, P = M / 16;
    static_assert(N == 512, "one DFT-16 per thread and pass");
    static_assert(L >= 2, "the sums use buffer 1, the rotation copy buffer 0");
    static_assert(NT / 16 >= JOBS && NT >= NMAC && MAC_DEPTH <= ROWS - NS, "thread layout");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* S = reinterpret_cast<cplx*>(smem_raw);                               // [L][JOBS][M]   (storage index s ↔ level s+1)
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);                     // [B][G][N] rotation copy = buffer 0
    cplx* wT = S + (size_t)L * C::s_cplx;
    int* rot_sm = reinterpret_cast<int*>(wT + tab_len(N));                     // [2][B]
    cplx* kst = reinterpret_cast<cplx*>(reinterpret_cast<unsigned char*>(rot_sm) + 64);
    const uint32_t kbar = kstage::smem_u32(reinterpret_cast<unsigned char*>(rot_sm) + 32);
    uint32_t kuses = 0;
    constexpr uint32_t ROW_BYTES = (uint32_t)(C::G * C::M * sizeof(cplx));
    const int tid = threadIdx.x;
    if (NS > 0 && tid == 0) {
        kstage::mbar_init(kbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int job = tid >> 4, t = tid & 15;
    const bool active = job < JOBS;
    const int ct0 = blockIdx.x * B;
    const int n1 = n + 1;
    auto switched = [&](int b, int i) -> int {
        const int ct = ct0 + b;
        if (ct >= nct) return 0;
        uint64_t a = __ldg(lwe_small + (size_t)ct * n1 + i);
        if (i == n) a += (1ull << 62);
        return modswitch(a, LogN<N>::v);
    };
    uint32_t* tslot = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(rot_sm) + 48);
    const int warp = tid >> 5;
#if TAC_MG_PARK
    if (warp == 0) tmem::alloc(tslot, 128);
#endif
    for (int i = tid; i < tab_len(N); i += NT) wT[i] = g_wT[i];
    if (tid < B) { rot_sm[tid] = switched(tid, 0); rot_sm[B + tid] = switched(tid, n); }
#if TAC_MG_PARK
    tmem::fence_before();
#endif
    __syncthreads();
#if TAC_MG_PARK
    tmem::fence_after();
    // parking row of this thread: TMEM lane quarter of its warp, 64 columns per warp of the pair (w, w + 4)
    const uint32_t taddr = *tslot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
#endif
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int b = idx / (C::G * N), rem = idx - b * C::G * N, p = rem / N, j = rem - p * N;
        uint64_t v = 0;
        if (p == K) {
            const int s = (j + rot_sm[B + b]) & (2 * N - 1);
            v = (s < N) ? (0ull - alpha) : alpha;
        }
        acc[idx] = v;
    }
    __syncthreads();
    uint64_t* Rj = acc + (size_t)(active ? job : 0) * N;                       // this group's polynomial (rotation copy)
    // coefficients t + 16m (ownd[2m]) and t + 16m + M (ownd[2m+1]) of this thread, as bit patterns in 64-bit registers
    double ownd[2 * P];
    static_for<0, P>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        ownd[2 * m] = __longlong_as_double((long long)Rj[t + 16 * m]);
        ownd[2 * m + 1] = __longlong_as_double((long long)Rj[t + 16 * m + M]);
    });
#if TAC_MG_PARK
    tmem::st_f64<2 * P>(taddr, ownd);
    tmem::wait_st();
#endif
    const DecompFast dc = make_decomp_fast(TAC_MG_BLOG ? TAC_MG_BLOG : base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
#ifdef TAC_DBG_KEY_ONE_ROW
    auto row_ptr = [&](const cplx* ggsw, int) { return ggsw; };
#else
    auto row_ptr = [&](const cplx* ggsw, int r) { return ggsw + (size_t)((L - 1 - r / C::G) * C::G + (r % C::G)) * C::G * C::M; };
#endif
    auto unpack = [&](uint32_t w, double& a, double& b) {
#if TAC_MG_I2F
        const int half = 1 << (dc.b - 1);
        a = __int2double_rn((int)(w & 0xFFFFu) - half);
        b = __int2double_rn((int)(w >> 16) - half);
#else
        unpack_digits(w, dc, a, b);
#endif
    };
    cplx* Sjob = S + (size_t)(active ? job : 0) * M;                           // + s·JOBS·M for buffer s
    constexpr int SUMS = 1;
    for (int i = 0; i < n; i++) {
        const int rot = rot_sm[(i & 1) * B + (active ? job / C::G : 0)];
#ifdef TAC_DBG_KEY_ONE_ROW
        const cplx* ggsw = bsk;
#else
        const cplx* ggsw = bsk + ggsw_sz * i;
#endif
        if (tid < B && i + 1 < n) rot_sm[((i + 1) & 1) * B + tid] = switched(tid, i + 1);      // consumed after >= 1 barrier
#if TAC_MG_PF > 0
        if (tid < NMAC) {
#pragma unroll
            for (int r = 0; r < TAC_MG_PF; r++)
#pragma unroll
                for (int c = 0; c < C::G; c++) {
#if TAC_MG_PF_MODE == 0
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(row_ptr(ggsw, r) + (size_t)c * M + tid));
#else
                    uint32_t sink;
                    asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(sink) : "l"(row_ptr(ggsw, r) + (size_t)c * M + tid));
#endif
                }
        }
#endif
        if (NS > 0 && tid == 0) {
            kstage::mbar_expect_tx(kbar, NS * ROW_BYTES);
#pragma unroll
            for (int r = 0; r < NS; r++) kstage::bulk_g2s(kstage::smem_u32(kst) + r * ROW_BYTES, row_ptr(ggsw, ROWS - NS + r), ROW_BYTES, kbar);
        }
#ifdef TAC_MG_TIMING
        long long mg_tprev = clock64();
#endif
        // ---- digits of all levels: dg[s][m] packs the level-(s+1) digits of samples t+16m (low) and t+16m+M (high)
        uint32_t dg[L][P];
#if TAC_MG_DIGITS == 0
        if (active) {
            constexpr int CH = TAC_MERGED_CH;
            constexpr int LOGN = LogN<N>::v;
            static_for<0, P, CH>([&](auto cc) {
                constexpr int c0 = decltype(cc)::value;
                uint64_t v0[CH], v1[CH];
                uint32_t n0[CH], n1_[CH];
                static_for<0, CH>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const int jj = t + 16 * (c0 + k);
                    const uint32_t s0 = (uint32_t)(jj - rot) & (uint32_t)(2 * N - 1);
                    const uint32_t i0 = s0 & (uint32_t)(N - 1), i1 = i0 ^ (uint32_t)(N / 2);
                    n0[k] = s0 >> LOGN; n1_[k] = n0[k] ^ (i0 >> (LOGN - 1));
                    v0[k] = Rj[i0]; v1[k] = Rj[i1];
                });
                static_for<0, CH>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, m = c0 + k;
                    const uint32_t m0 = 0u - n0[k], m1 = 0u - n1_[k];
                    const uint64_t w0 = ((uint64_t)((uint32_t)(v0[k] >> 32) ^ m0) << 32) | ((uint32_t)v0[k] ^ m0);
                    const uint64_t w1 = ((uint64_t)((uint32_t)(v1[k] >> 32) ^ m1) << 32) | ((uint32_t)v1[k] ^ m1);
                    const uint64_t x0 = (w0 + n0[k]) - (uint64_t)__double_as_longlong(ownd[2 * m]), x1 = (w1 + n1_[k]) - (uint64_t)__double_as_longlong(ownd[2 * m + 1]);
                    uint32_t w[L];
                    decompose_pair<L>(x0, x1, dc, w);
                    static_for<0, L>([&](auto sc) { constexpr int s = decltype(sc)::value; dg[s][m] = w[s]; });
                });
            });
        }
#else
        // one basic block: the closed-form digits of all 16 pairs first, ONE test for ties afterwards (the exact replay of
        // a tied pair re-reads its operands, which are still in place)
        if (active) {
            constexpr int CH = TAC_MERGED_CH;
            uint32_t probe_any = 0;
            auto operands = [&](int m, uint64_t& x0, uint64_t& x1, uint64_t o0, uint64_t o1) {
                rot_diff_pair_own<N>(Rj, t + 16 * m, rot, o0, o1, x0, x1);
            };
            static_for<0, P, CH>([&](auto cc) {
                constexpr int c0 = decltype(cc)::value;
                uint64_t v0[CH], v1[CH];
                uint32_t n0[CH], n1_[CH];
                constexpr int LOGN = LogN<N>::v;
                static_for<0, CH>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const int jj = t + 16 * (c0 + k);
                    const uint32_t s0 = (uint32_t)(jj - rot) & (uint32_t)(2 * N - 1);
                    const uint32_t i0 = s0 & (uint32_t)(N - 1), i1 = i0 ^ (uint32_t)(N / 2);
                    n0[k] = s0 >> LOGN; n1_[k] = n0[k] ^ (i0 >> (LOGN - 1));
                    v0[k] = Rj[i0]; v1[k] = Rj[i1];
                });
                static_for<0, CH>([&](auto kc) {
                    constexpr int k = decltype(kc)::value, m = c0 + k;
                    const uint32_t m0 = 0u - n0[k], m1 = 0u - n1_[k];
                    const uint64_t w0 = ((uint64_t)((uint32_t)(v0[k] >> 32) ^ m0) << 32) | ((uint32_t)v0[k] ^ m0);
                    const uint64_t w1 = ((uint64_t)((uint32_t)(v1[k] >> 32) ^ m1) << 32) | ((uint32_t)v1[k] ^ m1);
                    const uint64_t x0 = (w0 + n0[k]) - (uint64_t)__double_as_longlong(ownd[2 * m]), x1 = (w1 + n1_[k]) - (uint64_t)__double_as_longlong(ownd[2 * m + 1]);
                    const uint64_t y0 = x0 + dc.add, y1 = x1 + dc.add;
                    static_for<0, L>([&](auto sc) {
                        constexpr int s = decltype(sc)::value, l = s + 1;
                        const uint32_t f0 = (uint32_t)(y0 >> (64 - dc.b * l)) & dc.mask;
                        const uint32_t f1 = (uint32_t)(y1 >> (64 - dc.b * l)) & dc.mask;
                        dg[s][m] = f0 | (f1 << 16);
                        probe_any |= dg[s][m] - 0x00010001u;
                    });
                });
            });
            if (probe_any & 0x80008000u) {
                static_for<0, P>([&](auto mc) {
                    constexpr int m = decltype(mc)::value;
                    uint32_t probe = 0;
                    static_for<0, L>([&](auto sc) { probe |= dg[decltype(sc)::value][m] - 0x00010001u; });
                    if (probe & 0x80008000u) {
                        uint64_t x0, x1;
                        operands(m, x0, x1, (uint64_t)__double_as_longlong(ownd[2 * m]), (uint64_t)__double_as_longlong(ownd[2 * m + 1]));
                        const uint64_t p0 = decompose_digits_slow<L>(x0, dc.b), p1 = decompose_digits_slow<L>(x1, dc.b);
                        static_for<0, L>([&](auto sc) {
                            constexpr int s = decltype(sc)::value;
                            dg[s][m] = ((uint32_t)(p0 >> (16 * s)) & 0xFFFFu) | (((uint32_t)(p1 >> (16 * s)) & 0xFFFFu) << 16);
                        });
                    }
                });
            }
        }
#endif
#if TAC_MG_PIPE == 0
        MG_T(0);
        auto p1 = [&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if (active)
                xfft_fwd_pass1_m<N>(t, [&](auto mc, double& a, double& b) { unpack(dg[s][decltype(mc)::value], a, b); },
                                   Sjob + (size_t)s * JOBS * M);
        };
        auto p2 = [&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if (active) fft_fwd_pass2<N>(t, wT, Sjob + (size_t)s * JOBS * M);
        };
        auto batched = [&]() {
            // pass 1 of every level; buffer 0 (the rotation copy) is overwritten last, after the whole group has read it
            static_for<0, L>([&](auto ic) {
                constexpr int s = L - 1 - decltype(ic)::value;
                if (s == 0) __syncwarp();
                p1(std::integral_constant<int, s>{});
            });
            __syncwarp();
            MG_T(1);
            static_for<0, L>([&](auto ic) { p2(std::integral_constant<int, L - 1 - decltype(ic)::value>{}); });
        };
        auto interleaved = [&]() {
            static_for<0, L>([&](auto ic) {
                constexpr int s = L - 1 - decltype(ic)::value;
                if (s == 0 && L == 1) __syncwarp();
                p1(std::integral_constant<int, s>{});
                __syncwarp();
                p2(std::integral_constant<int, s>{});
            });
        };
#if TAC_MG_ORDER == 0
        batched();
#elif TAC_MG_ORDER == 1
        interleaved();
#else
        if (warp < NT / 64) interleaved(); else batched();
#endif
#elif TAC_MG_PIPE == 1
        // pass 1 as above; pass 2 with the loads of the next level issued before the arithmetic of the current one
        static_for<0, L>([&](auto ic) {
            constexpr int s = L - 1 - decltype(ic)::value;
            if (s == 0) __syncwarp();
            if (active)
                xfft_fwd_pass1_m<N>(t, [&](auto mc, double& a, double& b) { unpack(dg[s][decltype(mc)::value], a, b); },
                                   Sjob + (size_t)s * JOBS * M);
        });
        __syncwarp();
#if TAC_MG_RING_EARLY
        cplx g[MAC_DEPTH][C::G];
        if (tid < NMAC) {
#pragma unroll
            for (int r = 0; r < MAC_DEPTH; r++) mac_load_row<C, NMAC>(row_ptr(ggsw, r), 0, tid, g[r]);
        }
#endif
        MG_T(1);
        if (active) {
            static_assert(L == 3, "pipeline written for three levels");
            cplx base[4], va[16], vb[16];
            fwd2_base<N>(t, wT, base);
            fwd2_load<N>(t, Sjob + (size_t)2 * JOBS * M, va);
            fwd2_load<N>(t, Sjob + (size_t)1 * JOBS * M, vb);
            fwd2_math<N>(base, va);
            fwd2_store<N>(t, Sjob + (size_t)2 * JOBS * M, va);
            fwd2_load<N>(t, Sjob + (size_t)0 * JOBS * M, va);
            fwd2_math<N>(base, vb);
            fwd2_store<N>(t, Sjob + (size_t)1 * JOBS * M, vb);
            fwd2_math<N>(base, va);
            fwd2_store<N>(t, Sjob + (size_t)0 * JOBS * M, va);
        }
#elif TAC_MG_PIPE == 2
        // the pass-2 loads of level L are in flight during the pass-1 arithmetic of level 1
        {
            static_assert(L == 3, "pipeline written for three levels");
            cplx base[4], va[16], v1[16];
            if (active) {
                xfft_fwd_pass1_m<N>(t, [&](auto mc, double& a, double& b) { unpack(dg[2][decltype(mc)::value], a, b); }, Sjob + (size_t)2 * JOBS * M);
                xfft_fwd_pass1_m<N>(t, [&](auto mc, double& a, double& b) { unpack(dg[1][decltype(mc)::value], a, b); }, Sjob + (size_t)1 * JOBS * M);
            }
            __syncwarp();
            MG_T(1);
            if (active) {
                fwd2_base<N>(t, wT, base);
                fwd2_load<N>(t, Sjob + (size_t)2 * JOBS * M, va);
                fwd1_math<N>([&](auto mc, double& a, double& b) { unpack(dg[0][decltype(mc)::value], a, b); }, v1);
                fwd1_store<N>(t, Sjob, v1);
                fwd2_math<N>(base, va);
                fwd2_store<N>(t, Sjob + (size_t)2 * JOBS * M, va);
                fwd2_load<N>(t, Sjob + (size_t)1 * JOBS * M, va);
            }
            __syncwarp();
            if (active) {
                fwd2_load<N>(t, Sjob, v1);
                fwd2_math<N>(base, va);
                fwd2_store<N>(t, Sjob + (size_t)1 * JOBS * M, va);
                fwd2_math<N>(base, v1);
                fwd2_store<N>(t, Sjob, v1);
            }
        }
#endif
#if !TAC_MG_RING_EARLY
        cplx g[MAC_DEPTH][C::G];
        if (tid < NMAC) {
#pragma unroll
            for (int r = 0; r < MAC_DEPTH; r++) mac_load_row<C, NMAC>(row_ptr(ggsw, r), 0, tid, g[r]);
        }
#endif
        MG_T(2);
        __syncthreads();
        MG_T(3);
        // ---- Fourier MAC over all L·G key rows
        if (tid < NMAC) {
            cplx out[B][C::G];
#pragma unroll
            for (int b = 0; b < B; b++)
#pragma unroll
                for (int c = 0; c < C::G; c++) out[b][c] = mk(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < ROWS - NS; r++) {
                const int s = L - 1 - r / C::G, p = r % C::G;
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const cplx x = S[((size_t)s * JOBS + b * C::G + p) * M + tid];
#pragma unroll
                    for (int c = 0; c < C::G; c++) cfma(out[b][c], x, g[r % MAC_DEPTH][c]);
                }
                if (r + MAC_DEPTH < ROWS - NS) mac_load_row<C, NMAC>(row_ptr(ggsw, r + MAC_DEPTH), 0, tid, g[r % MAC_DEPTH]);
            }
            if constexpr (NS > 0) {
                kstage::mbar_wait(kbar, kuses & 1u);
#pragma unroll
                for (int r = ROWS - NS; r < ROWS; r++) {
                    const int s = L - 1 - r / C::G, p = r % C::G;
                    cplx row[C::G];
#pragma unroll
                    for (int c = 0; c < C::G; c++) row[c] = kst[(size_t)((r - (ROWS - NS)) * C::G + c) * M + tid];
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        const cplx x = S[((size_t)s * JOBS + b * C::G + p) * M + tid];
#pragma unroll
                        for (int c = 0; c < C::G; c++) cfma(out[b][c], x, row[c]);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < B; b++)
#pragma unroll
                for (int c = 0; c < C::G; c++) S[((size_t)SUMS * JOBS + b * C::G + c) * M + tid] = out[b][c];
        }
        if (NS > 0) kuses++;
        MG_T(4);
        __syncthreads();
        MG_T(5);
#if TAC_MG_PARK == 2
        // the parked coefficients travel back while the inverse transform runs (waited for before the accumulate)
        uint32_t pr[4 * P];
        tmem::wait_st();
        tmem::ld<4 * P>(taddr, pr);
#endif
        // ---- inverse transform of the sums; accumulate in registers; refresh the rotation copy
        if (active) fft_inv_passA<N>(t, wT, Sjob + (size_t)SUMS * JOBS * M);
        __syncwarp();
        MG_T(6);
#if TAC_MG_PARK == 1
        tmem::ld_f64<2 * P>(taddr, ownd);
#elif TAC_MG_PARK == 2
        tmem::wait_ld();
        static_for<0, 2 * P>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            ownd[i] = __hiloint2double((int)pr[2 * i + 1], (int)pr[2 * i]);
        });
#endif
        if (active)
            xfft_inv_passB_m<N>(t, Sjob + (size_t)SUMS * JOBS * M, [&](auto mc, double re, double im) {
                constexpr int m = decltype(mc)::value;
                const uint64_t a0 = (uint64_t)__double_as_longlong(ownd[2 * m]) + f64_to_torus(re);
                const uint64_t a1 = (uint64_t)__double_as_longlong(ownd[2 * m + 1]) + f64_to_torus(im);
                ownd[2 * m] = __longlong_as_double((long long)a0);
                ownd[2 * m + 1] = __longlong_as_double((long long)a1);
                Rj[t + 16 * m] = a0;
                Rj[t + 16 * m + M] = a1;
            });
#if TAC_MG_PARK
        tmem::st_f64<2 * P>(taddr, ownd);
#if TAC_MG_PARK == 1
        tmem::wait_st();
#endif
#endif
        __syncwarp();
        MG_T(7);
    }
#if TAC_MG_PARK
    tmem::wait_st();
    tmem::fence_before();
#endif
    __syncthreads();
#if TAC_MG_PARK
    if (warp == 0) tmem::dealloc(*tslot, 128);
#endif
    constexpr int LW = K * N + 1;
    for (int idx = tid; idx < B * LW; idx += NT) {
        const int b = idx / LW, e = idx - b * LW, ct = ct0 + b;
        if (ct >= nct) continue;
        uint64_t v = sample_extract_elem<C>(acc + (size_t)b * C::G * N, e);
        if (e == K * N) v += alpha;
        out_big[(size_t)ct * LW + e] = v;
    }
}

}  // namespace tac
