// kernels_wide_tma.cuh — experiment: the level-parallel (latency) PBS kernel with the WHOLE key stream carried by bulk
// asynchronous copies into a ring of R shared-memory rows, and no register ring.
//
// pbs_wide_kernel's MAC is bound by the SM's L2 ingest: one ciphertext per CTA means the 307 KB GGSW of a step is used once,
// and it can only be requested when the MAC is about to start (the register ring is 3 rows deep).  Here the rows of ALL
// steps form one continuous sequence k = 15·step + r that flows through R shared-memory slots:
//   full[s]   mbarrier, completes when the bytes of the row in slot s have landed (cp.async.bulk complete_tx)
//   empty[s]  mbarrier, completes when all 256 MAC threads have consumed the row in slot s
// Thread 0 is the producer: after consuming row k it waits for empty[k % R] and requests row k + R into that slot.  The
// copies run ahead across the phase boundaries: while the inverse transform, the digits and the forward transforms of the
// next step execute, R rows of the next GGSW arrive, so the MAC starts with R of its 15 rows in place.
#pragma once
#include "kernels_ep.cuh"

namespace tac {

template <class C, int R> struct WideTmaSmem {
    static constexpr size_t bytes = C::acc_words * 8 + (size_t)C::L * C::s_cplx * 16 + (size_t)C::JOBS * C::L * C::M * 4 + (size_t)tab_len(C::N) * 16 + 256 +
                                    (size_t)R * C::G * C::M * sizeof(cplx);
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <int N, int K, int L, int NT, int R>
__global__ void __launch_bounds__(NT, 1)
pbs_wide_tma_kernel(const uint64_t* __restrict__ lwe_small, int nct, int n, const cplx* __restrict__ bsk, int base_log, uint64_t alpha,
                    const cplx* __restrict__ g_wT, uint64_t* __restrict__ out_big) {
    constexpr int B = 1;
    typedef EpCfg<N, K, L, B> C;
    constexpr int JOBS = C::JOBS, ROWS = L * C::G, NMAC = C::M;
    constexpr int NC = NMAC;                       // compute threads; NT = NC (thread 0 also produces) or NC + 32 (producer warp)
    constexpr bool PWARP = NT > NC;
    static_assert(NC / 16 >= L * JOBS && (NT == NC || NT == NC + 32), "one group per (level, polynomial), one MAC thread per slot");
    auto cta_sync = [&]() { if constexpr (PWARP) asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); else __syncthreads(); };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);                    // [G][N]
    cplx* S = reinterpret_cast<cplx*>(acc + C::acc_words);                     // [L][JOBS][M]
    uint32_t* dig = reinterpret_cast<uint32_t*>(S + (size_t)L * C::s_cplx);    // [JOBS][L][M]
    cplx* wT = reinterpret_cast<cplx*>(dig + (size_t)JOBS * L * C::M);
    unsigned char* ctl = reinterpret_cast<unsigned char*>(wT + tab_len(C::N));      // 256 bytes: rot_sm[2], full[R], empty[R]
    int* rot_sm = reinterpret_cast<int*>(ctl);
    const uint32_t full0 = kstage::smem_u32(ctl + 16), empty0 = kstage::smem_u32(ctl + 16 + 8 * R);
    static_assert(16 + 16 * R <= 256, "control block");
    cplx* KR = reinterpret_cast<cplx*>(ctl + 256);                             // [R][G][M]
    constexpr uint32_t ROW_BYTES = (uint32_t)(C::G * C::M * sizeof(cplx));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < R; s++) { kstage::mbar_init(full0 + 8 * s, 1); kstage::mbar_init(empty0 + 8 * s, NC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int grp = tid >> 4, t = tid & 15;
    const int part = grp / JOBS, job = grp - part * JOBS;
    const bool active = grp < L * JOBS;
    const int ct = blockIdx.x;
    const int n1 = n + 1;
    auto switched = [&](int i) -> int {
        if (ct >= nct) return 0;
        uint64_t a = __ldg(lwe_small + (size_t)ct * n1 + i);
        if (i == n) a += (1ull << 62);
        return modswitch(a, LogN<N>::v);
    };
    for (int i = tid; i < tab_len(C::N); i += NT) wT[i] = g_wT[i];
    if (tid == 0) { rot_sm[0] = switched(0); rot_sm[1] = 0; rot_sm[2] = switched(n); }
    __syncthreads();
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int p = idx / N, j = idx - p * N;
        uint64_t v = 0;
        if (p == K) {
            const int s = (j + rot_sm[2]) & (2 * N - 1);
            v = (s < N) ? (0ull - alpha) : alpha;
        }
        acc[idx] = v;
    }
    const DecompFast dc = make_decomp_fast(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    constexpr int P = C::M / 16;
    const int m0 = part * P / L, m1 = (part + 1) * P / L;
    // global row k = ROWS·step + r, MAC order inside a step: level L first, polynomial p inside
    auto row_src = [&](int k) {
        const int step = k / ROWS, r = k - step * ROWS;
        return bsk + ggsw_sz * step + (size_t)((L - 1 - r / C::G) * C::G + (r % C::G)) * C::G * C::M;
    };
    const int total_rows = n * ROWS;
    auto request = [&](int k) {                    // thread 0: row k into slot k % R
        const int s = k % R;
        kstage::mbar_expect_tx(full0 + 8 * s, ROW_BYTES);
        kstage::bulk_g2s(kstage::smem_u32(KR) + (uint32_t)s * ROW_BYTES, row_src(k), ROW_BYTES, full0 + 8 * s);
    };
    if (tid == 0)
        for (int k = 0; k < R && k < total_rows; k++) request(k);
    __syncthreads();
    if constexpr (PWARP) {
        if (tid >= NC) {                           // producer warp: refill every slot as soon as all MAC threads have released it
            if (tid == NC)
                for (int kk = 0; kk + R < total_rows; kk++) {
                    kstage::mbar_wait(empty0 + 8 * (kk % R), (uint32_t)(kk / R) & 1u);
                    request(kk + R);
                }
            return;
        }
    }
    int k = 0;                                     // next row to consume (all threads count alike)
    for (int i = 0; i < n; i++) {
        const int rot = rot_sm[i & 1];
        if (tid == 0 && i + 1 < n) rot_sm[(i + 1) & 1] = switched(i + 1);          // consumed after >= 1 barrier
        // ---- P0: digits
        if (active) {
            const uint64_t* poly = acc + (size_t)job * N;
            uint32_t* dj = dig + (size_t)job * L * C::M;
#pragma unroll 2
            for (int m = m0; m < m1; m++) {
                const int jj = t + 16 * m;
                uint32_t w[L];
                uint64_t x0, x1;
                rot_diff_pair<N>(poly, jj, rot, x0, x1);
                decompose_pair<L>(x0, x1, dc, w);
#pragma unroll
                for (int s = 0; s < L; s++) dj[(size_t)s * C::M + jj] = w[s];
            }
        }
        cta_sync();
        // ---- P1: forward FFT of level part+1 of polynomial job
        {
            const uint32_t* d = dig + ((size_t)job * L + part) * C::M;
            cplx* Sj = S + ((size_t)part * JOBS + job) * C::M;
            if (active) fft_fwd_pass1<N>(t, [&](int jj, double& a, double& b) { unpack_digits(d[jj], dc, a, b); }, Sj);
            __syncwarp();
            if (active) fft_fwd_pass2<N>(t, wT, Sj);
        }
        cta_sync();
        // ---- P2: Fourier MAC, the key rows from the shared-memory ring
        {
            cplx out[C::G];
#pragma unroll
            for (int c = 0; c < C::G; c++) out[c] = mk(0.0, 0.0);
            const uint32_t pbase = (uint32_t)(k / R);                         // ROWS % R == 0: slot and phase of row r are r % R, pbase + r / R
            static_assert(ROWS % R == 0, "unrolled MAC: the ring length divides the rows of a step");
            static_for<0, ROWS>([&](auto rc) {
                constexpr int r = decltype(rc)::value, s = r % R;
                constexpr int lev_s = L - 1 - r / C::G, p = r % C::G;
                const uint32_t par = (pbase + (uint32_t)(r / R)) & 1u;
                const cplx x = S[((size_t)lev_s * JOBS + p) * C::M + tid];
                kstage::mbar_wait(full0 + 8 * s, par);
                const cplx* row = KR + (size_t)s * C::G * C::M;
                cplx kv[C::G];
#pragma unroll
                for (int c = 0; c < C::G; c++) kv[c] = row[(size_t)c * C::M + tid];
#pragma unroll
                for (int c = 0; c < C::G; c++) cfma(out[c], x, kv[c]);
                mbar_arrive(empty0 + 8 * s);
                if constexpr (!PWARP) {
                    if (tid == 0 && k + r + R < total_rows) {
                        kstage::mbar_wait(empty0 + 8 * s, par);
                        request(k + r + R);
                    }
                }
            });
            k += ROWS;
#pragma unroll
            for (int c = 0; c < C::G; c++) S[(size_t)c * C::M + tid] = out[c];
        }
        cta_sync();
        // ---- P3: inverse FFT and accumulate (the groups of part 0)
        if (active && part == 0) grp_inv1<C>(t, job, wT, S);
        __syncwarp();
        if (active && part == 0) grp_inv2<C>(t, job, S, acc);
        cta_sync();
    }
    constexpr int LW = K * N + 1;
    for (int idx = tid; idx < LW; idx += NC) {
        if (ct >= nct) continue;
        uint64_t v = sample_extract_elem<C>(acc, idx);
        if (idx == K * N) v += alpha;
        out_big[(size_t)ct * LW + idx] = v;
    }
}

}  // namespace tac
