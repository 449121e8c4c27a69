// kernels_wide512.cuh — EXPERIMENT (not part of the product): the level-parallel small-batch PBS kernel on 512 threads.
//
// MEASURED SLOWER (B200, tools/pbs_bench.cu, gpurun_out/r2_exp13.log): one wave of 128 ciphertexts 6.0 ms against 3.9 ms for
// pbs_wide_kernel (256 threads, 254 registers); 6144 ciphertexts 250 ms against 164 ms.  At 128 registers per thread the
// FFT passes lose the room to batch their shared-memory loads, and the latency of one pass — what a lone ciphertext per SM
// is bound by — grows by more than the wider digit and MAC phases save.
#pragma once
#include "kernels_ep.cuh"

namespace tac {

// ================================================================================================ PBS, level-parallel, 512 threads
// The same four phases as pbs_wide_kernel for ONE ciphertext per CTA, with every phase that can use them spread over 16
// warps (128 registers): the digits of all 1280 coefficient pairs over all 512 threads, and the Fourier MAC over lane
// PAIRS — lane h = 0 of a pair accumulates the real parts of a frequency slot, h = 1 the imaginary parts; both lanes
// read the same key row and transform values (one coalesced access), and the two FMAs per product run in cfma's order,
// so the words are those of the other kernels.  The FFT phases keep one 16-thread group per (level, polynomial).
template <int N, int K, int L, int MAC_DEPTH = 3>
__global__ void __launch_bounds__(512, 1)
pbs_wide512_kernel(const uint64_t* __restrict__ lwe_small, int nct, int n, const cplx* __restrict__ bsk, int base_log, uint64_t alpha,
                   const cplx* __restrict__ g_wT, uint64_t* __restrict__ out_big) {
    typedef EpCfg<N, K, L, 1> C;
    constexpr int NT = 512, JOBS = C::JOBS, ROWS = L * C::G;
    static_assert(C::M == 256 && NT / 16 >= L * JOBS && MAC_DEPTH <= ROWS, "shape");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);                    // [G][N]
    cplx* S = reinterpret_cast<cplx*>(acc + C::acc_words);                     // [L][JOBS][M]
    uint32_t* dig = reinterpret_cast<uint32_t*>(S + (size_t)L * C::s_cplx);    // [JOBS][L][M]
    cplx* wT = reinterpret_cast<cplx*>(dig + (size_t)JOBS * L * C::M);
    int* rot_sm = reinterpret_cast<int*>(wT + tab_len(C::N));                  // [2]
    const int tid = threadIdx.x;
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
    if (tid == 0) { rot_sm[0] = switched(0); rot_sm[1] = switched(n); }
    __syncthreads();
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int p = idx / N, j = idx - p * N;
        uint64_t v = 0;
        if (p == K) {
            const int s = (j + rot_sm[1]) & (2 * N - 1);
            v = (s < N) ? (0ull - alpha) : alpha;
        }
        acc[idx] = v;
    }
    __syncthreads();
    const DecompFast dc = make_decomp_fast(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    auto row_ptr = [&](const cplx* ggsw, int r) { return ggsw + (size_t)((L - 1 - r / C::G) * C::G + (r % C::G)) * C::G * C::M; };
    const int tau = tid >> 1;
    const bool im = tid & 1;
    for (int i = 0; i < n; i++) {
        const int rot = rot_sm[i & 1];
        const cplx* ggsw = bsk + ggsw_sz * i;
        if (tid == 0 && i + 1 < n) rot_sm[(i + 1) & 1] = switched(i + 1);                      // consumed after >= 1 barrier
        // ---- P0: digits of all G·M coefficient pairs, one pair per thread and round
#pragma unroll
        for (int idx = tid; idx < JOBS * C::M; idx += NT) {
            const int pj = idx / C::M, jj = idx - pj * C::M;
            uint32_t w[L];
            uint64_t x0, x1;
            rot_diff_pair<N>(acc + (size_t)pj * N, jj, rot, x0, x1);
            decompose_pair<L>(x0, x1, dc, w);
#pragma unroll
            for (int s = 0; s < L; s++) dig[((size_t)pj * L + s) * C::M + jj] = w[s];
        }
        __syncthreads();
        // ---- P1: forward FFT of level part+1 of polynomial job
        {
            const uint32_t* d = dig + ((size_t)job * L + part) * C::M;
            cplx* Sj = S + ((size_t)part * JOBS + job) * C::M;
            if (active) fft_fwd_pass1<N>(t, [&](int jj, double& a, double& b) { unpack_digits(d[jj], dc, a, b); }, Sj);
            __syncwarp();
            if (active) fft_fwd_pass2<N>(t, wT, Sj);
        }
        cplx g[MAC_DEPTH][C::G];
#pragma unroll
        for (int r = 0; r < MAC_DEPTH; r++) mac_load_row<C, C::M>(row_ptr(ggsw, r), 0, tau, g[r]);
        __syncthreads();
        // ---- P2: Fourier MAC over all L·G key rows, real / imaginary split over lane pairs
        {
            double out[C::G];
#pragma unroll
            for (int c = 0; c < C::G; c++) out[c] = 0.0;
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const int s = L - 1 - r / C::G, p = r % C::G;
                const cplx x = S[((size_t)s * JOBS + p) * C::M + tau];
                const double xa = im ? x.y : x.x, xb = im ? x.x : -x.y;
#pragma unroll
                for (int c = 0; c < C::G; c++) {
                    out[c] = fma(xa, g[r % MAC_DEPTH][c].x, out[c]);
                    out[c] = fma(xb, g[r % MAC_DEPTH][c].y, out[c]);
                }
                if (r + MAC_DEPTH < ROWS) mac_load_row<C, C::M>(row_ptr(ggsw, r + MAC_DEPTH), 0, tau, g[r % MAC_DEPTH]);
            }
            __syncthreads();                     // all slots of all buffers have been read: buffer 0 may take the result
            double* Sd = reinterpret_cast<double*>(S);
#pragma unroll
            for (int c = 0; c < C::G; c++) Sd[2 * ((size_t)c * C::M + tau) + (tid & 1)] = out[c];
        }
        __syncthreads();
        // ---- P3: inverse FFT and accumulate (the groups of part 0)
        if (active && part == 0) grp_inv1<C>(t, job, wT, S);
        __syncwarp();
        if (active && part == 0) grp_inv2<C>(t, job, S, acc);
        __syncthreads();
    }
    constexpr int LW = K * N + 1;
    for (int e = tid; e < LW; e += NT) {
        if (ct >= nct) continue;
        uint64_t v = sample_extract_elem<C>(acc, e);
        if (e == K * N) v += alpha;
        out_big[(size_t)ct * LW + e] = v;
    }
}

}  // namespace tac
