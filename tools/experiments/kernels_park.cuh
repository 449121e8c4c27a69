// kernels_park.cuh — pbs_park_kernel: the persistent blind rotation of pbs_kernel with THREE warps per scheduler.
//
// pbs_kernel runs 8 warps at 255 registers: two warps per scheduler cannot overlap the load / store bursts of the FFT
// passes with the FP64 work of another warp (ncu: FP64 pipe 46 %, LSU data pipe 63 %, every phase at about half of its
// own bound).  The register budget is spent on state that is idle most of the time: the Fourier-domain accumulators
// `out` (B·(k+1) complex per MAC thread) are only touched in the MAC phases but live through every FFT pass.  Here they
// are PARKED in tensor memory between the MAC phases (tcgen05.st after a MAC, tcgen05.ld issued before the barrier that
// precedes the next one), which brings every phase under 168 registers: 12 warps, B = 4 ciphertexts per CTA (20 forward
// FFT jobs in flight instead of 15, every BSK row reused four times instead of three).
//
// Same arithmetic in the same order as pbs_kernel (ep_step.cuh phases): bit-identical outputs.
//
// MEASURED SLOWER (B200, 6144 ciphertexts, tools/pbs_bench.cu, gpurun_out/r2_exp9.log): pbs_kernel B=3/256 thr 112.4 ms;
// parked B=3/256 thr 127.3 ms (the parking itself costs 13 %); parked B=3/384 thr (168 registers, spills) 146.3 ms;
// parked + split MAC B=4/512 thr (128 registers, 12 B of spills) 136.1 ms.  Kept as an experiment, not part of the product.
#pragma once
#include "kernels_ep.cuh"
#include "tmem_ops.cuh"

namespace tac {

// ---- real / imaginary split: a LANE PAIR owns one frequency slot, lane h = 0 accumulates the real parts, h = 1 the
// imaginary parts.  Both lanes read the same key row and the same transform values (one coalesced / broadcast access), so
// nothing is fetched twice; each thread carries B·G doubles instead of B·G complex values.  The two FMAs per product run in
// the order cfma uses, so the split MAC produces the same words as the one-thread-per-slot MAC.
template <class C, int MAC_DEPTH>
TAC_HD void ph_mac_prefetch_split(int tid, int lev, const cplx* __restrict__ ggsw, cplx (&g)[MAC_DEPTH][C::G]) {
    const cplx* gl = ggsw + (size_t)(lev - 1) * C::G * C::G * C::M;
#pragma unroll
    for (int p = 0; p < MAC_DEPTH && p < C::G; p++) mac_load_row<C, C::M>(gl, p, tid >> 1, g[p]);
}
template <class C, int MAC_DEPTH>
TAC_HD void ph_mac_split(int tid, int lev, const cplx* __restrict__ ggsw, const cplx* __restrict__ S, double (&out)[C::B][C::G], cplx (&g)[MAC_DEPTH][C::G]) {
    const cplx* gl = ggsw + (size_t)(lev - 1) * C::G * C::G * C::M;
    const int tau = tid >> 1;
    const bool im = tid & 1;
#pragma unroll
    for (int p = 0; p < C::G; p++) {
#pragma unroll
        for (int b = 0; b < C::B; b++) {
            const cplx x = S[(size_t)(b * C::G + p) * C::M + tau];
            const double xa = im ? x.y : x.x, xb = im ? x.x : -x.y;
#pragma unroll
            for (int c = 0; c < C::G; c++) {
                out[b][c] = fma(xa, g[p % MAC_DEPTH][c].x, out[b][c]);
                out[b][c] = fma(xb, g[p % MAC_DEPTH][c].y, out[b][c]);
            }
        }
        if (p + MAC_DEPTH < C::G) mac_load_row<C, C::M>(gl, p + MAC_DEPTH, tau, g[p % MAC_DEPTH]);
    }
}
template <class C>
TAC_HD void ph_outw_split(int tid, cplx* __restrict__ S, double (&out)[C::B][C::G]) {
    const int tau = tid >> 1;
    double* Sd = reinterpret_cast<double*>(S);
#pragma unroll
    for (int b = 0; b < C::B; b++)
#pragma unroll
        for (int c = 0; c < C::G; c++) Sd[2 * ((size_t)(b * C::G + c) * C::M + tau) + (tid & 1)] = out[b][c];
}

// group → job: the first 16 groups (warps 0-7) take one job each, further jobs take the LOWER half of warps 8, 9, …, so that
// every scheduler sees the same number of busy warps
template <int JOBS, int NGROUPS>
__device__ __forceinline__ int park_job_of_group(int grp) {
    if constexpr (JOBS <= 16 || 16 + 2 * (JOBS - 16) > NGROUPS) return grp < JOBS ? grp : -1;
    else {
        if (grp < 16) return grp;
        const int e = grp - 16;
        return ((e & 1) == 0 && 16 + (e >> 1) < JOBS) ? 16 + (e >> 1) : -1;
    }
}

// SPLIT: the MAC runs on 2·M threads with the real / imaginary split of ep_step.cuh (half the accumulator registers per thread)
template <int N, int K, int L, int B, int NT, int MAC_DEPTH = 3, bool SPLIT = false>
__global__ void __launch_bounds__(NT, 1)
pbs_park_kernel(const uint64_t* __restrict__ lwe_small, int nct, int n, const cplx* __restrict__ bsk, int base_log, uint64_t alpha,
                const cplx* __restrict__ g_wT, uint64_t* __restrict__ out_big) {
    typedef EpCfg<N, K, L, B> C;
    constexpr int NMAC = SPLIT ? 2 * C::M : C::M;                // MAC threads: warps 0 .. NMAC/32-1
    constexpr int NREG = B * C::G * (SPLIT ? 2 : 4);             // 32-bit registers of `out`
    constexpr int WPQ = NMAC / 128;                              // MAC warps per TMEM lane quarter
    constexpr uint32_t TCOLS = NREG * WPQ <= 128 ? 128u : NREG * WPQ <= 256 ? 256u : 512u;
    static_assert(C::M == 256 && NT >= NMAC && NT / 16 >= C::JOBS, "one 16-thread group per operand polynomial");
    static_assert(NREG * WPQ <= 512, "parked accumulators exceed tensor memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EpSmem<C> sm(smem_raw);
    int* rot_sm = reinterpret_cast<int*>(sm.extra);               // [2][B]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(rot_sm + 2 * B);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ct0 = blockIdx.x * B;
    const int n1 = n + 1;
    auto switched = [&](int b, int i) -> int {
        const int ct = ct0 + b;
        if (ct >= nct) return 0;
        uint64_t a = __ldg(lwe_small + (size_t)ct * n1 + i);
        if (i == n) a += (1ull << 62);
        return modswitch(a, LogN<N>::v);
    };
    if (warp == 0) tmem::alloc(tslot, TCOLS);
    for (int i = tid; i < tab_len(N); i += NT) sm.wT[i] = g_wT[i];
    if (tid < B) { rot_sm[tid] = switched(tid, 0); rot_sm[B + tid] = switched(tid, n); }
    tmem::fence_before();
    __syncthreads();
    tmem::fence_after();
    // this thread's parking row: TMEM lane quarter of its warp, column block of the warp pair
    const uint32_t taddr = *tslot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * NREG);          // (only MAC warps use it)
    for (int idx = tid; idx < (int)C::acc_words; idx += NT) {
        const int b = idx / (C::G * N), rem = idx - b * C::G * N, p = rem / N, j = rem - p * N;
        uint64_t v = 0;
        if (p == K) {
            const int s = (j + rot_sm[B + b]) & (2 * N - 1);
            v = (s < N) ? (0ull - alpha) : alpha;
        }
        sm.acc[idx] = v;
    }
    __syncthreads();
    const int job = park_job_of_group<C::JOBS, NT / 16>(tid >> 4), t = tid & 15;
    const bool active = job >= 0, mac = tid < NMAC;
    const DecompFast dc = make_decomp_fast(base_log, L);
    const size_t ggsw_sz = (size_t)L * C::G * C::G * C::M;
    for (int i = 0; i < n; i++) {
        const int* rot = rot_sm + (i & 1) * B;
        const cplx* ggsw = bsk + ggsw_sz * i;
        if (tid < B && i + 1 < n) rot_sm[((i + 1) & 1) * B + tid] = switched(tid, i + 1);      // consumed after >= 1 barrier
        cplx g[MAC_DEPTH][C::G];
        typename std::conditional<SPLIT, double[B][C::G], cplx[1][B][C::G]>::type out;
        double* outd = reinterpret_cast<double*>(&out[0][0]);
        // ---- level L: decomposition fused into forward pass 1
        if (active) grp_decomp_fwd1<C>(t, job, [&](int jj, uint64_t& x0, uint64_t& x1) { rot_diff_pair<N>(sm.acc + (size_t)job * N, jj, rot[job / C::G], x0, x1); },
                                       dc, sm.dig, sm.S);
        __syncwarp();
        if (active) grp_fwd2<C>(t, job, sm.wT, sm.S);
        if (mac) {
            if constexpr (SPLIT) ph_mac_prefetch_split<C, MAC_DEPTH>(tid, L, ggsw, g);
            else ph_mac_prefetch<C, NMAC, MAC_DEPTH>(tid, L, ggsw, g);
#pragma unroll
            for (int q = 0; q < NREG / 2; q++) outd[q] = 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int lev = L; lev >= 1; lev--) {
            if (lev < L) {
                if (active) grp_fwd1<C>(t, job, lev, dc, sm.dig, sm.S);
                __syncwarp();
                if (active) grp_fwd2<C>(t, job, sm.wT, sm.S);
                if (mac) {                                                                  // the key rows travel during the barrier
                    if constexpr (SPLIT) ph_mac_prefetch_split<C, MAC_DEPTH>(tid, lev, ggsw, g);
                    else ph_mac_prefetch<C, NMAC, MAC_DEPTH>(tid, lev, ggsw, g);
                }
                __syncthreads();
                // un-park after the barrier: across it only the key ring is live
                if (mac) tmem::ld_f64<NREG / 2>(taddr, outd);
            }
            if (mac) {
                if constexpr (SPLIT) ph_mac_split<C, MAC_DEPTH>(tid, lev, ggsw, sm.S, out, g);
                else ph_mac<C, NMAC, 1, MAC_DEPTH>(tid, lev, ggsw, sm.S, out, g);
                if (lev > 1) {
                    tmem::st_f64<NREG / 2>(taddr, outd);
                    tmem::wait_st();
                } else {
                    if constexpr (SPLIT) ph_outw_split<C>(tid, sm.S, out);
                    else ph_outw<C, NMAC, 1>(tid, sm.S, out);
                }
            }
            __syncthreads();
        }
        if (active) grp_inv1<C>(t, job, sm.wT, sm.S);
        __syncwarp();
        if (active) grp_inv2<C>(t, job, sm.S, sm.acc);
        __syncwarp();
    }
    tmem::fence_before();
    __syncthreads();
    if (warp == 0) tmem::dealloc(*tslot, TCOLS);
    constexpr int LW = K * N + 1;
    for (int idx = tid; idx < B * LW; idx += NT) {
        const int b = idx / LW, e = idx - b * LW, ct = ct0 + b;
        if (ct >= nct) continue;
        uint64_t v = sample_extract_elem<C>(sm.acc + (size_t)b * C::G * N, e);
        if (e == K * N) v += alpha;
        out_big[(size_t)ct * LW + e] = v;
    }
}
template <class C> struct ParkSmem { static constexpr size_t bytes = EpSmem<C>::bytes + 2 * C::B * sizeof(int) + 16; };

}  // namespace tac
