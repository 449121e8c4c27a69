// tests/cpu/ep_emul.cpp — runs the CUDA kernels' phase functions (csrc/ep_core.cuh, ep_step.cuh) thread by thread on the
// CPU, in the same phase order the kernels use, so that FFT passes / swizzles / slot order / rotation / decomposition
// can be checked against the oracle without a GPU.  Built by tests/test_ep_emulation.py with g++.
#include "../../tfhe-aes-2_b200/csrc/ep_step.cuh"

#include <cmath>
#include <vector>

using namespace tac;
constexpr int MAC_DEPTH = 5;      // ring depth the shipped PBS kernels use

template <int N, int K, int L, int B, int NT>
static void emul_step(const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
    typedef EpCfg<N, K, L, B> C;
    typedef MacCfg<C, NT> MC;
    std::vector<cplx> wT(tab_len(N)); build_wT(N, wT.data());
    // Fourier GGSW with the kernels' own key transform
    const int polys = L * C::G * C::G;
    std::vector<cplx> gf((size_t)polys * C::M), kbuf(C::M);
    for (int q = 0; q < polys; q++) {
        for (int t = 0; t < 16; t++) key_fft_pass1<N>(t, ggsw_std + (size_t)q * N, 1.0 / C::M, kbuf.data());
        for (int t = 0; t < 16; t++) fft_fwd_pass2<N>(t, wT.data(), kbuf.data());
        for (int s = 0; s < C::M; s++) gf[(size_t)q * C::M + s] = kbuf[s];
    }
    struct Regs { cplx v[MC::SPT][C::B][C::G]; cplx g[MAC_DEPTH][C::G]; };
    std::vector<Regs> regs(NT);
    for (auto& r : regs) for (int a = 0; a < MC::SPT; a++) for (int b = 0; b < C::B; b++) for (int c = 0; c < C::G; c++) r.v[a][b][c] = mk(0, 0);
    std::vector<cplx> S(C::s_cplx);
    std::vector<uint32_t> dig(C::dig_words + 1);
    const DecompFast dc = make_decomp_fast(base_log, L);
    // phase order of ep_step_device (kernels_ep.cuh); a group-local __syncwarp() is modelled by finishing a pass for all
    // threads before the next pass starts
    auto groups = [&](auto fn) { for (int tid = 0; tid < NT; tid++) { const int job = tid >> 4, t = tid & 15; if (job < C::JOBS) fn(t, job); } };
    groups([&](int t, int job) {
        grp_decomp_fwd1<C>(t, job, [&](int jj, uint64_t& x0, uint64_t& x1) { rot_diff_pair<N>(acc + (size_t)job * N, jj, rot[job / C::G], x0, x1); }, dc, dig.data(),
                           S.data());
    });
    groups([&](int t, int job) { grp_fwd2<C>(t, job, wT.data(), S.data()); });
    for (int tid = 0; tid < NT; tid++) ph_mac_prefetch<C, MC::NT_MAC, MAC_DEPTH>(tid, L, gf.data(), regs[tid].g);
    for (int tid = 0; tid < NT; tid++) ph_mac<C, MC::NT_MAC, MC::SPT, MAC_DEPTH>(tid, L, gf.data(), S.data(), regs[tid].v, regs[tid].g);
    for (int lev = L - 1; lev >= 1; lev--) {
        groups([&](int t, int job) { grp_fwd1<C>(t, job, lev, dc, dig.data(), S.data()); });
        groups([&](int t, int job) { grp_fwd2<C>(t, job, wT.data(), S.data()); });
        for (int tid = 0; tid < NT; tid++) ph_mac_prefetch<C, MC::NT_MAC, MAC_DEPTH>(tid, lev, gf.data(), regs[tid].g);
        for (int tid = 0; tid < NT; tid++) ph_mac<C, MC::NT_MAC, MC::SPT, MAC_DEPTH>(tid, lev, gf.data(), S.data(), regs[tid].v, regs[tid].g);
    }
    for (int tid = 0; tid < NT; tid++) ph_outw<C, MC::NT_MAC, MC::SPT>(tid, S.data(), regs[tid].v);
    groups([&](int t, int job) { grp_inv1<C>(t, job, wT.data(), S.data()); });
    groups([&](int t, int job) { grp_inv2<C>(t, job, S.data(), acc); });
}

// one step in the order of pbs_merged_kernel (kernels_ep.cuh): the threads' accumulator coefficients in "registers", the
// rotation copy aliased onto FFT buffer 0, all levels in one barrier interval.  acc is [B][G][N] as for emul_step.
template <int N, int K, int L, int B, int NT>
static void emul_step_merged(const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
    typedef EpCfg<N, K, L, B> C;
    constexpr int P = C::M / 16, JOBS = C::JOBS, M = C::M, SUMS = 1, DEPTH = 4;
    std::vector<cplx> wT(tab_len(N)); build_wT(N, wT.data());
    const int polys = L * C::G * C::G;
    std::vector<cplx> gf((size_t)polys * M), kbuf(M);
    for (int q = 0; q < polys; q++) {
        for (int t = 0; t < 16; t++) key_fft_pass1<N>(t, ggsw_std + (size_t)q * N, 1.0 / M, kbuf.data());
        for (int t = 0; t < 16; t++) fft_fwd_pass2<N>(t, wT.data(), kbuf.data());
        for (int s = 0; s < M; s++) gf[(size_t)q * M + s] = kbuf[s];
    }
    std::vector<cplx> S((size_t)L * C::s_cplx);
    uint64_t* R = reinterpret_cast<uint64_t*>(S.data());                 // buffer 0 doubles as the rotation copy
    for (size_t i = 0; i < C::acc_words; i++) R[i] = acc[i];
    struct Regs { uint64_t own0[P], own1[P]; uint32_t dg[L][P]; cplx g[DEPTH][C::G]; };
    std::vector<Regs> regs(NT);
    auto groups = [&](auto fn) { for (int tid = 0; tid < NT; tid++) { const int job = tid >> 4, t = tid & 15; if (job < JOBS) fn(tid, t, job); } };
    groups([&](int tid, int t, int job) { for (int m = 0; m < P; m++) { regs[tid].own0[m] = R[(size_t)job * N + t + 16 * m]; regs[tid].own1[m] = R[(size_t)job * N + t + 16 * m + M]; } });
    const DecompFast dc = make_decomp_fast(base_log, L);
    // a group-local __syncwarp() is modelled by finishing a pass for all threads before the next pass starts
    groups([&](int tid, int t, int job) { mg_digits<C>(t, R + (size_t)job * N, rot[job / C::G], regs[tid].own0, regs[tid].own1, dc, regs[tid].dg); });
    static_for<0, L>([&](auto ic) {
        constexpr int s = L - 1 - decltype(ic)::value;
        groups([&](int tid, int t, int job) {
            fft_fwd_pass1_m<N>(t, [&](auto mc, double& a, double& b) { unpack_digits(regs[tid].dg[s][decltype(mc)::value], dc, a, b); },
                               S.data() + ((size_t)s * JOBS + job) * M);
        });
    });
    for (int s = L - 1; s >= 0; s--) groups([&](int, int t, int job) { fft_fwd_pass2<N>(t, wT.data(), S.data() + ((size_t)s * JOBS + job) * M); });
    for (int tid = 0; tid < M; tid++) mg_mac_prefetch<C, DEPTH>(tid, gf.data(), regs[tid].g);
    for (int tid = 0; tid < M; tid++) mg_mac<C, DEPTH, SUMS>(tid, gf.data(), S.data(), regs[tid].g);
    groups([&](int, int t, int job) { fft_inv_passA<N>(t, wT.data(), S.data() + ((size_t)SUMS * JOBS + job) * M); });
    groups([&](int tid, int t, int job) { mg_inv2<C>(t, S.data() + ((size_t)SUMS * JOBS + job) * M, R + (size_t)job * N, regs[tid].own0, regs[tid].own1); });
    for (size_t i = 0; i < C::acc_words; i++) acc[i] = R[i];
}

// forward transform of a real polynomial given as doubles; returns slot-ordered spectrum + the frequency held by each slot
template <int N>
static void emul_fft(const double* in, double* out_re, double* out_im, int* slot_freq) {
    const int M = N / 2, P = M / 16;
    std::vector<cplx> wT(tab_len(N)), S(M); build_wT(N, wT.data());
    for (int t = 0; t < 16; t++) fft_fwd_pass1<N>(t, [&](int jj, double& a, double& b) { a = in[jj]; b = in[jj + M]; }, S.data());
    for (int t = 0; t < 16; t++) fft_fwd_pass2<N>(t, wT.data(), S.data());
    for (int s = 0; s < M; s++) { out_re[s] = S[s].x; out_im[s] = S[s].y; }
    for (int q = 0; q < P; q++) for (int i = 0; i < 16; i++) slot_freq[slot_of(q, i)] = q + P * i;
    // inverse back into `in`-shaped output appended after the spectrum (roundtrip check): out_re[M..M+N)
    for (int t = 0; t < 16; t++) fft_inv_passA<N>(t, wT.data(), S.data());
    for (int t = 0; t < 16; t++) fft_inv_passB<N>(t, S.data(), [&](int jj, double re, double im) { out_re[M + jj] = re / M; out_re[M + jj + M] = im / M; });
}

template <int L> static void digits_t(uint64_t x, int b, double* out) {
    uint32_t w[L];
    const DecompFast dc = make_decomp_fast(b, L);
    decompose_pair<L>(x, ~x, dc, w);
    for (int s = 0; s < L; s++) { double a, c; unpack_digits(w[s], dc, a, c); out[s] = a; }
}

extern "C" {
int emul_cmux_step(int N, int K, int L, int B, int NT, const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
#define CASE(n, k, l, b, nt) if (N == n && K == k && L == l && B == b && NT == nt) { emul_step<n, k, l, b, nt>(ggsw_std, base_log, rot, acc); return 0; }
    CASE(512, 4, 3, 3, 256) CASE(512, 4, 2, 3, 256) CASE(512, 4, 3, 4, 320) CASE(512, 4, 3, 2, 256) CASE(512, 4, 1, 3, 256) CASE(512, 4, 1, 4, 320) CASE(512, 4, 1, 1, 256)
    CASE(1024, 2, 2, 2, 256) CASE(1024, 2, 4, 2, 256) CASE(1024, 2, 1, 2, 256) CASE(1024, 2, 1, 1, 96)
#undef CASE
    return -1;
}
int emul_cmux_step_merged(int N, int K, int L, int B, int NT, const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
#define CASE(n, k, l, b, nt) if (N == n && K == k && L == l && B == b && NT == nt) { emul_step_merged<n, k, l, b, nt>(ggsw_std, base_log, rot, acc); return 0; }
    CASE(512, 4, 3, 3, 256) CASE(512, 4, 2, 3, 256)
#undef CASE
    return -1;
}
int emul_fft_fwd_inv(int N, const double* in, double* out_re, double* out_im, int* slot_freq) {
    if (N == 512) { emul_fft<512>(in, out_re, out_im, slot_freq); return 0; }
    if (N == 1024) { emul_fft<1024>(in, out_re, out_im, slot_freq); return 0; }
    return -1;
}
uint64_t emul_f64_to_torus(double x) { return f64_to_torus(x); }
void emul_digits(uint64_t x, int b, int l, double* out) {
    if (l == 1) digits_t<1>(x, b, out); else if (l == 2) digits_t<2>(x, b, out); else if (l == 3) digits_t<3>(x, b, out); else digits_t<4>(x, b, out);
}
}
