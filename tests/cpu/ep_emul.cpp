// tests/cpu/ep_emul.cpp — runs the CUDA kernels' phase functions (csrc/ep_core.cuh, ep_step.cuh) thread by thread on the
// CPU, in the same phase order the kernels use, so that FFT passes / swizzles / slot order / rotation / decomposition
// can be checked against the oracle without a GPU.  Built by tests/test_ep_emulation.py with g++.
#include "../../tfhe-aes-2_b200/csrc/ep_step.cuh"

#include <cmath>
#include <vector>

using namespace tac;

template <int N>
static void build_tables(std::vector<cplx>& twist, std::vector<cplx>& wM) {
    const int M = N / 2;
    const long double pi = 3.141592653589793238462643383279502884L;
    twist.resize(M); wM.resize(M);
    for (int j = 0; j < M; j++) {
        twist[j] = mk((double)cosl(pi * j / N), (double)sinl(pi * j / N));
        wM[j] = mk((double)cosl(-2.0L * pi * j / M), (double)sinl(-2.0L * pi * j / M));
    }
}

template <int N, int K, int L, int B, int NT>
static void emul_step(const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
    typedef EpCfg<N, K, L, B> C;
    typedef MacCfg<C, NT> MC;
    std::vector<cplx> twist, wM; build_tables<N>(twist, wM);
    // Fourier GGSW with the kernels' own key transform
    const int polys = L * C::G * C::G;
    std::vector<cplx> gf((size_t)polys * C::M), buf(C::M);
    for (int q = 0; q < polys; q++) {
        for (int t = 0; t < 16; t++) key_fft_pass1<N>(t, ggsw_std + (size_t)q * N, 1.0 / C::M, twist.data(), wM.data(), buf.data());
        for (int t = 0; t < 16; t++) fft_fwd_pass2<N>(t, buf.data());
        for (int s = 0; s < C::M; s++) gf[(size_t)q * C::M + s] = buf[s];
    }
    struct Regs { cplx v[MC::SPT][C::B][C::G]; };
    std::vector<Regs> regs(NT);
    for (auto& r : regs) for (int a = 0; a < MC::SPT; a++) for (int b = 0; b < C::B; b++) for (int c = 0; c < C::G; c++) r.v[a][b][c] = mk(0, 0);
    std::vector<cplx> S(C::s_cplx);
    const DecompF64 dc = make_decomp(base_log, L);
    for (int lev = L; lev >= 1; lev--) {
        for (int tid = 0; tid < NT; tid++) ph_fwd1<C>(tid, NT, lev, acc, [&](int b) { return rot[b]; }, dc, twist.data(), wM.data(), S.data());
        for (int tid = 0; tid < NT; tid++) ph_fwd2<C>(tid, NT, S.data());
        for (int tid = 0; tid < NT; tid++) ph_mac<C, MC::NT_MAC, MC::SPT>(tid, lev, gf.data(), S.data(), regs[tid].v);
    }
    for (int tid = 0; tid < NT; tid++) ph_outw<C, MC::NT_MAC, MC::SPT>(tid, S.data(), regs[tid].v);
    for (int tid = 0; tid < NT; tid++) ph_inv1<C>(tid, NT, wM.data(), S.data());
    for (int tid = 0; tid < NT; tid++) ph_inv2<C>(tid, NT, twist.data(), S.data(), acc);
}

// forward transform of a real polynomial given as doubles; returns slot-ordered spectrum + the frequency held by each slot
template <int N>
static void emul_fft(const double* in, double* out_re, double* out_im, int* slot_freq) {
    const int M = N / 2, P = M / 16;
    std::vector<cplx> twist, wM, S(M); build_tables<N>(twist, wM);
    for (int t = 0; t < 16; t++) fft_fwd_pass1<N>(t, [&](int j) { return in[j]; }, twist.data(), wM.data(), S.data());
    for (int t = 0; t < 16; t++) fft_fwd_pass2<N>(t, S.data());
    for (int s = 0; s < M; s++) { out_re[s] = S[s].x; out_im[s] = S[s].y; }
    for (int q = 0; q < P; q++) for (int i = 0; i < 16; i++) slot_freq[slot_of(q, i)] = q + P * bitrev<16>(i);
    // inverse back into `in`-shaped output appended after the spectrum (roundtrip check): out_re[M..M+N)
    for (int t = 0; t < 16; t++) fft_inv_passA<N>(t, wM.data(), S.data());
    for (int t = 0; t < 16; t++) fft_inv_passB<N>(t, twist.data(), S.data(), 1.0 / M, [&](int j, double v) { out_re[M + j] = v; });
}

extern "C" {
int emul_cmux_step(int N, int K, int L, int B, int NT, const uint64_t* ggsw_std, int base_log, const int* rot, uint64_t* acc) {
#define CASE(n, k, l, b, nt) if (N == n && K == k && L == l && B == b && NT == nt) { emul_step<n, k, l, b, nt>(ggsw_std, base_log, rot, acc); return 0; }
    CASE(512, 4, 3, 4, 320) CASE(512, 4, 3, 2, 256) CASE(512, 4, 1, 4, 320) CASE(512, 4, 1, 1, 256)
    CASE(1024, 2, 2, 2, 256) CASE(1024, 2, 4, 2, 256) CASE(1024, 2, 1, 2, 256) CASE(1024, 2, 1, 1, 96)
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
    const DecompF64 dc = make_decomp(b, l);
    for (int lev = 1; lev <= l; lev++) out[lev - 1] = (l == 1) ? digit_f64<1>(x, dc, lev) : (l == 2) ? digit_f64<2>(x, dc, lev) : (l == 3) ? digit_f64<3>(x, dc, lev) : digit_f64<4>(x, dc, lev);
}
}
