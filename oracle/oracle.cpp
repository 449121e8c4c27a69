// oracle/oracle.cpp — CPU restatement of the tfhe-aes-2 WoP-PBS hot path.  TEST INFRASTRUCTURE ONLY.
// (see oracle.hpp / README.md; "[U]" = restated from the published algorithm of tfhe 0.11.2, which is not
//  vendored in /root/reference — Cargo.lock:721-724)
#include "oracle.hpp"

#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <algorithm>

namespace orc {

// ============================================================================================ parameters
// reference src/tfhe/shortint_woppbs_1bit/parameters.rs:29-61 (lvl_1), :77-109 (lvl_4), :125-157 (lvl_64), :173-205 (lvl_256)
bool params_preset(int id, Params* o) {
    switch (id) {
        case 1:   *o = Params{671, 2, 1024, 2, 15, 4, 3, 1, 10, 1, 24, 1, 4.7280002450549286e-05, 3.162026630747649e-16, 3.162026630747649e-16}; return true;
        case 4:   *o = Params{679, 2, 1024, 2, 15, 4, 3, 1, 11, 2, 16, 4, 4.7280002450549286e-05, 3.162026630747649e-16, 3.162026630747649e-16}; return true;
        case 64:  *o = Params{677, 4, 512, 3, 12, 4, 3, 1, 13, 2, 16, 64, 4.7280002450549286e-05, 0.00000000000000022148688116005568, 0.00000000000000022148688116005568}; return true;
        case 256: *o = Params{665, 2, 1024, 4, 9, 6, 2, 1, 14, 3, 12, 256, 4.7280002450549286e-05, 3.162026630747649e-16, 3.162026630747649e-16}; return true;
        default: return false;
    }
}

// ============================================================================================ ChaCha20
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define ORC_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

void ChaCha20::init(const uint8_t key[32], uint64_t nonce, uint64_t counter) {
    st[0] = 0x61707865; st[1] = 0x3320646e; st[2] = 0x79622d32; st[3] = 0x6b206574;
    for (int i = 0; i < 8; i++) {
        st[4 + i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
    }
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32);
    st[14] = (uint32_t)nonce;   st[15] = (uint32_t)(nonce >> 32);
    pos = 16;
}
void ChaCha20::refill() {
    uint32_t x[16];
    memcpy(x, st, sizeof(x));
    for (int r = 0; r < 10; r++) {
        ORC_QR(x[0], x[4], x[8], x[12]) ORC_QR(x[1], x[5], x[9], x[13]) ORC_QR(x[2], x[6], x[10], x[14]) ORC_QR(x[3], x[7], x[11], x[15])
        ORC_QR(x[0], x[5], x[10], x[15]) ORC_QR(x[1], x[6], x[11], x[12]) ORC_QR(x[2], x[7], x[8], x[13]) ORC_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) buf[i] = x[i] + st[i];
    if (++st[12] == 0) ++st[13];
    pos = 0;
}
void ChaCha20::bytes(uint8_t* out, size_t n) {
    // byte-granular draw (little-endian words), used for the rand_chacha-compatible test stream
    size_t i = 0;
    while (i < n) {
        uint32_t w = next_u32();
        for (int b = 0; b < 4 && i < n; b++, i++) out[i] = (uint8_t)(w >> (8 * b));
    }
}
void rng_key(uint64_t seed, uint32_t domain, uint8_t key[32]) {
    static const char tag[21] = "tfhe-aes-b200 rng v1";
    for (int i = 0; i < 8; i++) key[i] = (uint8_t)(seed >> (8 * i));
    for (int i = 0; i < 4; i++) key[8 + i] = (uint8_t)(domain >> (8 * i));
    memcpy(key + 12, tag, 20);
}

// Box-Muller pair from two u64 draws (spec shared with the product client library)
static inline void gaussian_pair(uint64_t x, uint64_t y, double* z0, double* z1) {
    const double u1 = (double)((x >> 11) + 1) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(y >> 11) * (1.0 / 9007199254740992.0);
    const double r = std::sqrt(-2.0 * std::log(u1));
    const double th = 6.283185307179586476925286766559 * u2;
    *z0 = r * std::cos(th);
    *z1 = r * std::sin(th);
}
static inline uint64_t noise_to_torus(double z, double sigma_scaled) {
    const double v = z * sigma_scaled;
    return (uint64_t)(int64_t)std::llrint(v);
}
static const double TWO64 = 18446744073709551616.0;

// ============================================================================================ decomposer
// [U] tfhe core_crypto/commons/math/decomposition/decomposer.rs::SignedDecomposer::closest_representable
uint64_t closest_representable(uint64_t x, int b, int l) {
    const int non_rep = 64 - b * l;
    const uint64_t msb = (x >> (non_rep - 1)) & 1ull;
    uint64_t res = x >> non_rep;
    res += msb;
    return res << non_rep;
}
// [U] decomposer.rs::init_decomposer_state — rounding + "balanced" tie handling (tfhe >= 0.10).  The tie rule
// (state == B^l/2 is mapped to -B^l/2 when the rounding bit is 1) cannot be checked against the crate here;
// either variant recomposes to the same torus value, so decrypt-level parity is unaffected.  This choice
// defines bit-exactness for the CUDA integer kernels.
uint64_t decomp_init_state(uint64_t x, int b, int l) {
    const int rep = b * l;
    const int non_rep = 64 - rep;
    uint64_t res = x >> (non_rep - 1);
    const uint64_t rounding_bit = res & 1ull;
    res += 1ull;
    res >>= 1;
    res &= (~0ull) >> (64 - rep);
    const uint64_t half = 1ull << (rep - 1);
    const uint64_t need_balance = (res > half || (res == half && rounding_bit == 1ull)) ? 1ull : 0ull;
    return res - (need_balance << rep);
}

// ============================================================================================ FFT
NegFFT::NegFFT(int N_) : N(N_), M(N_ / 2) {
    tw_re.resize(M); tw_im.resize(M); w_re.resize(M / 2); w_im.resize(M / 2);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int j = 0; j < M; j++) {
        tw_re[j] = (double)cosl(pi * j / N);
        tw_im[j] = (double)sinl(pi * j / N);
    }
    for (int k = 0; k < M / 2; k++) {
        w_re[k] = (double)cosl(-2.0L * pi * k / M);
        w_im[k] = (double)sinl(-2.0L * pi * k / M);
    }
}
// [U] tfhe-fft 0.7.0 forward plan: unnormalised, output order unordered (here: bit-reversed); any consistent
// order works because all Fourier-domain operations are pointwise.
__attribute__((target_clones("avx512f", "avx2", "default")))
void NegFFT::fft(double* re, double* im) const {
    for (int len = M; len >= 2; len >>= 1) {
        const int half = len >> 1, stride = M / len;
        for (int s = 0; s < M; s += len) {
            double* ar = re + s; double* ai = im + s; double* br = ar + half; double* bi = ai + half;
            for (int j = 0; j < half; j++) {
                const double wr = w_re[j * stride], wi = w_im[j * stride];
                const double ur = ar[j], ui = ai[j], vr = br[j], vi = bi[j];
                ar[j] = ur + vr; ai[j] = ui + vi;
                const double dr = ur - vr, di = ui - vi;
                br[j] = dr * wr - di * wi;
                bi[j] = dr * wi + di * wr;
            }
        }
    }
}
__attribute__((target_clones("avx512f", "avx2", "default")))
void NegFFT::ifft(double* re, double* im) const {
    for (int len = 2; len <= M; len <<= 1) {
        const int half = len >> 1, stride = M / len;
        for (int s = 0; s < M; s += len) {
            double* ar = re + s; double* ai = im + s; double* br = ar + half; double* bi = ai + half;
            for (int j = 0; j < half; j++) {
                const double wr = w_re[j * stride], wi = -w_im[j * stride];
                const double vr = br[j] * wr - bi[j] * wi;
                const double vi = br[j] * wi + bi[j] * wr;
                const double ur = ar[j], ui = ai[j];
                ar[j] = ur + vr; ai[j] = ui + vi;
                br[j] = ur - vr; bi[j] = ui - vi;
            }
        }
    }
}
// [U] tfhe core_crypto/fft_impl/fft64/math/fft/mod.rs::{convert_forward_integer, forward_as_integer}
void NegFFT::fwd_int(const int64_t* p, double* re, double* im) const {
    for (int j = 0; j < M; j++) {
        const double a = (double)p[j], b = (double)p[j + M];
        re[j] = a * tw_re[j] - b * tw_im[j];
        im[j] = a * tw_im[j] + b * tw_re[j];
    }
    fft(re, im);
}
// [U] fft/mod.rs::{convert_forward_torus, forward_as_torus}: torus value read as i64, scaled by 2^-64
void NegFFT::fwd_torus(const uint64_t* p, double* re, double* im) const {
    const double sc = 1.0 / TWO64;
    for (int j = 0; j < M; j++) {
        const double a = (double)(int64_t)p[j] * sc, b = (double)(int64_t)p[j + M] * sc;
        re[j] = a * tw_re[j] - b * tw_im[j];
        im[j] = a * tw_im[j] + b * tw_re[j];
    }
    fft(re, im);
}
static inline uint64_t f64_to_torus(double x) {
    const double fr = x - std::nearbyint(x);
    const double t = fr * TWO64;
    if (t >= 9223372036854775808.0) return 1ull << 63;
    return (uint64_t)(int64_t)std::nearbyint(t);
}
// [U] fft/mod.rs::{convert_add_backward_torus, add_backward_as_torus}
void NegFFT::add_bwd_torus(uint64_t* out, double* re, double* im) const {
    ifft(re, im);
    const double norm = 1.0 / (double)M;
    for (int j = 0; j < M; j++) {
        const double a = re[j] * norm, b = im[j] * norm;
        const double xr = a * tw_re[j] + b * tw_im[j];    // multiply by conj(twist)
        const double xi = b * tw_re[j] - a * tw_im[j];
        out[j] += f64_to_torus(xr);
        out[j + M] += f64_to_torus(xi);
    }
}

// ============================================================================================ polynomial helpers
// out[j] = (p * X^d)[j], d in [0, 2N).  [U] polynomial_algorithms.rs::polynomial_wrapping_monic_monomial_mul
static inline void monomial_mul(const uint64_t* p, int N, int d, uint64_t* out) {
    const int twoN = 2 * N;
    for (int j = 0; j < N; j++) {
        int src = j - d; src %= twoN; if (src < 0) src += twoN;
        out[j] = (src < N) ? p[src] : (0ull - p[src - N]);
    }
}
// body[t] += sum over set key bits j of a * X^j  (negacyclic a * S with binary S)
__attribute__((target_clones("avx512f", "avx2", "default")))
static void negacyclic_mul_binary_add(const uint64_t* a, const uint64_t* S, int N, uint64_t* body) {
    for (int j = 0; j < N; j++) {
        if (!(S[j] & 1ull)) continue;
        for (int t = 0; t < j; t++) body[t] -= a[t - j + N];
        for (int t = j; t < N; t++) body[t] += a[t - j];
    }
}

// ============================================================================================ key generation
KeySet::~KeySet() { delete fft; }

static void draw_secret(uint64_t seed, uint32_t domain, size_t count, std::vector<uint64_t>& out) {
    uint8_t key[32]; rng_key(seed, domain, key);
    ChaCha20 c; c.init(key, 0);
    out.resize(count);
    for (size_t i = 0; i < count; i++) out[i] = c.next_u64() & 1ull;
}
// GLWE encryption of plaintext polynomial m under S (k polys), stream (domain, index).
// [U] tfhe core_crypto/algorithms/glwe_encryption.rs::encrypt_glwe_ciphertext
static void glwe_encrypt(const Params& p, const uint64_t* S, const uint64_t* m, double sigma, const uint8_t key[32],
                         uint64_t index, uint64_t* out) {
    ChaCha20 c; c.init(key, index);
    const int N = p.N, k = p.k;
    for (int i = 0; i < k * N; i++) out[i] = c.next_u64();
    uint64_t* body = out + (size_t)k * N;
    const double ss = sigma * TWO64;
    for (int t = 0; t < N; t += 2) {
        const uint64_t x = c.next_u64(), y = c.next_u64();
        double z0, z1; gaussian_pair(x, y, &z0, &z1);
        body[t] = m[t] + noise_to_torus(z0, ss);
        body[t + 1] = m[t + 1] + noise_to_torus(z1, ss);
    }
    for (int i = 0; i < k; i++) negacyclic_mul_binary_add(out + (size_t)i * N, S + (size_t)i * N, N, body);
}
// [U] lwe_encryption.rs::encrypt_lwe_ciphertext
static void lwe_encrypt(const uint64_t* s, int dim, uint64_t m, double sigma, const uint8_t key[32], uint64_t index, uint64_t* out) {
    ChaCha20 c; c.init(key, index);
    uint64_t acc = 0;
    for (int i = 0; i < dim; i++) { out[i] = c.next_u64(); acc += out[i] * s[i]; }
    const uint64_t x = c.next_u64(), y = c.next_u64();
    double z0, z1; gaussian_pair(x, y, &z0, &z1);
    out[dim] = acc + m + noise_to_torus(z0, sigma * TWO64);
}

void KeySet::build_fourier() {
    if (!fft) fft = new NegFFT(p.N);
    const int M = p.N / 2, G = p.k + 1;
    const size_t polys = (size_t)p.n * p.pbs_l * G * G;
    bsk_re.resize(polys * M); bsk_im.resize(polys * M);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < (long)polys; q++) fft->fwd_torus(bsk.data() + (size_t)q * p.N, bsk_re.data() + (size_t)q * M, bsk_im.data() + (size_t)q * M);
}

// reference shortint_woppbs_1bit.rs:245-268 → [U] shortint::gen_keys + WopbsKey::new_wopbs_key_only_for_wopbs
KeySet* keygen(const Params& p, uint64_t seed) {
    KeySet* ks = new KeySet();
    ks->p = p; ks->seed = seed;
    const int N = p.N, k = p.k, G = k + 1, big = p.big();
    draw_secret(seed, D_SK_GLWE, (size_t)big, ks->sk_glwe);
    draw_secret(seed, D_SK_LWE, (size_t)p.n, ks->sk_lwe);
    const uint64_t* S = ks->sk_glwe.data();

    // --- bootstrapping key: GGSW(s_i) under the GLWE key.  [U] ggsw_encryption.rs::encrypt_constant_ggsw_ciphertext
    ks->bsk.assign(ks->bsk_len(), 0);
    {
        uint8_t key[32]; rng_key(seed, D_BSK, key);
        const long total = (long)p.n * p.pbs_l * G;
#pragma omp parallel
        {
            std::vector<uint64_t> m(N);
#pragma omp for schedule(dynamic, 16)
            for (long q = 0; q < total; q++) {
                const int r = (int)(q % G); const int s = (int)((q / G) % p.pbs_l); const int i = (int)(q / ((long)G * p.pbs_l));
                const uint64_t g = 1ull << (64 - p.pbs_b * (s + 1));
                const uint64_t f = ks->sk_lwe[i] * g;
                if (r < k) { for (int t = 0; t < N; t++) m[t] = 0ull - S[(size_t)r * N + t] * f; }
                else       { std::fill(m.begin(), m.end(), 0ull); m[0] = f; }
                glwe_encrypt(p, S, m.data(), p.s_glwe, key, (uint64_t)q, ks->bsk.data() + (size_t)q * G * N);
            }
        }
    }
    // --- keyswitch key big -> small.  [U] lwe_keyswitch_key_generation.rs::generate_lwe_keyswitch_key
    ks->ksk.assign(ks->ksk_len(), 0);
    {
        uint8_t key[32]; rng_key(seed, D_KSK, key);
        const long total = (long)big * p.ks_l;
#pragma omp parallel for schedule(static)
        for (long q = 0; q < total; q++) {
            const int s = (int)(q % p.ks_l); const int i = (int)(q / p.ks_l);
            const int level = p.ks_l - s;
            const uint64_t m = S[i] << (64 - p.ks_b * level);
            lwe_encrypt(ks->sk_lwe.data(), p.n, m, p.s_lwe, key, (uint64_t)q, ks->ksk.data() + (size_t)q * (p.n + 1));
        }
    }
    // --- circuit-bootstrap PFPKSKs.  [U] lwe_private_functional_packing_keyswitch_key_generation.rs +
    //     lwe_wopbs.rs::generate_circuit_bootstrap_lwe_pfpksk_list (f_j = -x with polynomial S_j for j<k; identity with 1 for j=k)
    ks->pfpksk.assign(ks->pfpksk_len(), 0);
    {
        uint8_t key[32]; rng_key(seed, D_PFPKSK, key);
        const long per_key = (long)(big + 1) * p.pfks_l;
        const long total = (long)G * per_key;
#pragma omp parallel
        {
            std::vector<uint64_t> m(N);
#pragma omp for schedule(dynamic, 16)
            for (long q = 0; q < total; q++) {
                const int s = (int)(q % p.pfks_l); const int i = (int)((q / p.pfks_l) % (big + 1)); const int j = (int)(q / per_key);
                const uint64_t kb = (i < big) ? S[i] : ~0ull;           // "-1" for the body position
                const uint64_t g = 1ull << (64 - p.pfks_b * (s + 1));
                if (j < k) { const uint64_t f = (0ull - kb) * g; for (int t = 0; t < N; t++) m[t] = S[(size_t)j * N + t] * f; }
                else       { std::fill(m.begin(), m.end(), 0ull); m[0] = kb * g; }
                glwe_encrypt(p, S, m.data(), p.s_pfks, key, (uint64_t)q, ks->pfpksk.data() + (size_t)q * G * N);
            }
        }
    }
    ks->build_fourier();
    return ks;
}

KeySet* keyset_from_raw(const Params& p, const uint64_t* sk_glwe, const uint64_t* sk_lwe, const uint64_t* bsk,
                        const uint64_t* ksk, const uint64_t* pfpksk) {
    KeySet* ks = new KeySet();
    ks->p = p;
    ks->sk_glwe.assign(sk_glwe, sk_glwe + p.big());
    ks->sk_lwe.assign(sk_lwe, sk_lwe + p.n);
    ks->bsk.assign(bsk, bsk + ks->bsk_len());
    ks->ksk.assign(ksk, ksk + ks->ksk_len());
    ks->pfpksk.assign(pfpksk, pfpksk + ks->pfpksk_len());
    ks->build_fourier();
    return ks;
}

// reference ClientKey::encrypt shortint_woppbs_1bit.rs:200-217 (big key, sigma_lwe)
void encrypt_bit(const KeySet& ks, int bit, uint64_t index, uint64_t* out) {
    uint8_t key[32]; rng_key(ks.seed, D_CLIENT, key);
    lwe_encrypt(ks.sk_glwe.data(), ks.p.big(), encode_bit((uint64_t)bit), ks.p.s_lwe, key, index, out);
}
// reference ClientKey::decrypt :219-225 (phase only; decode_bit applied by the caller)
uint64_t decrypt_phase(const KeySet& ks, const uint64_t* ct) {
    const int big = ks.p.big();
    uint64_t acc = 0;
    for (int i = 0; i < big; i++) acc += ct[i] * ks.sk_glwe[i];
    return ct[big] - acc;
}

// ============================================================================================ stages
// [U] lwe_keyswitch.rs::keyswitch_lwe_ciphertext  (reference call: shortint_woppbs_1bit.rs:342-349, extract_bits with 1 bit)
void keyswitch(const KeySet& ks, const uint64_t* in, uint64_t* out) {
    const Params& p = ks.p; const int big = p.big(), n1 = p.n + 1;
    std::fill(out, out + n1, 0ull);
    out[p.n] = in[big];
    for (int i = 0; i < big; i++) {
        uint64_t st = decomp_init_state(in[i], p.ks_b, p.ks_l);
        const uint64_t* blk = ks.ksk.data() + (size_t)i * p.ks_l * n1;
        for (int s = 0; s < p.ks_l; s++) {
            const uint64_t d = (uint64_t)decomp_next(st, p.ks_b);
            if (d == 0) continue;
            const uint64_t* row = blk + (size_t)s * n1;
            for (int t = 0; t < n1; t++) out[t] -= d * row[t];
        }
    }
}

// [U] fft64/crypto/ggsw.rs::add_external_product_assign
void external_product_add(const NegFFT& f, int k, int levels, int b, const double* g_re, const double* g_im,
                          const uint64_t* glwe_in, uint64_t* acc) {
    const int N = f.N, M = f.M, G = k + 1;
    std::vector<uint64_t> st((size_t)G * N);
    std::vector<int64_t> dig(N);
    std::vector<double> fr(M), fi(M), o_re((size_t)G * M, 0.0), o_im((size_t)G * M, 0.0);
    for (int t = 0; t < G * N; t++) st[t] = decomp_init_state(glwe_in[t], b, levels);
    for (int lev = levels; lev >= 1; lev--) {
        const int s = lev - 1;
        for (int r = 0; r < G; r++) {
            for (int t = 0; t < N; t++) dig[t] = decomp_next(st[(size_t)r * N + t], b);
            f.fwd_int(dig.data(), fr.data(), fi.data());
            for (int c = 0; c < G; c++) {
                const double* gr = g_re + ((size_t)(s * G + r) * G + c) * M;
                const double* gi = g_im + ((size_t)(s * G + r) * G + c) * M;
                double* orr = o_re.data() + (size_t)c * M; double* oi = o_im.data() + (size_t)c * M;
                for (int j = 0; j < M; j++) {
                    orr[j] += fr[j] * gr[j] - fi[j] * gi[j];
                    oi[j] += fr[j] * gi[j] + fi[j] * gr[j];
                }
            }
        }
    }
    for (int c = 0; c < G; c++) f.add_bwd_torus(acc + (size_t)c * N, o_re.data() + (size_t)c * M, o_im.data() + (size_t)c * M);
}

// [U] lwe_programmable_bootstrapping pbs_modulus_switch
static inline int modswitch(uint64_t a, int logN) {
    const int lg = logN + 1;
    return (int)((a + (1ull << (64 - lg - 1))) >> (64 - lg));
}
static inline int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }

// [U] glwe_sample_extraction.rs::extract_lwe_sample_from_glwe_ciphertext(.., MonomialDegree(0))
void sample_extract0(int k, int N, const uint64_t* glwe, uint64_t* lwe) {
    for (int i = 0; i < k; i++) {
        const uint64_t* a = glwe + (size_t)i * N; uint64_t* o = lwe + (size_t)i * N;
        o[0] = a[0];
        for (int j = 1; j < N; j++) o[j] = 0ull - a[N - j];
    }
    lwe[(size_t)k * N] = glwe[(size_t)k * N];
}

// [U] fft64/crypto/wop_pbs.rs::homomorphic_shift_boolean (delta_log = 63 → pre-shift multiplier 1) wrapping
//     fft64/crypto/bootstrap.rs::{blind_rotate_assign, bootstrap}
void pbs_shift_boolean(const KeySet& ks, const uint64_t* in_small, uint64_t* out_big) {
    pbs_sign(ks, in_small, 1ull << (63 - ks.p.cbs_b * ks.p.cbs_l), out_big);
}
// the bootstrap both callers share: accumulator = trivial GLWE with −alpha in every coefficient, input body + 2^62 (centres
// the error on the negacyclic step), output body + alpha  ⇒  LWE of (sign bit of the input)·2·alpha under the big key
void pbs_sign(const KeySet& ks, const uint64_t* in_small, uint64_t alpha, uint64_t* out_big) {
    const Params& p = ks.p; const int N = p.N, k = p.k, G = k + 1, M = N / 2, logN = ilog2(N);
    std::vector<uint64_t> acc((size_t)G * N, 0ull), ct1((size_t)G * N), tmp(N);
    // accumulator = trivial GLWE, body = -alpha in every coefficient, rotated by X^{-b~}
    const uint64_t body_in = in_small[p.n] + (1ull << 62);
    const int bt = modswitch(body_in, logN);
    for (int t = 0; t < N; t++) tmp[t] = 0ull - alpha;
    monomial_mul(tmp.data(), N, (2 * N - bt) % (2 * N), acc.data() + (size_t)k * N);
    const size_t ggsw_sz = (size_t)p.pbs_l * G * G * M;
    for (int i = 0; i < p.n; i++) {
        if (in_small[i] == 0) continue;
        const int at = modswitch(in_small[i], logN);
        for (int c = 0; c < G; c++) {
            monomial_mul(acc.data() + (size_t)c * N, N, at, ct1.data() + (size_t)c * N);
            for (int t = 0; t < N; t++) ct1[(size_t)c * N + t] -= acc[(size_t)c * N + t];
        }
        external_product_add(*ks.fft, k, p.pbs_l, p.pbs_b, ks.bsk_re.data() + ggsw_sz * i, ks.bsk_im.data() + ggsw_sz * i, ct1.data(), acc.data());
    }
    sample_extract0(k, N, acc.data(), out_big);
    out_big[(size_t)k * N] += alpha;
}

// [U] fft64/crypto/wop_pbs.rs::extract_bits — the general bit-extraction chain (the AES path calls it with delta_log = 63 and
// one bit, where it degenerates to the keyswitch alone; reference shortint_woppbs_1bit.rs:342-349; the 8-bit model extracts 8
// bits at delta_log 56, shortint_woppbs_8bit.rs:271-275).  For bit i = 0 … n_bits−1, least significant first:
//   shifted = remaining · 2^(64 − delta_log − i − 1)            the bit becomes the most significant one
//   out[n_bits − 1 − i] = keyswitch(shifted)                     small key; the list ends up most significant bit first
//   (last bit: done)  bit_big = pbs_sign(keyswitch output, alpha = 2^(delta_log + i − 1))      = bit · 2^(delta_log + i)
//   remaining −= bit_big                                         clears the bit for the next round
void extract_bits(const KeySet& ks, const uint64_t* in_big, int delta_log, int n_bits, uint64_t* out_small /* [n_bits][n+1] */) {
    const Params& p = ks.p; const int L = p.big() + 1, S = p.n + 1;
    std::vector<uint64_t> rem(in_big, in_big + L), shifted(L), small(S), bit_big(L);
    for (int i = 0; i < n_bits; i++) {
        const int shift = 64 - delta_log - i - 1;
        for (int t = 0; t < L; t++) shifted[t] = rem[t] << shift;
        keyswitch(ks, shifted.data(), small.data());
        memcpy(out_small + (size_t)(n_bits - 1 - i) * S, small.data(), sizeof(uint64_t) * S);
        if (i == n_bits - 1) break;
        pbs_sign(ks, small.data(), 1ull << (delta_log + i - 1), bit_big.data());
        for (int t = 0; t < L; t++) rem[t] -= bit_big[t];
    }
}

// [U] lwe_private_functional_packing_keyswitch.rs::private_functional_keyswitch_lwe_ciphertext_into_glwe_ciphertext, for all k+1 keys
__attribute__((target_clones("arch=skylake-avx512", "avx2", "default")))
static void row_submul(uint64_t* out, const uint64_t* row, uint64_t d, int len) {
    for (int t = 0; t < len; t++) out[t] -= d * row[t];
}
void pfks_all(const KeySet& ks, const uint64_t* in_big, uint64_t* out) {
    const Params& p = ks.p; const int big = p.big(), G = p.k + 1, W = G * p.N;
    std::fill(out, out + (size_t)G * W, 0ull);
    std::vector<uint64_t> digs((size_t)(big + 1) * p.pfks_l);
    for (int i = 0; i <= big; i++) {
        const uint64_t rounded = closest_representable(in_big[i], p.pfks_b, p.pfks_l);
        uint64_t st = decomp_init_state(rounded, p.pfks_b, p.pfks_l);
        // decomposition yields level l first; key block stores level 1 first and is iterated reversed
        for (int lev = p.pfks_l; lev >= 1; lev--) digs[(size_t)i * p.pfks_l + (lev - 1)] = (uint64_t)decomp_next(st, p.pfks_b);
    }
    for (int j = 0; j < G; j++) {
        uint64_t* o = out + (size_t)j * W;
        const uint64_t* key = ks.pfpksk.data() + (size_t)j * (big + 1) * p.pfks_l * W;
        for (int i = 0; i <= big; i++)
            for (int s = 0; s < p.pfks_l; s++) {
                const uint64_t d = digs[(size_t)i * p.pfks_l + s];
                if (d) row_submul(o, key + ((size_t)i * p.pfks_l + s) * W, d, W);
            }
    }
}

// [U] wop_pbs.rs::circuit_bootstrap_boolean (cbs level index 0 ↔ decomposition level 1; the four shipped sets have cbs_l = 1)
void circuit_bootstrap_boolean(const KeySet& ks, const uint64_t* in_small, uint64_t* ggsw_std) {
    const Params& p = ks.p; const int G = p.k + 1, W = G * p.N;
    std::vector<uint64_t> lwe(p.big() + 1);
    // one bootstrap + (k+1) private functional keyswitches per level: level matrix s of the GGSW carries bit · q / B^(s+1)
    for (int lv = 1; lv <= p.cbs_l; lv++) {
        pbs_sign(ks, in_small, 1ull << (63 - p.cbs_b * lv), lwe.data());
        pfks_all(ks, lwe.data(), ggsw_std + (size_t)(lv - 1) * G * W);
    }
}
// [U] ggsw.rs::FourierGgswCiphertext::fill_with_forward_fourier
void ggsw_to_fourier(const KeySet& ks, const uint64_t* ggsw_std, int levels, FourierGgsw& out) {
    const Params& p = ks.p; const int G = p.k + 1, M = p.N / 2;
    const size_t polys = (size_t)levels * G * G;
    out.re.resize(polys * M); out.im.resize(polys * M);
    for (size_t q = 0; q < polys; q++) ks.fft->fwd_torus(ggsw_std + q * p.N, out.re.data() + q * M, out.im.data() + q * M);
}

// [U] ggsw.rs::cmux : c1 -= c0 ; c0 += G ⊡ c1
static void cmux(const KeySet& ks, uint64_t* c0, uint64_t* c1, const FourierGgsw& g) {
    const Params& p = ks.p; const int W = (p.k + 1) * p.N;
    for (int t = 0; t < W; t++) c1[t] -= c0[t];
    external_product_add(*ks.fft, p.k, p.cbs_l, p.cbs_b, g.re.data(), g.im.data(), c1, c0);
}

// [U] wop_pbs.rs::{vertical_packing, cmux_tree_memory_optimized, blind_rotate_assign}
void vertical_packing(const KeySet& ks, const uint64_t* lut, size_t lut_len, const std::vector<FourierGgsw>& ggsws, uint64_t* out_big) {
    const Params& p = ks.p; const int N = p.N, k = p.k, W = (k + 1) * N;
    const int n_polys = (int)(lut_len / N);
    int log_polys = 0; while ((1 << (log_polys + 1)) <= n_polys) log_polys++;
    const int tree_bits = (log_polys > (int)ggsws.size()) ? 0 : log_polys;
    // CMux tree, level order: layer j uses GGSW (tree_bits-1-j); node = cmux(child 2i, child 2i+1)
    std::vector<std::vector<uint64_t>> nodes((size_t)1 << tree_bits, std::vector<uint64_t>(W, 0ull));
    for (int i = 0; i < (1 << tree_bits); i++) memcpy(nodes[i].data() + (size_t)k * N, lut + (size_t)i * N, sizeof(uint64_t) * N);
    int cnt = 1 << tree_bits;
    for (int j = 0; j < tree_bits; j++) {
        const FourierGgsw& g = ggsws[tree_bits - 1 - j];
        for (int i = 0; i < cnt / 2; i++) {
            cmux(ks, nodes[2 * i].data(), nodes[2 * i + 1].data(), g);
            if (i != 2 * i) nodes[i].swap(nodes[2 * i]);
        }
        cnt /= 2;
    }
    std::vector<uint64_t>& T = nodes[0];
    std::vector<uint64_t> c1(W);
    int deg = 1;
    for (int gi = (int)ggsws.size() - 1; gi >= tree_bits; gi--) {
        for (int c = 0; c <= k; c++) monomial_mul(T.data() + (size_t)c * N, N, (2 * N - deg) % (2 * N), c1.data() + (size_t)c * N);
        deg <<= 1;
        cmux(ks, T.data(), c1.data(), ggsws[gi]);
    }
    sample_extract0(k, N, T.data(), out_big);
}

size_t lut_len_per_output(int n_in, int N) {
    const int logN = ilog2(N);
    const int tree_bits = n_in > logN ? n_in - logN : 0;
    return (size_t)N << tree_bits;
}
// reference generate_multivariate_luts, shortint_woppbs_1bit.rs:366-403
void generate_lut(int n_in, int n_out, int N, const uint64_t* f_table, uint64_t* out) {
    const size_t len = lut_len_per_output(n_in, N);
    std::fill(out, out + len * n_out, 0ull);
    for (int o = 0; o < n_out; o++)
        for (size_t val = 0; val < ((size_t)1 << n_in); val++) {
            // util::u64_to_bits(f(val))[o + 64 - n_out], MSB first (util.rs:54-62)
            const uint64_t bit = (f_table[val] >> (n_out - 1 - o)) & 1ull;
            out[(size_t)o * len + val] = encode_bit(bit);
        }
}

// reference FheContext::circuit_bootstrap, shortint_woppbs_1bit.rs:292-336
void circuit_bootstrap(const KeySet& ks, const uint64_t* in_bits, int n_in, const uint64_t* lut, int n_out, uint64_t* out) {
    const Params& p = ks.p; const int big1 = p.big() + 1, G = p.k + 1, W = G * p.N;
    std::vector<FourierGgsw> ggsws(n_in);
    std::vector<uint64_t> small(p.n + 1), ggsw_std((size_t)p.cbs_l * G * W);
    for (int i = 0; i < n_in; i++) {
        keyswitch(ks, in_bits + (size_t)i * big1, small.data());                 // extract_dual_bit_from_bit :339-363
        circuit_bootstrap_boolean(ks, small.data(), ggsw_std.data());
        ggsw_to_fourier(ks, ggsw_std.data(), p.cbs_l, ggsws[i]);
    }
    const size_t len = lut_len_per_output(n_in, p.N);
    for (int o = 0; o < n_out; o++) vertical_packing(ks, lut + (size_t)o * len, len, ggsws, out + (size_t)o * big1);
}

// ============================================================================================ clear AES
// reference src/aes_128.rs:18-56
const uint8_t SBOX[256] = {
    0x63, 0x7c, 0x77, 0x7b, 0xf2, 0x6b, 0x6f, 0xc5, 0x30, 0x01, 0x67, 0x2b, 0xfe, 0xd7, 0xab, 0x76, 0xca, 0x82, 0xc9, 0x7d, 0xfa, 0x59, 0x47,
    0xf0, 0xad, 0xd4, 0xa2, 0xaf, 0x9c, 0xa4, 0x72, 0xc0, 0xb7, 0xfd, 0x93, 0x26, 0x36, 0x3f, 0xf7, 0xcc, 0x34, 0xa5, 0xe5, 0xf1, 0x71, 0xd8,
    0x31, 0x15, 0x04, 0xc7, 0x23, 0xc3, 0x18, 0x96, 0x05, 0x9a, 0x07, 0x12, 0x80, 0xe2, 0xeb, 0x27, 0xb2, 0x75, 0x09, 0x83, 0x2c, 0x1a, 0x1b,
    0x6e, 0x5a, 0xa0, 0x52, 0x3b, 0xd6, 0xb3, 0x29, 0xe3, 0x2f, 0x84, 0x53, 0xd1, 0x00, 0xed, 0x20, 0xfc, 0xb1, 0x5b, 0x6a, 0xcb, 0xbe, 0x39,
    0x4a, 0x4c, 0x58, 0xcf, 0xd0, 0xef, 0xaa, 0xfb, 0x43, 0x4d, 0x33, 0x85, 0x45, 0xf9, 0x02, 0x7f, 0x50, 0x3c, 0x9f, 0xa8, 0x51, 0xa3, 0x40,
    0x8f, 0x92, 0x9d, 0x38, 0xf5, 0xbc, 0xb6, 0xda, 0x21, 0x10, 0xff, 0xf3, 0xd2, 0xcd, 0x0c, 0x13, 0xec, 0x5f, 0x97, 0x44, 0x17, 0xc4, 0xa7,
    0x7e, 0x3d, 0x64, 0x5d, 0x19, 0x73, 0x60, 0x81, 0x4f, 0xdc, 0x22, 0x2a, 0x90, 0x88, 0x46, 0xee, 0xb8, 0x14, 0xde, 0x5e, 0x0b, 0xdb, 0xe0,
    0x32, 0x3a, 0x0a, 0x49, 0x06, 0x24, 0x5c, 0xc2, 0xd3, 0xac, 0x62, 0x91, 0x95, 0xe4, 0x79, 0xe7, 0xc8, 0x37, 0x6d, 0x8d, 0xd5, 0x4e, 0xa9,
    0x6c, 0x56, 0xf4, 0xea, 0x65, 0x7a, 0xae, 0x08, 0xba, 0x78, 0x25, 0x2e, 0x1c, 0xa6, 0xb4, 0xc6, 0xe8, 0xdd, 0x74, 0x1f, 0x4b, 0xbd, 0x8b,
    0x8a, 0x70, 0x3e, 0xb5, 0x66, 0x48, 0x03, 0xf6, 0x0e, 0x61, 0x35, 0x57, 0xb9, 0x86, 0xc1, 0x1d, 0x9e, 0xe1, 0xf8, 0x98, 0x11, 0x69, 0xd9,
    0x8e, 0x94, 0x9b, 0x1e, 0x87, 0xe9, 0xce, 0x55, 0x28, 0xdf, 0x8c, 0xa1, 0x89, 0x0d, 0xbf, 0xe6, 0x42, 0x68, 0x41, 0x99, 0x2d, 0x0f, 0xb0,
    0x54, 0xbb, 0x16};
const uint8_t RC[11] = {0x00, 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};

uint8_t gf_256_mul(uint8_t a, uint8_t b) {   // aes_128.rs:42-56
    uint8_t res = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) res ^= a;
        const uint8_t hi = a & 0x80;
        a <<= 1;
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return res;
}
// reference src/aes_128/plain.rs:105-136 (words stored as 4 consecutive bytes)
void plain_key_schedule(const uint8_t key[16], uint8_t ek[176]) {
    memcpy(ek, key, 16);
    for (int i = 4; i < 44; i++) {
        uint8_t w[4]; memcpy(w, ek + 4 * (i - 1), 4);
        if (i % 4 == 0) {
            const uint8_t t = w[0]; w[0] = SBOX[w[1]]; w[1] = SBOX[w[2]]; w[2] = SBOX[w[3]]; w[3] = SBOX[t];
            w[0] ^= RC[i / 4];
        }
        for (int b = 0; b < 4; b++) ek[4 * i + b] = ek[4 * (i - 4) + b] ^ w[b];
    }
}
// reference src/aes_128/plain.rs:75-103 — note the final AddRoundKey always uses words 40..44, even if rounds < 10
void plain_encrypt_block(const uint8_t ek[176], const uint8_t in[16], int rounds, uint8_t out[16]) {
    uint8_t s[16]; for (int i = 0; i < 16; i++) s[i] = in[i] ^ ek[i];
    auto sub_shift = [&](uint8_t* st) {
        uint8_t t[16];
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) t[4 * c + r] = SBOX[st[4 * ((c + r) % 4) + r]];
        memcpy(st, t, 16);
    };
    for (int rd = 1; rd < rounds; rd++) {
        sub_shift(s);
        uint8_t t[16];
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++)
            t[4 * c + r] = gf_256_mul(s[4 * c + r], 2) ^ gf_256_mul(s[4 * c + (r + 3) % 4], 1) ^ gf_256_mul(s[4 * c + (r + 2) % 4], 1) ^ gf_256_mul(s[4 * c + (r + 1) % 4], 3);
        for (int i = 0; i < 16; i++) s[i] = t[i] ^ ek[16 * rd + i];
    }
    sub_shift(s);
    for (int i = 0; i < 16; i++) out[i] = s[i] ^ ek[160 + i];
}

// ============================================================================================ FHE AES
// LUT closures of reference fhe_impls/shortint_woppbs_1bit.rs:18-44, :94-129
static void make_aes_luts(int N, std::vector<uint64_t>& lut24, std::vector<uint64_t>& lut8, std::vector<uint64_t>& lut1) {
    std::vector<uint64_t> f(256);
    for (int b = 0; b < 256; b++) f[b] = ((uint64_t)gf_256_mul(SBOX[b], 1) << 16) | ((uint64_t)gf_256_mul(SBOX[b], 2) << 8) | (uint64_t)gf_256_mul(SBOX[b], 3);
    lut24.resize(lut_len_per_output(8, N) * 24); generate_lut(8, 24, N, f.data(), lut24.data());
    for (int b = 0; b < 256; b++) f[b] = SBOX[b];
    lut8.resize(lut_len_per_output(8, N) * 8); generate_lut(8, 8, N, f.data(), lut8.data());
    uint64_t id[2] = {0, 1};
    lut1.resize(lut_len_per_output(1, N)); generate_lut(1, 1, N, id, lut1.data());
}

static inline void lwe_add(uint64_t* a, const uint64_t* b, int len) { for (int t = 0; t < len; t++) a[t] += b[t]; }

// reference fhe_sbox_gal_mul_pbs.rs:84-132 (encrypt_block_for_rounds) with data_model.rs:165-281 re-expressed on a flat
// [16 bytes][8 bits][big+1] tensor (byte index = 4*col + row); noise bookkeeping of shortint_woppbs_1bit.rs:63-77 done statically.
int aes_encrypt_blocks(const KeySet& ks, const uint64_t* key_sched, int n_blocks, int rounds, const uint64_t* in, uint64_t* out, int in_noise_sq) {
    const Params& p = ks.p; const int L = p.big() + 1; const size_t BYTE = (size_t)8 * L, BLK = 16 * BYTE;
    std::vector<uint64_t> lut24, lut8, lut1; make_aes_luts(p.N, lut24, lut8, lut1);
    // static noise check (squared levels): rk = 1 (fresh or boot output)
    {
        int lvl = in_noise_sq + 1;
        if (lvl > p.max_noise_sq) return -1;
        for (int rd = 1; rd < rounds; rd++) { lvl = 8 * 4 + 1; if (lvl > p.max_noise_sq) return -1; }
        lvl = 8 + 1; if (lvl > p.max_noise_sq) return -1;
    }
    std::vector<uint64_t> state((size_t)n_blocks * BLK);
    memcpy(state.data(), in, sizeof(uint64_t) * state.size());
    for (int b = 0; b < n_blocks; b++) lwe_add(state.data() + (size_t)b * BLK, key_sched, (int)BLK);          // xor_state(rk[0..4]) :96-99
    std::vector<uint64_t> muls((size_t)n_blocks * 16 * 24 * L);
    for (int rd = 1; rd < rounds; rd++) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int q = 0; q < n_blocks * 16; q++)                                                                 // sub_bytes_with_gal_mul :27-48
            circuit_bootstrap(ks, state.data() + (size_t)q * BYTE, 8, lut24.data(), 24, muls.data() + (size_t)q * 24 * L);
        // shift_rows on the three states (:106-108) + mix_columns (:61-82) + xor_state (:112-117)
        for (int b = 0; b < n_blocks; b++) {
            const uint64_t* mb = muls.data() + (size_t)b * 16 * 24 * L;
            uint64_t* sb = state.data() + (size_t)b * BLK;
            auto src = [&](int which, int row, int col) {   // byte after ShiftRows: new[row][col] = old[row][(col+row)%4]
                const int old_byte = 4 * ((col + row) % 4) + row;
                return mb + ((size_t)old_byte * 24 + (size_t)which * 8) * L;
            };
            for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) {
                uint64_t* dst = sb + (size_t)(4 * c + r) * BYTE;
                memcpy(dst, src(1, r, c), sizeof(uint64_t) * BYTE);               // mul2[i]
                lwe_add(dst, src(0, (r + 3) % 4, c), (int)BYTE);                   // ^ mul1[i-1]
                lwe_add(dst, src(0, (r + 2) % 4, c), (int)BYTE);                   // ^ mul1[i-2]
                lwe_add(dst, src(2, (r + 1) % 4, c), (int)BYTE);                   // ^ mul3[i-3]
            }
            lwe_add(sb, key_sched + (size_t)rd * BLK, (int)BLK);
        }
    }
    // last round :119-129 — sub_bytes, shift_rows, xor rk[40..44]
    std::vector<uint64_t> sub((size_t)n_blocks * BLK);
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < n_blocks * 16; q++) circuit_bootstrap(ks, state.data() + (size_t)q * BYTE, 8, lut8.data(), 8, sub.data() + (size_t)q * BYTE);
    for (int b = 0; b < n_blocks; b++) {
        uint64_t* ob = out + (size_t)b * BLK;
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++)
            memcpy(ob + (size_t)(4 * c + r) * BYTE, sub.data() + (size_t)b * BLK + (size_t)(4 * ((c + r) % 4) + r) * BYTE, sizeof(uint64_t) * BYTE);
        lwe_add(ob, key_sched + (size_t)10 * BLK, (int)BLK);
    }
    return 0;
}

// reference fhe_sbox_gal_mul_pbs.rs:134-191 (key_schedule, boot_word, sub_word)
int aes_key_schedule(const KeySet& ks, const uint64_t* key_bits, uint64_t* ek) {
    const Params& p = ks.p; const int L = p.big() + 1; const size_t BYTE = (size_t)8 * L, WORD = 4 * BYTE;
    std::vector<uint64_t> lut24, lut8, lut1; make_aes_luts(p.N, lut24, lut8, lut1);
    if (9 > p.max_noise_sq) return -1;
    memcpy(ek, key_bits, sizeof(uint64_t) * 4 * WORD);
    std::vector<uint64_t> tmp(WORD), w(WORD);
    for (int i = 4; i < 44; i++) {
        uint64_t* cur = ek + (size_t)i * WORD;
        const uint64_t* prev = ek + (size_t)(i - 1) * WORD;
        if (i % 4 == 0) {
#pragma omp parallel for schedule(dynamic, 1)
            for (int b = 0; b < 4; b++)                                    // sub_word(rotate_left(1))
                circuit_bootstrap(ks, prev + (size_t)((b + 1) % 4) * BYTE, 8, lut8.data(), 8, tmp.data() + (size_t)b * BYTE);
            for (int bit = 0; bit < 8; bit++)                              // ^= trivial(RC[i/4]) on byte 0 (:154), MSB first
                if (RC[i / 4] & (0x80 >> bit)) tmp[(size_t)bit * L + (L - 1)] += encode_bit(1);
        } else {
            memcpy(tmp.data(), prev, sizeof(uint64_t) * WORD);
        }
        memcpy(w.data(), ek + (size_t)(i - 4) * WORD, sizeof(uint64_t) * WORD);
        lwe_add(w.data(), tmp.data(), (int)WORD);
#pragma omp parallel for schedule(dynamic, 1)
        for (int q = 0; q < 32; q++)                                       // boot_word :166-180 → bootstrap_assign (1→1 identity)
            circuit_bootstrap(ks, w.data() + (size_t)q * L, 1, lut1.data(), 1, cur + (size_t)q * L);
    }
    return 0;
}

}  // namespace orc
