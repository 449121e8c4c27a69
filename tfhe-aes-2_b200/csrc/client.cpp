// client.cpp — client side of the model (host CPU, as in the reference): secret keys, evaluation-key generation,
// bit encryption / decryption, and the pure-integer helpers (parameter presets, encodings, LUT generation).
//
// Reference: src/tfhe/shortint_woppbs_1bit.rs:189-268 (ClientKey, generate_keys_with_params), :125-132 (encodings),
// :366-403 (generate_multivariate_luts); parameters.rs.  Key material formats are those of tfhe 0.11.2's core_crypto
// (layouts in include/tfhe_aes_cuda.h).  Randomness: the reference seeds tfhe-csprng from the OS (engine.rs:164-168);
// here every object (one GLWE / LWE ciphertext) draws from its own ChaCha20 stream keyed by (seed, domain) with the
// object index as nonce, so generation is reproducible and order-independent across threads.
#include "../../include/tfhe_aes_cuda.h"
#include "tac_common.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <cstdio>
#include <sys/random.h>
#include <vector>

namespace {

// ---------------------------------------------------------------- ChaCha20 block function (64-bit counter ‖ 64-bit nonce)
struct Stream {
    uint32_t in[16];
    uint32_t out[16];
    int used = 16;
    // master: 256-bit ChaCha key of this client (OS entropy, or derived from a test seed); bytes 8..11 separate the domains
    Stream(const uint8_t master[32], uint32_t domain, uint64_t nonce) {
        static const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        uint8_t key[32];
        memcpy(key, master, 32);
        for (int i = 0; i < 4; i++) key[8 + i] ^= (uint8_t)(domain >> (8 * i));
        memcpy(in, sigma, 16);
        memcpy(in + 4, key, 32);          // little-endian host
        in[12] = 0; in[13] = 0;
        in[14] = (uint32_t)nonce; in[15] = (uint32_t)(nonce >> 32);
    }
    static inline uint32_t rl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    static inline void quarter(uint32_t* x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 7);
    }
    void block() {
        uint32_t x[16];
        memcpy(x, in, 64);
        for (int round = 0; round < 20; round += 2) {
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
        if (++in[12] == 0) ++in[13];
        used = 0;
    }
    inline uint64_t u64() {
        if (used == 16) block();
        const uint64_t lo = out[used], hi = out[used + 1];
        used += 2;
        return lo | (hi << 32);
    }
};
enum : uint32_t { DOM_SK_GLWE = 1, DOM_SK_LWE = 2, DOM_BSK = 3, DOM_KSK = 4, DOM_PFPKSK = 5, DOM_CLIENT = 6 };

const double kTwo64 = 18446744073709551616.0;

// two standard normal samples from two 64-bit draws (Box–Muller)
inline void normal_pair(uint64_t x, uint64_t y, double& z0, double& z1) {
    const double u1 = (double)((x >> 11) + 1) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(y >> 11) * (1.0 / 9007199254740992.0);
    const double r = std::sqrt(-2.0 * std::log(u1));
    const double th = 6.283185307179586476925286766559 * u2;
    z0 = r * std::cos(th);
    z1 = r * std::sin(th);
}
inline uint64_t torus_noise(double z, double sigma_times_2_64) {
    const double v = z * sigma_times_2_64;
    return (uint64_t)(int64_t)std::llrint(v);
}

template <class F>
void parallel_for(long count, int threads, F f) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<long>(threads, std::max<long>(1, count));
    std::atomic<long> next(0);
    auto worker = [&]() {
        for (;;) {
            const long begin = next.fetch_add(32);
            if (begin >= count) return;
            const long end = std::min(count, begin + 32);
            for (long q = begin; q < end; q++) f(q);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
}

// Deterministic (test / benchmark) master key: seed ‖ 0000 ‖ fixed label.  Anyone who knows the seed can recompute the
// secret keys, so this mode is for reproducible tests only; tac_client_keygen_os draws all 256 bits from the OS.
void master_from_seed(uint64_t seed, uint8_t master[32]) {
    for (int i = 0; i < 8; i++) master[i] = (uint8_t)(seed >> (8 * i));
    memset(master + 8, 0, 4);
    memcpy(master + 12, "tfhe-aes-b200 rng v1", 20);
}
bool master_from_os(uint8_t master[32]) {
    size_t got = 0;
    while (got < 32) {
        const ssize_t r = getrandom(master + got, 32 - got, 0);
        if (r <= 0) break;
        got += (size_t)r;
    }
    if (got == 32) return true;
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return false;
    const bool ok = fread(master, 1, 32, f) == 32;
    fclose(f);
    return ok;
}

struct Client {
    TacParams p;
    uint8_t seed[32];              // master key of every random stream of this client
    std::vector<uint64_t> sk_glwe, sk_lwe, bsk, ksk, pfpksk;
    std::vector<std::vector<int>> glwe_support;   // positions of the 1-bits of each GLWE key polynomial
    bool have_eval = false;
    int big() const { return p.k * p.N; }

    // body += Σ_i a_i · S_i (negacyclic), using the sparse support of the binary key polynomials
    void add_mask_times_key(const uint64_t* mask, uint64_t* body) const {
        const int N = p.N;
        for (int i = 0; i < p.k; i++) {
            const uint64_t* a = mask + (size_t)i * N;
            for (int pos : glwe_support[i]) {
                // a · X^pos : coefficient t receives a[t-pos] (t >= pos) or -a[t-pos+N] (t < pos)
                const uint64_t* hi = a + (N - pos);
                for (int t = 0; t < pos; t++) body[t] -= hi[t];
                uint64_t* dst = body + pos;
                for (int t = 0; t < N - pos; t++) dst[t] += a[t];
            }
        }
    }
    // GLWE encryption of message polynomial `msg` ([U] glwe_encryption.rs::encrypt_glwe_ciphertext)
    void glwe_encrypt(const uint64_t* msg, double sigma, uint32_t domain, uint64_t index, uint64_t* out) const {
        Stream rng(seed, domain, index);
        const int N = p.N, kN = p.k * p.N;
        for (int i = 0; i < kN; i++) out[i] = rng.u64();
        uint64_t* body = out + kN;
        const double s = sigma * kTwo64;
        for (int t = 0; t < N; t += 2) {
            const uint64_t x = rng.u64(), y = rng.u64();
            double z0, z1;
            normal_pair(x, y, z0, z1);
            body[t] = msg[t] + torus_noise(z0, s);
            body[t + 1] = msg[t + 1] + torus_noise(z1, s);
        }
        add_mask_times_key(out, body);
    }
    // LWE encryption under key `s` ([U] lwe_encryption.rs::encrypt_lwe_ciphertext)
    void lwe_encrypt(const uint64_t* s, int dim, uint64_t msg, double sigma, uint32_t domain, uint64_t index, uint64_t* out) const {
        Stream rng(seed, domain, index);
        uint64_t dot = 0;
        for (int i = 0; i < dim; i++) { const uint64_t a = rng.u64(); out[i] = a; dot += a * s[i]; }
        const uint64_t x = rng.u64(), y = rng.u64();
        double z0, z1;
        normal_pair(x, y, z0, z1);
        out[dim] = dot + msg + torus_noise(z0, sigma * kTwo64);
    }
};

size_t key_len(const TacParams& p, int which) {
    const size_t big = (size_t)p.k * p.N, G = (size_t)p.k + 1;
    switch (which) {
        case 0: return big;
        case 1: return (size_t)p.n;
        case 2: return (size_t)p.n * p.pbs_l * G * G * p.N;
        case 3: return big * p.ks_l * ((size_t)p.n + 1);
        case 4: return G * (big + 1) * p.pfks_l * G * p.N;
        default: return 0;
    }
}

}  // namespace

extern "C" {

int tac_params_preset(int id, tac_params* out) {
    // reference src/tfhe/shortint_woppbs_1bit/parameters.rs:29-61, :77-109, :125-157, :173-205
    const double s_lwe = 4.7280002450549286e-05, s_n1024 = 3.162026630747649e-16, s_n512 = 0.00000000000000022148688116005568;
    switch (id) {
        case 1:   *out = tac_params{671, 2, 1024, 2, 15, 4, 3, 1, 10, 1, 24, 1, s_lwe, s_n1024, s_n1024}; return TAC_OK;
        case 4:   *out = tac_params{679, 2, 1024, 2, 15, 4, 3, 1, 11, 2, 16, 2 * 2, s_lwe, s_n1024, s_n1024}; return TAC_OK;
        case 64:  *out = tac_params{677, 4, 512, 3, 12, 4, 3, 1, 13, 2, 16, 8 * 8, s_lwe, s_n512, s_n512}; return TAC_OK;
        case 256: *out = tac_params{665, 2, 1024, 4, 9, 6, 2, 1, 14, 3, 12, 16 * 16, s_lwe, s_n1024, s_n1024}; return TAC_OK;
        default: return TAC_ERR_ARG;
    }
}

uint64_t tac_encode_bit(uint64_t bit) { return bit << 63; }                                               // :125-128
uint64_t tac_decode_bit(uint64_t e) { return ((e + (1ull << 62)) & (1ull << 63)) >> 63; }                  // :130-132

size_t tac_lut_len(int n_in, int N) {
    const int logN = tac::ilog2(N);
    return (size_t)N << (n_in > logN ? n_in - logN : 0);                                                  // :373-378
}
int tac_generate_lut(int n_in, int n_out, int N, const uint64_t* f_table, uint64_t* out) {
    if (n_in <= 0 || n_in > 16 || n_out <= 0 || n_out > 64 || (N & (N - 1))) return TAC_ERR_ARG;          // :372-376
    const size_t len = tac_lut_len(n_in, N);
    memset(out, 0, sizeof(uint64_t) * len * (size_t)n_out);
    for (int o = 0; o < n_out; o++) {
        uint64_t* small = out + (size_t)o * len;
        const int shift = n_out - 1 - o;           // util::u64_to_bits(f(val))[o + 64 - n_out], MSB first (:395-397)
        for (size_t val = 0; val < ((size_t)1 << n_in); val++) small[val] = tac_encode_bit((f_table[val] >> shift) & 1ull);
    }
    return TAC_OK;
}

size_t tac_key_len(const tac_params* p, int which) { return key_len(*reinterpret_cast<const TacParams*>(p), which); }

static void index_glwe_support(Client* c) {
    c->glwe_support.assign(c->p.k, {});
    for (int i = 0; i < c->p.k; i++)
        for (int t = 0; t < c->p.N; t++)
            if (c->sk_glwe[(size_t)i * c->p.N + t]) c->glwe_support[i].push_back(t);
}
static tac_client_key* client_new(const tac_params* pp, const uint8_t master[32]) {
    Client* c = new Client();
    c->p = *reinterpret_cast<const TacParams*>(pp);
    memcpy(c->seed, master, 32);
    c->sk_glwe.resize(c->big());
    c->sk_lwe.resize(c->p.n);
    { Stream r(c->seed, DOM_SK_GLWE, 0); for (auto& w : c->sk_glwe) w = r.u64() & 1ull; }
    { Stream r(c->seed, DOM_SK_LWE, 0); for (auto& w : c->sk_lwe) w = r.u64() & 1ull; }
    index_glwe_support(c);
    return reinterpret_cast<tac_client_key*>(c);
}
tac_client_key* tac_client_keygen(const tac_params* pp, uint64_t seed) {
    uint8_t master[32];
    master_from_seed(seed, master);
    return client_new(pp, master);
}
tac_client_key* tac_client_keygen_os(const tac_params* pp) {
    uint8_t master[32];
    if (!master_from_os(master)) return nullptr;
    return client_new(pp, master);
}
tac_client_key* tac_client_from_secret_keys(const tac_params* pp, const uint64_t* sk_glwe, const uint64_t* sk_lwe) {
    uint8_t master[32];
    if (!master_from_os(master)) return nullptr;               // encryption masks / noise of this instance: fresh entropy
    Client* c = new Client();
    c->p = *reinterpret_cast<const TacParams*>(pp);
    memcpy(c->seed, master, 32);
    c->sk_glwe.assign(sk_glwe, sk_glwe + c->big());
    c->sk_lwe.assign(sk_lwe, sk_lwe + c->p.n);
    for (uint64_t w : c->sk_glwe) if (w > 1) { delete c; return nullptr; }
    for (uint64_t w : c->sk_lwe) if (w > 1) { delete c; return nullptr; }
    index_glwe_support(c);
    return reinterpret_cast<tac_client_key*>(c);
}
void tac_client_free(tac_client_key* ck) { delete reinterpret_cast<Client*>(ck); }

int tac_client_gen_eval_keys(tac_client_key* ck, int threads) {
    Client& c = *reinterpret_cast<Client*>(ck);
    if (c.have_eval) return TAC_OK;
    const TacParams& p = c.p;
    const int N = p.N, k = p.k, G = k + 1, big = c.big();
    const size_t glwe_words = (size_t)G * N;
    // bootstrapping key: GGSW(s_i) — row r<k carries -s_i·g·S_r, row k carries s_i·g  ([U] ggsw_encryption.rs)
    c.bsk.assign(key_len(p, 2), 0);
    parallel_for((long)p.n * p.pbs_l * G, threads, [&](long q) {
        const int r = (int)(q % G), s = (int)((q / G) % p.pbs_l), i = (int)(q / ((long)G * p.pbs_l));
        const uint64_t factor = c.sk_lwe[i] << (64 - p.pbs_b * (s + 1));
        std::vector<uint64_t> msg(N, 0);
        if (r == k) msg[0] = factor;
        else for (int pos : c.glwe_support[r]) msg[pos] = 0ull - factor;
        c.glwe_encrypt(msg.data(), p.s_glwe, DOM_BSK, (uint64_t)q, c.bsk.data() + (size_t)q * glwe_words);
    });
    // keyswitch key big → small, level l stored first ([U] lwe_keyswitch_key_generation.rs)
    c.ksk.assign(key_len(p, 3), 0);
    parallel_for((long)big * p.ks_l, threads, [&](long q) {
        const int s = (int)(q % p.ks_l), i = (int)(q / p.ks_l);
        const uint64_t msg = c.sk_glwe[i] << (64 - p.ks_b * (p.ks_l - s));
        c.lwe_encrypt(c.sk_lwe.data(), p.n, msg, p.s_lwe, DOM_KSK, (uint64_t)q, c.ksk.data() + (size_t)q * (p.n + 1));
    });
    // circuit-bootstrap PFPKSKs: key j<k applies x ↦ -x·S_j, key k the identity; the body position uses "-1" as key bit
    // ([U] lwe_private_functional_packing_keyswitch_key_generation.rs, lwe_wopbs.rs::generate_circuit_bootstrap_lwe_pfpksk_list)
    c.pfpksk.assign(key_len(p, 4), 0);
    const long per_key = (long)(big + 1) * p.pfks_l;
    parallel_for((long)G * per_key, threads, [&](long q) {
        const int s = (int)(q % p.pfks_l), i = (int)((q / p.pfks_l) % (big + 1)), j = (int)(q / per_key);
        const uint64_t key_bit = (i < big) ? c.sk_glwe[i] : ~0ull;
        const uint64_t g = 1ull << (64 - p.pfks_b * (s + 1));
        std::vector<uint64_t> msg(N, 0);
        if (j == k) msg[0] = key_bit * g;
        else { const uint64_t f = (0ull - key_bit) * g; for (int pos : c.glwe_support[j]) msg[pos] = f; }
        c.glwe_encrypt(msg.data(), p.s_pfks, DOM_PFPKSK, (uint64_t)q, c.pfpksk.data() + (size_t)q * glwe_words);
    });
    c.have_eval = true;
    return TAC_OK;
}

const uint64_t* tac_client_key_ptr(tac_client_key* ck, int which) {
    Client& c = *reinterpret_cast<Client*>(ck);
    switch (which) {
        case 0: return c.sk_glwe.data();
        case 1: return c.sk_lwe.data();
        case 2: return c.have_eval ? c.bsk.data() : nullptr;
        case 3: return c.have_eval ? c.ksk.data() : nullptr;
        case 4: return c.have_eval ? c.pfpksk.data() : nullptr;
        default: return nullptr;
    }
}

int tac_client_encrypt_bits(tac_client_key* ck, const uint8_t* bits, size_t n, uint64_t first_index, uint64_t* out) {
    Client& c = *reinterpret_cast<Client*>(ck);
    const int big = c.big();
    for (size_t i = 0; i < n; i++)
        if (bits[i] > 1) return TAC_ERR_ARG;                              // "cleartext out of bounds" (:126)
    parallel_for((long)n, 0, [&](long i) {
        c.lwe_encrypt(c.sk_glwe.data(), big, tac_encode_bit(bits[i]), c.p.s_lwe, DOM_CLIENT, first_index + (uint64_t)i,
                      out + (size_t)i * (big + 1));
    });
    return TAC_OK;
}
int tac_client_decrypt_phases(tac_client_key* ck, const uint64_t* cts, size_t n, uint64_t* phases) {
    Client& c = *reinterpret_cast<Client*>(ck);
    const int big = c.big();
    parallel_for((long)n, 0, [&](long i) {
        const uint64_t* ct = cts + (size_t)i * (big + 1);
        uint64_t dot = 0;
        for (int t = 0; t < big; t++) dot += ct[t] * c.sk_glwe[t];
        phases[i] = ct[big] - dot;
    });
    return TAC_OK;
}
int tac_client_decrypt_bits(tac_client_key* ck, const uint64_t* cts, size_t n, uint8_t* bits) {
    std::vector<uint64_t> ph(n);
    const int rc = tac_client_decrypt_phases(ck, cts, n, ph.data());
    if (rc) return rc;
    for (size_t i = 0; i < n; i++) bits[i] = (uint8_t)tac_decode_bit(ph[i]);
    return TAC_OK;
}

}  // extern "C"
