// oracle/oracle.hpp — CPU restatement of the tfhe-aes-2 WoP-PBS hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this
// code; the CUDA product never links it.  See oracle/README.md for the parity status ("parity unpinned"
// for intermediate ciphertexts; pinned on every golden vector the reference's own tests hold).
//
// Conventions (shared with include/tfhe_aes_cuda.h so that identical keys can be fed to both sides):
//   torus            = uint64_t, wrapping arithmetic, q = 2^64
//   LWE              = mask[dim] ‖ body
//   GLWE             = k mask polynomials ‖ body polynomial, N coefficients each
//   GGSW (standard)  = [level s=0..l-1 (decomposition level s+1)] [row r=0..k] [poly c=0..k] [N]
//   BSK (standard)   = [i=0..n-1] GGSW as above
//   KSK              = [i=0..kN-1] [s=0..l-1 (decomposition level l-s)] [n+1]
//   PFPKSK           = [j=0..k] [i=0..kN (last = body)] [s=0..l-1 (decomposition level s+1)] [(k+1)N]
//   LUT              = [n_out] [N << max(0, n_in - log2 N)]
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace orc {

// Mirrors WopbsParameters + max_noise_level_squared, reference src/tfhe/shortint_woppbs_1bit/parameters.rs:9-13
struct Params {
    int32_t n;        // lwe_dimension
    int32_t k;        // glwe_dimension
    int32_t N;        // polynomial_size
    int32_t pbs_l, pbs_b;
    int32_t ks_l, ks_b;
    int32_t cbs_l, cbs_b;
    int32_t pfks_l, pfks_b;
    int32_t max_noise_sq;
    double s_lwe, s_glwe, s_pfks;
    int big() const { return k * N; }
};

bool params_preset(int id, Params* out);

// ---------------------------------------------------------------- ChaCha20 (DJB layout: 64-bit counter, 64-bit nonce)
struct ChaCha20 {
    uint32_t st[16];
    uint32_t buf[16];
    int pos;
    void init(const uint8_t key[32], uint64_t nonce, uint64_t counter = 0);
    void refill();
    inline uint32_t next_u32() { if (pos == 16) refill(); return buf[pos++]; }
    inline uint64_t next_u64() { uint64_t lo = next_u32(); uint64_t hi = next_u32(); return lo | (hi << 32); }
    void bytes(uint8_t* out, size_t n);
};

// RNG domains (shared spec with the product's client library)
enum Domain : uint32_t { D_SK_GLWE = 1, D_SK_LWE = 2, D_BSK = 3, D_KSK = 4, D_PFPKSK = 5, D_CLIENT = 6 };
void rng_key(uint64_t seed, uint32_t domain, uint8_t key[32]);

// ---------------------------------------------------------------- decomposer
uint64_t closest_representable(uint64_t x, int b, int l);
uint64_t decomp_init_state(uint64_t x, int b, int l);
inline int64_t decomp_next(uint64_t& state, int b) {
    // [U] tfhe core_crypto/commons/math/decomposition/iter.rs::decompose_one_level
    const uint64_t mask = (1ull << b) - 1;
    uint64_t res = state & mask;
    state >>= b;
    uint64_t carry = ((res - 1ull) | state) & res;
    carry >>= (b - 1);
    state += carry;
    return (int64_t)(res - (carry << b));
}

// ---------------------------------------------------------------- negacyclic FFT
struct NegFFT {
    int N, M;                       // M = N/2 complex points
    std::vector<double> tw_re, tw_im;     // twist e^{i pi j / N}, j < M
    std::vector<double> w_re, w_im;       // e^{-2 pi i k / M}, k < M/2
    explicit NegFFT(int N);
    // integer polynomial (signed digits, as int64) -> Fourier (bit-reversed order), SoA
    void fwd_int(const int64_t* p, double* re, double* im) const;
    // torus polynomial scaled by 2^-64 -> Fourier
    void fwd_torus(const uint64_t* p, double* re, double* im) const;
    // Fourier -> torus, wrapping-add into out
    void add_bwd_torus(uint64_t* out, double* re, double* im) const;   // destroys re/im
    void fft(double* re, double* im) const;      // DIF, natural -> bit-reversed
    void ifft(double* re, double* im) const;     // DIT, bit-reversed -> natural, unnormalised
};

// ---------------------------------------------------------------- keys
struct KeySet {
    Params p;
    std::vector<uint64_t> sk_glwe;   // k*N bits (0/1)
    std::vector<uint64_t> sk_lwe;    // n bits
    std::vector<uint64_t> bsk;       // standard domain
    std::vector<uint64_t> ksk;
    std::vector<uint64_t> pfpksk;
    // Fourier BSK, SoA: [i][s][r][c][M] re then im
    std::vector<double> bsk_re, bsk_im;
    NegFFT* fft = nullptr;
    uint64_t seed = 0;
    ~KeySet();
    size_t bsk_len() const { return (size_t)p.n * p.pbs_l * (p.k + 1) * (p.k + 1) * p.N; }
    size_t ksk_len() const { return (size_t)p.big() * p.ks_l * (p.n + 1); }
    size_t pfpksk_len() const { return (size_t)(p.k + 1) * (p.big() + 1) * p.pfks_l * (p.k + 1) * p.N; }
    void build_fourier();
};

KeySet* keygen(const Params& p, uint64_t seed);
KeySet* keyset_from_raw(const Params& p, const uint64_t* sk_glwe, const uint64_t* sk_lwe, const uint64_t* bsk,
                        const uint64_t* ksk, const uint64_t* pfpksk);

// client side (reference shortint_woppbs_1bit.rs:189-226)
void encrypt_bit(const KeySet& ks, int bit, uint64_t index, uint64_t* out);
uint64_t decrypt_phase(const KeySet& ks, const uint64_t* ct);
inline uint64_t encode_bit(uint64_t bit) { return bit << 63; }                                   // :125-128
inline uint64_t decode_bit(uint64_t enc) { return ((enc + (1ull << 62)) & (1ull << 63)) >> 63; }   // :130-132

// ---------------------------------------------------------------- server-side stages
void keyswitch(const KeySet& ks, const uint64_t* in_big, uint64_t* out_small);
void pbs_shift_boolean(const KeySet& ks, const uint64_t* in_small, uint64_t* out_big);
void pbs_sign(const KeySet& ks, const uint64_t* in_small, uint64_t alpha, uint64_t* out_big);
void extract_bits(const KeySet& ks, const uint64_t* in_big, int delta_log, int n_bits, uint64_t* out_small);
void pfks_all(const KeySet& ks, const uint64_t* in_big, uint64_t* ggsw_level_out /* [k+1][(k+1)N] */);
// Fourier GGSW (one per input bit): [s][r][c][M] re / im
struct FourierGgsw { std::vector<double> re, im; };
void circuit_bootstrap_boolean(const KeySet& ks, const uint64_t* in_small, uint64_t* ggsw_std /* [cbs_l][k+1][(k+1)N] */);
void ggsw_to_fourier(const KeySet& ks, const uint64_t* ggsw_std, int levels, FourierGgsw& out);
// acc(GLWE, (k+1)N) += ggsw ⊡ glwe_in   (ggsw given by SoA pointers with `levels`, `b`)
void external_product_add(const NegFFT& f, int k, int levels, int b, const double* g_re, const double* g_im,
                          const uint64_t* glwe_in, uint64_t* acc);
void vertical_packing(const KeySet& ks, const uint64_t* lut_o, size_t lut_len, const std::vector<FourierGgsw>& ggsws,
                      uint64_t* out_big);
void sample_extract0(int k, int N, const uint64_t* glwe, uint64_t* lwe);

// reference FheContext::circuit_bootstrap (shortint_woppbs_1bit.rs:292-336) minus the noise bookkeeping
void circuit_bootstrap(const KeySet& ks, const uint64_t* in_bits /*[n_in][big+1]*/, int n_in, const uint64_t* lut,
                       int n_out, uint64_t* out /*[n_out][big+1]*/);

// reference generate_multivariate_luts (shortint_woppbs_1bit.rs:366-403); f_table[val] = f(val), val < 2^n_in
size_t lut_len_per_output(int n_in, int N);
void generate_lut(int n_in, int n_out, int N, const uint64_t* f_table, uint64_t* out);

// ---------------------------------------------------------------- AES (clear + FHE)
extern const uint8_t SBOX[256];
extern const uint8_t RC[11];
uint8_t gf_256_mul(uint8_t a, uint8_t b);
void plain_key_schedule(const uint8_t key[16], uint8_t out[176]);
void plain_encrypt_block(const uint8_t ek[176], const uint8_t in[16], int rounds, uint8_t out[16]);

// FHE AES on flat ciphertext tensors: block = [16 bytes][8 bits MSB first][big+1], key schedule = [44 words][4][8][big+1]
// returns 0 on success, <0 on a noise-bookkeeping violation (mirrors the reference panics)
int aes_encrypt_blocks(const KeySet& ks, const uint64_t* key_sched, int n_blocks, int rounds, const uint64_t* in,
                       uint64_t* out, int in_noise_sq);
int aes_key_schedule(const KeySet& ks, const uint64_t* key_bits /*[16][8][big+1]*/, uint64_t* out /*[44][4][8][big+1]*/);

}  // namespace orc
