// oracle/capi.cpp — C ABI over the CPU oracle (loaded with ctypes by tests/ and bench.py's cpu_baseline leg only).
// TEST INFRASTRUCTURE ONLY — see oracle/README.md.
#include "oracle.hpp"

#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

int orc_params_preset(int id, Params* out) { return params_preset(id, out) ? 0 : -1; }
void orc_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void* orc_keygen(const Params* p, uint64_t seed) { return keygen(*p, seed); }
void* orc_keyset_from_raw(const Params* p, uint64_t seed, const uint64_t* sk_glwe, const uint64_t* sk_lwe, const uint64_t* bsk,
                          const uint64_t* ksk, const uint64_t* pfpksk) {
    KeySet* ks = keyset_from_raw(*p, sk_glwe, sk_lwe, bsk, ksk, pfpksk);
    ks->seed = seed;
    return ks;
}
void orc_keyset_free(void* ks) { delete (KeySet*)ks; }
const uint64_t* orc_key_ptr(void* h, int which, size_t* len) {
    KeySet* ks = (KeySet*)h;
    const std::vector<uint64_t>* v = nullptr;
    switch (which) {
        case 0: v = &ks->sk_glwe; break;
        case 1: v = &ks->sk_lwe; break;
        case 2: v = &ks->bsk; break;
        case 3: v = &ks->ksk; break;
        case 4: v = &ks->pfpksk; break;
        default: *len = 0; return nullptr;
    }
    *len = v->size();
    return v->data();
}

void orc_encrypt_bits(void* h, const uint8_t* bits, int n, uint64_t first_index, uint64_t* out) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) encrypt_bit(*ks, bits[i], first_index + (uint64_t)i, out + (size_t)i * L);
}
void orc_decrypt_phases(void* h, const uint64_t* cts, int n, uint64_t* phases) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1;
    for (int i = 0; i < n; i++) phases[i] = decrypt_phase(*ks, cts + (size_t)i * L);
}
void orc_decrypt_bits(void* h, const uint64_t* cts, int n, uint8_t* bits) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1;
    for (int i = 0; i < n; i++) bits[i] = (uint8_t)decode_bit(decrypt_phase(*ks, cts + (size_t)i * L));
}
// phase of a small-key LWE (after keyswitch) — test helper
void orc_decrypt_phases_small(void* h, const uint64_t* cts, int n, uint64_t* phases) {
    KeySet* ks = (KeySet*)h; const int nn = ks->p.n;
    for (int i = 0; i < n; i++) {
        const uint64_t* ct = cts + (size_t)i * (nn + 1);
        uint64_t acc = 0;
        for (int t = 0; t < nn; t++) acc += ct[t] * ks->sk_lwe[t];
        phases[i] = ct[nn] - acc;
    }
}
// phase polynomial of a GLWE under the GLWE key — test helper (N outputs per ciphertext)
void orc_glwe_phase(void* h, const uint64_t* glwe, int n, uint64_t* phases) {
    KeySet* ks = (KeySet*)h; const int N = ks->p.N, k = ks->p.k;
    for (int q = 0; q < n; q++) {
        const uint64_t* ct = glwe + (size_t)q * (k + 1) * N;
        uint64_t* ph = phases + (size_t)q * N;
        for (int t = 0; t < N; t++) ph[t] = ct[(size_t)k * N + t];
        for (int i = 0; i < k; i++) {
            const uint64_t* a = ct + (size_t)i * N; const uint64_t* S = ks->sk_glwe.data() + (size_t)i * N;
            for (int j = 0; j < N; j++) {
                if (!S[j]) continue;
                for (int t = 0; t < j; t++) ph[t] += a[t - j + N];
                for (int t = j; t < N; t++) ph[t] -= a[t - j];
            }
        }
    }
}

void orc_decompose(uint64_t x, int b, int l, int closest_first, int64_t* digits /* index lev-1 */) {
    uint64_t v = closest_first ? closest_representable(x, b, l) : x;
    uint64_t st = decomp_init_state(v, b, l);
    for (int lev = l; lev >= 1; lev--) digits[lev - 1] = decomp_next(st, b);
}

void orc_keyswitch_batch(void* h, const uint64_t* in, int n, uint64_t* out) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1, S = ks->p.n + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; i++) keyswitch(*ks, in + (size_t)i * L, out + (size_t)i * S);
}
void orc_pbs_batch(void* h, const uint64_t* in_small, int n, uint64_t* out_big) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1, S = ks->p.n + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; i++) pbs_shift_boolean(*ks, in_small + (size_t)i * S, out_big + (size_t)i * L);
}
void orc_extract_bits_batch(void* h, const uint64_t* in_big, int n, int delta_log, int n_bits, uint64_t* out_small) {
    KeySet* ks = (KeySet*)h;
    const size_t L = (size_t)ks->p.big() + 1, S = (size_t)ks->p.n + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; i++) extract_bits(*ks, in_big + (size_t)i * L, delta_log, n_bits, out_small + (size_t)i * n_bits * S);
}
void orc_pfks_batch(void* h, const uint64_t* in_big, int n, uint64_t* out) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1; const size_t O = (size_t)(ks->p.k + 1) * (ks->p.k + 1) * ks->p.N;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; i++) pfks_all(*ks, in_big + (size_t)i * L, out + (size_t)i * O);
}
// acc += GGSW(standard domain, `levels` x base 2^b) ⊡ glwe_in
void orc_external_product(void* h, const uint64_t* ggsw_std, int levels, int b, const uint64_t* glwe_in, uint64_t* acc) {
    KeySet* ks = (KeySet*)h; const Params& p = ks->p; const int G = p.k + 1, M = p.N / 2;
    const size_t polys = (size_t)levels * G * G;
    std::vector<double> re(polys * M), im(polys * M);
    for (size_t q = 0; q < polys; q++) ks->fft->fwd_torus(ggsw_std + q * p.N, re.data() + q * M, im.data() + q * M);
    external_product_add(*ks->fft, p.k, levels, b, re.data(), im.data(), glwe_in, acc);
}
// vertical packing from standard-domain GGSWs (n_in of them, cbs_l levels each)
void orc_vertical_packing(void* h, const uint64_t* ggsw_std, int n_in, const uint64_t* lut, int n_out, uint64_t* out) {
    KeySet* ks = (KeySet*)h; const Params& p = ks->p; const int G = p.k + 1; const size_t GS = (size_t)p.cbs_l * G * G * p.N;
    std::vector<FourierGgsw> gg(n_in);
    for (int i = 0; i < n_in; i++) ggsw_to_fourier(*ks, ggsw_std + (size_t)i * GS, p.cbs_l, gg[i]);
    const size_t len = lut_len_per_output(n_in, p.N);
#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < n_out; o++) vertical_packing(*ks, lut + (size_t)o * len, len, gg, out + (size_t)o * (p.big() + 1));
}
void orc_circuit_bootstrap_batch(void* h, const uint64_t* in, int batch, int n_in, const uint64_t* lut, int n_out, uint64_t* out) {
    KeySet* ks = (KeySet*)h; const int L = ks->p.big() + 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < batch; q++) circuit_bootstrap(*ks, in + (size_t)q * n_in * L, n_in, lut, n_out, out + (size_t)q * n_out * L);
}

size_t orc_lut_len(int n_in, int N) { return lut_len_per_output(n_in, N); }
void orc_generate_lut(int n_in, int n_out, int N, const uint64_t* f_table, uint64_t* out) { generate_lut(n_in, n_out, N, f_table, out); }
uint64_t orc_encode_bit(uint64_t b) { return encode_bit(b); }
uint64_t orc_decode_bit(uint64_t e) { return decode_bit(e); }

void orc_chacha20_stream(const uint8_t key[32], uint64_t nonce, uint8_t* out, size_t n) {
    ChaCha20 c; c.init(key, nonce); c.bytes(out, n);
}
uint8_t orc_sbox(int i) { return SBOX[i & 255]; }
uint8_t orc_gf_256_mul(uint8_t a, uint8_t b) { return gf_256_mul(a, b); }
void orc_plain_key_schedule(const uint8_t key[16], uint8_t ek[176]) { plain_key_schedule(key, ek); }
void orc_plain_encrypt_block(const uint8_t ek[176], const uint8_t in[16], int rounds, uint8_t out[16]) { plain_encrypt_block(ek, in, rounds, out); }

int orc_aes_encrypt_blocks(void* h, const uint64_t* key_sched, int n_blocks, int rounds, const uint64_t* in, uint64_t* out, int in_noise_sq) {
    return aes_encrypt_blocks(*(KeySet*)h, key_sched, n_blocks, rounds, in, out, in_noise_sq);
}
int orc_aes_key_schedule(void* h, const uint64_t* key_bits, uint64_t* out) { return aes_key_schedule(*(KeySet*)h, key_bits, out); }

}  // extern "C"
