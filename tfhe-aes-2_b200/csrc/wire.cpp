// wire.cpp — flat key / ciphertext files (host only).
//
// The reference pulls raw key material out of tfhe-rs (`into_raw_parts`, src/tfhe/shortint_woppbs_1bit.rs:245-268) but
// never stores it; a drop-in replacement needs a way to take the SAME keys a tfhe-rs process generated, and a way to hand
// ciphertexts back and forth, so that intermediate values can be compared against a real reference build.  The format is
// deliberately trivial: every payload is the `u64` container of the corresponding tfhe 0.11.2 `core_crypto` object,
// `.as_ref()` order, little-endian (layouts in include/tfhe_aes_cuda.h; INTEGRATION.md shows the 30 lines of Rust that
// write the same file).
//
//   file    := header  section*
//   header  := magic[8] = "TACWIRE\x01" | u32 version = 1 | u32 n_sections | tac_params (12 × i32, 3 × f64; 72 bytes) | u64 reserved = 0
//   section := u32 id | u32 reserved = 0 | u64 n_words | u64 dim0 | u64 dim1 | u64 fnv1a64(payload) | payload[n_words × u64]
//   ids       1 GLWE secret key (kN words 0/1)      2 LWE secret key (n words 0/1)
//             3 bootstrap key, STANDARD domain       LweBootstrapKeyOwned<u64>                                [n][l][k+1][k+1][N]
//             4 keyswitch key big → small            LweKeyswitchKeyOwned<u64>                                [kN][l][n+1]
//             5 circuit-bootstrap PFPKSK list        LwePrivateFunctionalPackingKeyswitchKeyListOwned<u64>    [k+1][kN+1][l][(k+1)N]
//             6 LWE ciphertext list                  LweCiphertextListOwned<u64>, dim0 = count, dim1 = lwe_size
// Secret-key sections exist for tests and for moving a client between processes; a server-side file carries 3-5 only.
#include "../../include/tfhe_aes_cuda.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

const char kMagic[8] = {'T', 'A', 'C', 'W', 'I', 'R', 'E', 1};

struct SectionHeader {
    uint32_t id, reserved;
    uint64_t n_words, dim0, dim1, checksum;
};
static_assert(sizeof(SectionHeader) == 40, "section header layout");
static_assert(sizeof(tac_params) == 72, "tac_params layout");

uint64_t fnv1a64(const uint64_t* w, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(w);
    for (size_t i = 0; i < n * 8; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

struct File {
    FILE* f = nullptr;
    explicit File(const char* path, const char* mode) : f(fopen(path, mode)) {}
    ~File() { if (f) fclose(f); }
};

bool write_header(FILE* f, const tac_params* p, uint32_t n_sections) {
    const uint32_t version = 1;
    const uint64_t reserved = 0;
    tac_params zero;
    memset(&zero, 0, sizeof zero);
    return fwrite(kMagic, 1, 8, f) == 8 && fwrite(&version, 4, 1, f) == 1 && fwrite(&n_sections, 4, 1, f) == 1 &&
           fwrite(p ? p : &zero, sizeof(tac_params), 1, f) == 1 && fwrite(&reserved, 8, 1, f) == 1;
}
bool write_section(FILE* f, uint32_t id, const uint64_t* words, uint64_t n_words, uint64_t dim0, uint64_t dim1) {
    SectionHeader h{id, 0, n_words, dim0, dim1, fnv1a64(words, n_words)};
    return fwrite(&h, sizeof h, 1, f) == 1 && fwrite(words, 8, n_words, f) == n_words;
}
// positions the stream at the payload of section `id`; false if absent / malformed
bool find_section(FILE* f, uint32_t id, tac_params* p, SectionHeader* out, uint32_t* present) {
    char magic[8];
    uint32_t version, n_sections;
    uint64_t reserved;
    tac_params pp;
    if (fseek(f, 0, SEEK_SET) != 0 || fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0) return false;
    if (fread(&version, 4, 1, f) != 1 || version != 1 || fread(&n_sections, 4, 1, f) != 1) return false;
    if (fread(&pp, sizeof pp, 1, f) != 1 || fread(&reserved, 8, 1, f) != 1) return false;
    if (p) *p = pp;
    if (present) *present = 0;
    bool found = false;
    long found_pos = 0;
    for (uint32_t s = 0; s < n_sections; s++) {
        SectionHeader h;
        if (fread(&h, sizeof h, 1, f) != 1) return false;
        if (present && h.id < 32) *present |= 1u << h.id;
        if (h.id == id && !found) { found = true; found_pos = ftell(f); if (out) *out = h; }
        if (fseek(f, (long)(h.n_words * 8), SEEK_CUR) != 0) return false;
    }
    if (id == 0) return true;                     // header / presence query only
    if (!found) return false;
    return fseek(f, found_pos, SEEK_SET) == 0;
}
int read_section(const char* path, uint32_t id, uint64_t* out, size_t expect_words, const tac_params* expect_params) {
    File fl(path, "rb");
    if (!fl.f) return TAC_ERR_ARG;
    tac_params p;
    SectionHeader h;
    if (!find_section(fl.f, id, &p, &h, nullptr)) return TAC_ERR_STATE;
    if (expect_params && memcmp(&p, expect_params, sizeof p) != 0) return TAC_ERR_ARG;       // keys of another parameter set
    if (h.n_words != expect_words) return TAC_ERR_ARG;
    if (fread(out, 8, h.n_words, fl.f) != h.n_words) return TAC_ERR_STATE;
    if (fnv1a64(out, h.n_words) != h.checksum) return TAC_ERR_STATE;                          // corrupted payload
    return TAC_OK;
}

}  // namespace

extern "C" {

int tac_keys_save(const char* path, const tac_params* p, const uint64_t* sk_glwe, const uint64_t* sk_lwe, const uint64_t* bsk_std, const uint64_t* ksk,
                  const uint64_t* pfpksk) {
    if (!path || !p) return TAC_ERR_ARG;
    File fl(path, "wb");
    if (!fl.f) return TAC_ERR_ARG;
    const uint64_t* ptr[5] = {sk_glwe, sk_lwe, bsk_std, ksk, pfpksk};
    uint32_t n = 0;
    for (auto* q : ptr) n += q != nullptr;
    if (!write_header(fl.f, p, n)) return TAC_ERR_STATE;
    const uint64_t kN = (uint64_t)p->glwe_dimension * p->polynomial_size;
    const uint64_t d0[5] = {kN, (uint64_t)p->lwe_dimension, (uint64_t)p->lwe_dimension, kN, (uint64_t)p->glwe_dimension + 1};
    const uint64_t d1[5] = {1, 1, (uint64_t)p->pbs_level, (uint64_t)p->ks_level, kN + 1};
    for (int which = 0; which < 5; which++)
        if (ptr[which] && !write_section(fl.f, (uint32_t)which + 1, ptr[which], tac_key_len(p, which), d0[which], d1[which])) return TAC_ERR_STATE;
    return fflush(fl.f) == 0 ? TAC_OK : TAC_ERR_STATE;
}
int tac_keys_load_params(const char* path, tac_params* p, uint32_t* present_mask) {
    if (!path || !p) return TAC_ERR_ARG;
    File fl(path, "rb");
    if (!fl.f) return TAC_ERR_ARG;
    return find_section(fl.f, 0, p, nullptr, present_mask) ? TAC_OK : TAC_ERR_STATE;
}
int tac_keys_load(const char* path, const tac_params* p, uint64_t* sk_glwe, uint64_t* sk_lwe, uint64_t* bsk_std, uint64_t* ksk, uint64_t* pfpksk) {
    if (!path || !p) return TAC_ERR_ARG;
    uint64_t* ptr[5] = {sk_glwe, sk_lwe, bsk_std, ksk, pfpksk};
    for (int which = 0; which < 5; which++)
        if (ptr[which]) {
            const int rc = read_section(path, (uint32_t)which + 1, ptr[which], tac_key_len(p, which), p);
            if (rc) return rc;
        }
    return TAC_OK;
}
int tac_lwe_list_save(const char* path, uint64_t lwe_size, uint64_t count, const uint64_t* words) {
    if (!path || (!words && count)) return TAC_ERR_ARG;
    File fl(path, "wb");
    if (!fl.f) return TAC_ERR_ARG;
    if (!write_header(fl.f, nullptr, 1) || !write_section(fl.f, 6, words, lwe_size * count, count, lwe_size)) return TAC_ERR_STATE;
    return fflush(fl.f) == 0 ? TAC_OK : TAC_ERR_STATE;
}
int tac_lwe_list_load(const char* path, uint64_t* lwe_size, uint64_t* count, uint64_t* words, size_t capacity_words) {
    if (!path) return TAC_ERR_ARG;
    File fl(path, "rb");
    if (!fl.f) return TAC_ERR_ARG;
    SectionHeader h;
    if (!find_section(fl.f, 6, nullptr, &h, nullptr)) return TAC_ERR_STATE;
    if (h.dim0 * h.dim1 != h.n_words) return TAC_ERR_STATE;
    if (count) *count = h.dim0;
    if (lwe_size) *lwe_size = h.dim1;
    if (!words) return TAC_OK;                     // size query
    if (capacity_words < h.n_words) return TAC_ERR_ARG;
    if (fread(words, 8, h.n_words, fl.f) != h.n_words || fnv1a64(words, h.n_words) != h.checksum) return TAC_ERR_STATE;
    return TAC_OK;
}

}  // extern "C"
