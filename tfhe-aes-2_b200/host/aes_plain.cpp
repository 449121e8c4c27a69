// Clear AES-128 with a `rounds` argument — reference src/aes_128/plain.rs:75-136 and the constants of src/aes_128.rs:18-39.
#include "tfhe_aes.hpp"

#include <map>

namespace aes_128 {

// The S-box is generated from its definition — multiplicative inverse in GF(2^8) followed by the affine map — rather than
// tabulated; it equals the table of the reference (src/aes_128.rs:18-35), which the FIPS-197 and CLI vectors of the tests pin.
namespace {
struct SboxTable {
    uint8_t v[256];
    SboxTable() {
        uint8_t p = 1, q = 1;
        do {
            p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));                                  // p ← 3·p
            q ^= (uint8_t)(q << 1); q ^= (uint8_t)(q << 2); q ^= (uint8_t)(q << 4);                   // q ← q / 3
            if (q & 0x80) q ^= 0x09;
            const auto rotl = [&](int n) { return (uint8_t)((q << n) | (q >> (8 - n))); };
            v[p] = (uint8_t)(q ^ rotl(1) ^ rotl(2) ^ rotl(3) ^ rotl(4) ^ 0x63);
        } while (p != 1);
        v[0] = 0x63;
    }
};
const SboxTable kSbox;
}  // namespace
const uint8_t* const SBOX = kSbox.v;

namespace plain {
std::array<uint8_t, 176> key_schedule(const Key& key) {
    std::array<uint8_t, 176> ek{};
    std::copy(key.begin(), key.end(), ek.begin());
    for (int i = 4; i < 44; i++) {
        uint8_t w[4] = {ek[4 * (i - 1)], ek[4 * (i - 1) + 1], ek[4 * (i - 1) + 2], ek[4 * (i - 1) + 3]};
        if (i % 4 == 0) {
            const uint8_t t = w[0];
            w[0] = SBOX[w[1]] ^ RC[i / 4]; w[1] = SBOX[w[2]]; w[2] = SBOX[w[3]]; w[3] = SBOX[t];
        }
        for (int b = 0; b < 4; b++) ek[4 * i + b] = ek[4 * (i - 4) + b] ^ w[b];
    }
    return ek;
}
Block encrypt_block(const std::array<uint8_t, 176>& ek, const Block& block, int rounds) {
    Block s;
    for (int i = 0; i < 16; i++) s[i] = block[i] ^ ek[i];
    auto sub_shift = [&]() {
        Block t;
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) t[4 * c + r] = SBOX[s[4 * ((c + r) % 4) + r]];
        s = t;
    };
    for (int rd = 1; rd < rounds; rd++) {
        sub_shift();
        Block t;
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++)
            t[4 * c + r] = gf_256_mul(s[4 * c + r], 2) ^ s[4 * c + (r + 3) % 4] ^ s[4 * c + (r + 2) % 4] ^ gf_256_mul(s[4 * c + (r + 1) % 4], 3);
        for (int i = 0; i < 16; i++) s[i] = t[i] ^ ek[16 * rd + i];
    }
    sub_shift();
    for (int i = 0; i < 16; i++) s[i] ^= ek[160 + i];
    return s;
}
}  // namespace plain
}  // namespace aes_128
