// host_tests — the reference's model / AES tests that need keys, through the C++ host mirror (run on the GPU box by
// tests/test_gpu_host_cpp.py).  Mirrors src/tfhe/shortint_woppbs_1bit.rs:463-529 and fhe_impls/shortint_woppbs_1bit.rs:185-193.
#include "tfhe_aes.hpp"

#include <cstdio>
#include <map>

using namespace aes_128;
using namespace aes_128::fhe;
using namespace tfhe::cuda_woppbs_1bit;
using Enc = fhe_impls::cuda_woppbs_1bit::CudaWoppbs1BitSboxGalMulPbsAesEncrypt;

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)
template <class F> static bool panics_with(F f, const char* what) {
    try { f(); } catch (const Panic& e) { return std::string(e.what()).find(what) != std::string::npos; }
    return false;
}

int main() {
    const uint64_t seed = 424242;
    auto keys = generate_keys(64, &seed);
    auto& ck = keys.first; auto& ctx = keys.second;
    // test_bit_encrypt_decrypt / test_bit_xor
    BitCt b1 = ck.encrypt(0, ctx), b2 = ck.encrypt(1, ctx), b3 = ck.encrypt(0, ctx), b4 = ck.encrypt(1, ctx);
    CHECK(ck.decrypt(b1) == 0 && ck.decrypt(b2) == 1);
    CHECK(ck.decrypt(BitCt::trivial(1, ctx)) == 1 && ck.decrypt(BitCt::trivial(0, ctx)) == 0);
    CHECK(ck.decrypt(b1 ^ b2) == 1 && ck.decrypt(b1 ^ b3) == 0 && ck.decrypt(b2 ^ b4) == 0);
    BitCt t0 = BitCt::trivial(0, ctx);
    CHECK(ck.decrypt(b2 ^ t0) == 1);
    (void)(t0 ^ t0 ^ t0);                                                    // trivial does not accumulate noise
    CHECK(panics_with([&] { (void)(b1 ^ b1); }, "noise components not independent"));
    CHECK(panics_with([&] { (void)ck.encrypt(2, ctx); }, "cleartext out of bounds"));
    {   // NoiseTooBig: 65 independent fresh ciphertexts exceed max_noise_level_squared = 64
        BitCt acc = ck.encrypt(0, ctx);
        CHECK(panics_with([&] { for (int i = 0; i < 64; i++) acc ^= ck.encrypt(0, ctx); }, "NoiseTooBig"));
    }
    // multivariate parity (3 bits) and multivalued square (3 → 3), :531-582
    auto parity = ctx.generate_lookup_table(3, 1, [](uint16_t v) { return (uint64_t)(__builtin_popcount(v) & 1); });
    auto square = ctx.generate_lookup_table(3, 3, [](uint16_t v) { return (uint64_t)((v * v) % 8); });
    for (uint16_t word : {0b001, 0b000, 0b100, 0b101}) {
        BitCt x0 = ck.encrypt((word >> 2) & 1, ctx), x1 = ck.encrypt((word >> 1) & 1, ctx), x2 = ck.encrypt(word & 1, ctx);
        auto d = circuit_bootstrap(ctx, {&x0, &x1, &x2}, parity);
        CHECK(ck.decrypt(d[0]) == (uint64_t)(__builtin_popcount(word) & 1));
        CHECK(d[0].noise_level.noise_level_squared == 3);
        auto s = circuit_bootstrap(ctx, {&x0, &x1, &x2}, square);
        CHECK(((ck.decrypt(s[0]) << 2) | (ck.decrypt(s[1]) << 1) | ck.decrypt(s[2])) == (uint64_t)((word * word) % 8));
    }
    // test_light_gal_mul: 2 rounds, clear key schedule encrypted directly, block from the ChaCha20 seed-0 stream (test_helper.rs:86-120),
    // through the reference's GENERIC code path (ByteT policy) and through the fused device path
    const Key key_clear = {0x76, 0xb8, 0xe0, 0xad, 0xa0, 0xf1, 0x3d, 0x90, 0x40, 0x5d, 0x6a, 0xe5, 0x53, 0x86, 0xbd, 0x28};
    const Block block_clear = {0xbd, 0xd2, 0x19, 0xb8, 0xa0, 0x8d, 0xed, 0x1a, 0xa8, 0x36, 0xef, 0xcc, 0x8b, 0x77, 0x0d, 0xc7};
    const auto ek_clear = plain::key_schedule(key_clear);
    std::array<data_model::Word<BitCt>, 44> ek;
    for (int w = 0; w < 44; w++) for (int b = 0; b < 4; b++) ek[w][b] = fhe_encryption::encrypt_byte(ck, ctx, ek_clear[4 * w + b]);
    auto block = fhe_encryption::encrypt_byte_array(ck, ctx, block_clear);
    const Block want = plain::encrypt_block(ek_clear, block_clear, 2);
    const Block expect = {0x5c, 0x86, 0x4f, 0x98, 0x4d, 0xf1, 0x21, 0x13, 0xa0, 0x7c, 0x22, 0xa9, 0x9f, 0x49, 0xf0, 0xa1};
    CHECK(want == expect);
    auto enc_generic = Enc::encrypt_block_for_rounds(ctx, ek, block, 2);
    CHECK(fhe_encryption::decrypt_byte_array(ck, enc_generic) == expect);
    CHECK(enc_generic[0][0].noise_level.noise_level_squared == 9);           // SBOX output (8) + round key (1)
    auto enc_fused = Enc::encrypt_blocks_fused(ctx, ek, {block}, 2);
    CHECK(fhe_encryption::decrypt_byte_array(ck, enc_fused[0]) == expect);
    printf("host_tests: all passed\n");
    return 0;
}
