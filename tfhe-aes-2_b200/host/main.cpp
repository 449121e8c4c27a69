// tfhe_aes_cli — the reference's binary (src/bin/main.rs) on the B200 path:
//   tfhe_aes_cli --key 76b8e0ada0f13d90405d6ae55386bd28 --iv bdd219b8a08ded1a --number-of-outputs 10 [--generic] [--seed S]
// (--seed is for tests: without it the keys come from OS entropy, like the reference's)
// --generic runs the reference's generic per-byte code path (fhe_sbox_gal_mul_pbs over ByteT) instead of the fused device path.
#include "tfhe_aes.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>

using namespace aes_128;
using namespace aes_128::fhe;
using Enc = fhe_impls::cuda_woppbs_1bit::CudaWoppbs1BitSboxGalMulPbsAesEncrypt;

static std::vector<uint8_t> unhex(const std::string& s) {
    std::vector<uint8_t> v;
    for (size_t i = 0; i + 1 < s.size(); i += 2) v.push_back((uint8_t)std::stoul(s.substr(i, 2), nullptr, 16));
    return v;
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    std::map<std::string, std::string> a;
    bool generic = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--generic")) generic = true;
        else if (i + 1 < argc) { a[argv[i]] = argv[i + 1]; i++; }
    }
    if (!a.count("--key") || !a.count("--iv") || !a.count("--number-of-outputs")) {
        fprintf(stderr, "usage: %s --key <32 hex> --iv <16 hex> --number-of-outputs N [--generic] [--seed S]\n", argv[0]);
        return 2;
    }
    const auto kv = unhex(a["--key"]), iv = unhex(a["--iv"]);
    if (kv.size() != 16) { fprintf(stderr, "invalid key length, must be 16 bytes\n"); return 2; }
    if (iv.size() != 8) { fprintf(stderr, "invalid iv length, must be 8 bytes\n"); return 2; }
    const size_t n_out = std::stoul(a["--number-of-outputs"]);
    // keys come from OS entropy; --seed S (TEST ONLY) makes them reproducible and therefore publicly computable
    const bool have_seed = a.count("--seed") != 0;
    const uint64_t seed = have_seed ? std::stoull(a["--seed"]) : 0;
    printf("using implementation: CudaWoppbs1bit (%s path)\n", generic ? "generic per-byte" : "fused device");
    try {
        auto keys = tfhe::cuda_woppbs_1bit::generate_keys(64, have_seed ? &seed : nullptr);                           // generate_keys_sqrd_lvl_64 (main.rs:82-83)
        auto& client_key = keys.first; auto& ctx = keys.second;
        Key key_clear; std::copy(kv.begin(), kv.end(), key_clear.begin());
        // client side: FHE encrypt AES key and blocks (main.rs:107-116)
        auto key = fhe_encryption::encrypt_byte_array(client_key, ctx, key_clear);
        std::vector<Block> blocks_clear(n_out);
        std::vector<data_model::BlockT<Enc::Bit>> blocks;
        for (size_t ctr = 1; ctr <= n_out; ctr++) {
            Block& b = blocks_clear[ctr - 1];
            std::copy(iv.begin(), iv.end(), b.begin());
            for (int i = 0; i < 8; i++) b[8 + i] = (uint8_t)((uint64_t)ctr >> (8 * (7 - i)));
            blocks.push_back(fhe_encryption::encrypt_byte_array(client_key, ctx, b));
        }
        double t0 = now();
        auto key_schedule = Enc::key_schedule(ctx, key);                                        // main.rs:130-139
        printf("AES key expansion took: %.3fs\n", now() - t0);
        t0 = now();
        std::vector<data_model::BlockT<Enc::Bit>> enc;
        if (generic) for (auto& b : blocks) enc.push_back(Enc::encrypt_block(ctx, key_schedule, b));
        else enc = Enc::encrypt_blocks_fused(ctx, key_schedule, blocks, ROUNDS);
        printf("AES of #%zu outputs computed in: %.3fs\n", enc.size(), now() - t0);
        // client side: decrypt, compare with clear AES (main.rs:123-127)
        const auto ek = plain::key_schedule(key_clear);
        for (size_t i = 0; i < n_out; i++) {
            const Block got = fhe_encryption::decrypt_byte_array(client_key, enc[i]);
            const Block want = plain::encrypt_block(ek, blocks_clear[i], ROUNDS);
            for (uint8_t v : got) printf("%02x", v);
            printf("\n");
            if (got != want) { fprintf(stderr, "assertion failed: FHE result != clear AES for block %zu\n", i); return 1; }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "panicked: %s\n", e.what());
        return 101;
    }
    return 0;
}
