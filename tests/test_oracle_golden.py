"""Pins the oracle against every golden vector the reference's own tests hold for this path
and the product's pure-integer host helpers against the same vectors."""
import numpy as np
import pytest

E = lambda bits: [int(b) << 63 for b in bits]

CLI_KEY = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28")
CLI_IV = bytes.fromhex("bdd219b8a08ded1a")
CLI_OUT = [  # block = iv ‖ BE64(ctr), ctr = 1..10  (reference src/bin/main.rs:108-115)
    "175eea04466a73f066b5bc8195bf9e08", "ad0f0343f9591e31b9eb909c10525f0c", "057b7a00658d8b7e6ad77f1500846a7f",
    "91e7364153acab3fc062d308f102fb78", "441f9651d0c5a54b1bebe2ae64734265", "9db8ceeca32caec01905be6d3e944940",
    "238ff796a820ae20da0c98d2e4b04ee6", "5e88dabff3a50875a9d25729401bc536", "8f20bfd485fc13d0d221a2f5d99de35f",
    "a04854fb35e33887692abe7434f747a9"]


def test_encode_decode(ol, tac):
    # reference shortint_woppbs_1bit.rs:447-461
    for enc, dec in ((ol.lib().orc_encode_bit, ol.lib().orc_decode_bit), (tac.encode_bit, tac.decode_bit)):
        assert enc(0) == 0 and enc(1) == 1 << 63
        assert dec(0) == 0 and dec(1) == 0 and dec(2**64 - 1) == 0
        assert dec(1 << 63) == 1 and dec((1 << 63) - 1) == 1 and dec((1 << 63) + 1) == 1


def test_lut_layout_vertical_packing(ol, tac):
    # reference shortint_woppbs_1bit.rs:665-677
    for gen in (lambda *a: ol.generate_lut(*a), lambda i, o, n, f: tac.generate_multivariate_luts(i, o, n, f)):
        lut = gen(3, 2, 16, lambda v: v)
        assert lut.size == 16 * 2
        assert lut[0].tolist() == E([0, 0, 1, 1, 0, 0, 1, 1] + [0] * 8)
        assert lut[1].tolist() == E([0, 1, 0, 1, 0, 1, 0, 1] + [0] * 8)


def test_lut_layout_multipolynomial(ol, tac):
    # reference shortint_woppbs_1bit.rs:679-697
    for gen in (lambda *a: ol.generate_lut(*a), lambda i, o, n, f: tac.generate_multivariate_luts(i, o, n, f)):
        lut = gen(5, 2, 8, lambda v: v)
        assert lut.size == 8 * 4 * 2
        assert lut[0].tolist() == E([0, 0, 1, 1] * 8)
        assert lut[1].tolist() == E([0, 1] * 16)


def test_bit_order(ol, tac):
    # reference src/util.rs:91-95
    assert ol.u8_to_bits(0b01100011) == [0, 1, 1, 0, 0, 0, 1, 1]
    assert tac.u8_to_bits(0b01100011) == [0, 1, 1, 0, 0, 0, 1, 1]
    assert tac.bits_to_u8([0, 1, 1, 0, 0, 0, 1, 1]) == 0b01100011
    assert tac.u16_to_bits(0b1111000101100011) == [1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1]


def test_chacha20_seed0_stream(ol):
    # ChaCha20Rng::from_seed([0;32]) as drawn by reference test_helper.rs:29-36 (key | block1 | block2)
    s = ol.chacha20_stream(48)
    assert s.hex() == ("76b8e0ada0f13d90405d6ae55386bd28" "bdd219b8a08ded1aa836efcc8b770dc7" "da41597c5157488d7724e03fb8d84a37")


def test_plain_aes_vectors(ol):
    s = ol.chacha20_stream(48)
    key, b1, b2 = s[:16], s[16:32], s[32:48]
    # reference plain.rs:157-172 / test_helper.rs:47-50 (vs the `aes` crate)
    assert ol.plain_encrypt_block(key, b1).hex() == "c3763382db9e0b88b00b6d133fbd537a"
    assert ol.plain_encrypt_block(key, b2).hex() == "3ff5a50205db74d007cdb91899a0d7ee"
    # reduced rounds, final round always uses rk10 (plain.rs:75-103; test_light uses rounds = 2)
    assert ol.plain_encrypt_block(key, b1, 2).hex() == "5c864f984df12113a07c22a99f49f0a1"
    assert ol.plain_encrypt_block(key, b1, 1).hex() == "de3011192c24fd50c3b199187f869fa4"
    ek = ol.plain_key_schedule(key)
    assert bytes(ek[16:32]).hex() == "33c2d4409333e9d0d36e833580e83e1d"
    assert bytes(ek[160:176]).hex() == "c12086c64f5b1a09581000661e84ef01"
    # FIPS-197 C.1 (test_helper.rs:61-83)
    assert ol.plain_encrypt_block(bytes.fromhex("000102030405060708090a0b0c0d0e0f"),
                                  bytes.fromhex("00112233445566778899aabbccddeeff")).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"


def test_plain_aes_matches_cryptography(ol):
    crypto = pytest.importorskip("cryptography.hazmat.primitives.ciphers")
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    rng = np.random.default_rng(5)
    for _ in range(20):
        key, blk = rng.bytes(16), rng.bytes(16)
        enc = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
        assert ol.plain_encrypt_block(key, blk) == enc.update(blk)


def test_cli_counter_blocks(ol):
    for ctr, want in enumerate(CLI_OUT, start=1):
        assert ol.plain_encrypt_block(CLI_KEY, CLI_IV + ctr.to_bytes(8, "big")).hex() == want


def test_sbox_gal_mul_lut_spot_values(ol):
    from conftest import sbox_gal_mul_fn
    f = sbox_gal_mul_fn(ol)
    # SURVEY.md Appendix B (fhe_impls/shortint_woppbs_1bit.rs:98-111)
    assert [f"{f(b):06x}" for b in (0x00, 0x01, 0x53, 0xFF)] == ["63c6a5", "7cf884", "edc12c", "162c3a"]
    lut = ol.generate_lut(8, 24, 512, f)
    assert lut.shape == (24, 512)
    for b in (0x00, 0x53, 0xFF):
        word = sum(int(lut[o, b] >> 63) << (23 - o) for o in range(24))
        assert word == f(b)
    assert not lut[:, 256:].any()


def test_decomposer_recomposes(ol):
    rng = np.random.default_rng(9)
    for b, l in ((12, 3), (3, 4), (16, 2), (13, 1), (15, 2), (9, 4), (24, 1), (2, 6)):
        for x in [0, 1, 2**63, 2**64 - 1, 2**63 - 1] + [int(v) for v in rng.integers(0, 2**64, 300, dtype=np.uint64)]:
            d = np.zeros(l, dtype=np.int64)
            ol.lib().orc_decompose(x, b, l, 0, d)
            assert all(abs(int(v)) <= 1 << (b - 1) for v in d)
            rec = sum(int(d[i]) << (64 - b * (i + 1)) for i in range(l)) % 2**64
            nonrep = 64 - b * l
            want = (((x >> nonrep) + ((x >> (nonrep - 1)) & 1)) << nonrep) % 2**64
            assert rec == want


# ------------------------------------------------------------------ decrypt-level model tests of the reference, on the oracle
def test_oracle_encrypt_decrypt_and_trivial(oracle64):
    # reference shortint_woppbs_1bit.rs:463-482
    cts = oracle64.encrypt_bits([0, 1])
    assert oracle64.decrypt_bits(cts).tolist() == [0, 1]
    triv = np.zeros((2, oracle64.big1), dtype=np.uint64)
    triv[1, -1] = 1 << 63
    assert oracle64.decrypt_bits(triv).tolist() == [0, 1]


def test_oracle_xor_truth_table(oracle64):
    # reference :484-503 (leveled XOR = ciphertext add)
    b = oracle64.encrypt_bits([0, 1, 0, 1])
    assert oracle64.decrypt_bits((b[0] + b[1])[None])[0] == 1
    assert oracle64.decrypt_bits((b[0] + b[2])[None])[0] == 0
    assert oracle64.decrypt_bits((b[1] + b[3])[None])[0] == 0


def test_oracle_multivariate_parity_3(oracle64, ol):
    # reference :531-539
    parity = lambda v: bin(v).count("1") % 2
    lut = oracle64.generate_lookup_table(3, 1, parity)
    for word in (0b001, 0b000, 0b100, 0b101):
        bits = ol.u8_to_bits(word)[5:]
        out = oracle64.circuit_bootstrap(oracle64.encrypt_bits(bits), lut, 1)
        assert oracle64.decrypt_bits(out)[0] == parity(word)


def test_oracle_multivalued_square_3(oracle64, ol):
    # reference :574-582
    sq = lambda v: (v * v) % 8
    lut = oracle64.generate_lookup_table(3, 3, sq)
    for word in (0b101, 0b000, 0b100):
        bits = ol.u8_to_bits(word)[5:]
        out = oracle64.circuit_bootstrap(oracle64.encrypt_bits(bits), lut, 3)
        got = sum(int(b) << (2 - i) for i, b in enumerate(oracle64.decrypt_bits(out)))
        assert got == sq(word)


def test_oracle_sbox_gal_mul(oracle64, ol):
    # BASELINE config 2 on the CPU: one 8→24 SBOX·{1,2,3}
    from conftest import sbox_gal_mul_fn
    f = sbox_gal_mul_fn(ol)
    lut = oracle64.generate_lookup_table(8, 24, f)
    out = oracle64.circuit_bootstrap(oracle64.encrypt_bytes([0x53])[0], lut, 24)
    assert oracle64.decrypt_bytes(out).hex() == "edc12c"


def test_oracle_adder_with_trivial_carry(oracle64, ol):
    # reference :792-836 (2→2 LUT, trivial carry-in, exercises a noiseless input to circuit_bootstrap)
    add = lambda v: ((v >> 1) & 1) + (v & 1)
    lut = oracle64.generate_lookup_table(2, 2, add)
    carry = np.zeros(oracle64.big1, dtype=np.uint64)
    carry[-1] = 1 << 63
    bit = oracle64.encrypt_bits([1])[0]
    out = oracle64.circuit_bootstrap(np.stack([carry, bit]), lut, 2)
    assert oracle64.decrypt_bits(out).tolist() == [1, 0]


def test_oracle_increment_8bit_adder_9_to_9(oracle64, ol):
    # reference test_increment_8bit_adder (:838-877), shortened to the two low bytes (the GPU test runs all 16): the 9→9
    # LUT has n_in = log2 N — nine blind-rotation steps, the first by X^-256 — and a trivial carry-in
    add = lambda v: (v & 0xFF) + ((v >> 8) & 1)
    lut = oracle64.generate_lookup_table(9, 9, add)
    assert lut.shape == (9, 512)
    value = oracle64.encrypt_bytes([0x00, 0xFF])
    for _ in range(2):
        carry = np.zeros(oracle64.big1, dtype=np.uint64)
        carry[-1] = 1 << 63
        new = []
        for byte in value[::-1]:
            out = oracle64.circuit_bootstrap(np.concatenate([carry[None], byte]), lut, 9)
            carry, nb = out[0], out[1:]
            new.append(nb)
        value = np.stack(new[::-1])
    assert oracle64.decrypt_bytes(value.reshape(-1, oracle64.big1)) == bytes([0x01, 0x01])


def test_oracle_extract_bits_chain(oracle64):
    # [U] wop_pbs.rs::extract_bits beyond the one-bit case: 3-bit messages at delta_log 61 come back bit for bit (most significant
    # first), and (63, 1) — what extract_dual_bit_from_bit asks for (shortint_woppbs_1bit.rs:342-349) — is the keyswitch alone
    msgs = [0, 3, 5, 6]
    cts = oracle64.encrypt_bits([0] * len(msgs), first_index=900)
    cts[:, -1] += np.array(msgs, dtype=np.uint64) << np.uint64(61)
    out = oracle64.extract_bits(cts, 61, 3)
    ph = oracle64.phases_small(out.reshape(-1, out.shape[-1])).reshape(len(msgs), 3)
    bits = ((ph + np.uint64(1 << 62)) >> np.uint64(63)).astype(int)
    assert [int("".join(map(str, b)), 2) for b in bits.tolist()] == msgs
    assert np.array_equal(oracle64.extract_bits(cts[:2], 63, 1)[:, 0], oracle64.keyswitch(cts[:2]))


def test_oracle_two_level_circuit_bootstrap(oracle64, ol):
    # [U] wop_pbs.rs::circuit_bootstrap_boolean with cbs_level = 2 (the reference's 8-bit model uses 4 levels,
    # shortint_woppbs_8bit.rs; the 1-bit sets use 1): one bootstrap + (k+1) PFKS per level, vertical packing with 2-level GGSWs.
    # Evaluation keys do not depend on the circuit-bootstrap decomposition, so the lvl_64 keys are reused.
    p = ol.preset(64)
    p.cbs_l, p.cbs_b = 2, 8
    o = ol.Oracle(p, seed=0, raw=(oracle64.sk_glwe, oracle64.sk_lwe, oracle64.bsk, oracle64.ksk, oracle64.pfpksk))
    lut = o.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    out = o.circuit_bootstrap(oracle64.encrypt_bytes([0x53])[0], lut, 8)
    assert oracle64.decrypt_bytes(out) == bytes([ol.sbox(0x53)])


def test_oracle_cmux_tree_16_to_8(ol):
    # reference :626-659 — 16 inputs at N = 1024 (params_sqrd_lvl_1) exercises the real CMux tree (6 tree bits)
    o = ol.Oracle(1, seed=77)
    b1, b2 = 0b11000110, 0b10101010
    xor_fn = lambda v: (v >> 8) ^ (v & 0xFF)
    lut = o.generate_lookup_table(16, 8, xor_fn)
    assert lut.shape == (8, 1024 << 6)
    cts = o.encrypt_bytes([b1, b2]).reshape(16, -1)
    out = o.circuit_bootstrap(cts, lut, 8)
    assert o.decrypt_bytes(out) == bytes([b1 ^ b2])
