"""GPU parity tests (run on the B200 box with -m gpu): every stage of the hot path through the C ABI against the oracle
on identical keys and identical inputs.

Bars:
  * integer stages (keyswitch, PFKS, sample extraction, leveled adds, LUT indexing) — bit-exact;
  * one external-product step (f64 FFT) — |Δ| < 2^35 of 2^64 per coefficient (2^-29 relative; different FFT operation order);
  * multi-step f64 stages (PBS = 677 steps, vertical packing = 8 steps) — raw words of two correct runs diverge as soon as
    one low digit differs, so they are compared on the decrypted phase: same message, and phase error within the noise
    the oracle itself shows;
  * everything end to end — decrypts to clear AES / the clear function.
"""
import os

import numpy as np
import pytest

from conftest import SEED, sbox_gal_mul_fn

pytestmark = pytest.mark.gpu


def signed(x):
    return np.asarray(x, dtype=np.uint64).astype(np.int64).astype(np.float64)


# ---------------------------------------------------------------------------------------------- integer stages: bit-exact
@pytest.mark.parametrize("n", [1, 17, 300])
def test_keyswitch_bit_exact(gpu64, oracle64, n):
    ck, ctx = gpu64
    rng = np.random.default_rng(n)
    cts = ck.encrypt_bits(rng.integers(0, 2, n))
    got = ctx.stage_keyswitch(cts)
    assert np.array_equal(got, oracle64.keyswitch(cts))
    # and it still decrypts under the small key
    ph = oracle64.phases_small(got)
    assert ((ph + np.uint64(1 << 62)) >> np.uint64(63)).tolist() == ck.decrypt_bits(cts).tolist()


def test_keyswitch_edge_inputs(gpu64, oracle64):
    ck, ctx = gpu64
    L = ck.params.big_lwe_size
    cts = np.zeros((4, L), dtype=np.uint64)
    cts[1, :] = np.uint64(2**64 - 1)
    cts[2, :] = np.uint64(1 << 63)
    cts[3, ::2] = np.uint64((1 << 63) + (1 << 51))          # decomposition ties
    assert np.array_equal(ctx.stage_keyswitch(cts), oracle64.keyswitch(cts))


@pytest.mark.parametrize("n", [1, 20, 260])
def test_pfks_bit_exact(gpu64, oracle64, n):
    ck, ctx = gpu64
    rng = np.random.default_rng(100 + n)
    x = rng.integers(0, 2**64, (n, ck.params.big_lwe_size), dtype=np.uint64)
    x[0, :8] = [0, 2**64 - 1, 1 << 63, (1 << 63) - 1, 1 << 47, (1 << 47) - 1, (1 << 31), (1 << 63) + (1 << 31)]
    got = ctx.stage_pfks(x)
    assert np.array_equal(got, oracle64.pfks(x))


def test_pfks_tie_list_overflow_is_exact(gpu64, oracle64):
    """More +B/2 tie digits than the fix-up list holds (65536): every element 2^47 decomposes to the level-2 digit +2^15, so
    33 ciphertexts carry 33·2049 = 67617 of them.  The scan fallback must still give the oracle's words (it used to drop
    the excess silently)."""
    ck, ctx = gpu64
    x = np.full((33, ck.params.big_lwe_size), 1 << 47, dtype=np.uint64)
    x[1, ::3] = np.uint64(1 << 63)                                   # level-1 ties as well
    x[2, 5:900] = np.uint64(12345678901234567)                       # and ordinary elements in between
    assert np.array_equal(ctx.stage_pfks(x), oracle64.pfks(x))
    # and a batch just below the capacity still goes through the list path
    assert np.array_equal(ctx.stage_pfks(x[:3]), oracle64.pfks(x[:3]))


def test_sample_extract_bit_exact(gpu64):
    """[U] extract_lwe_sample_from_glwe_ciphertext(.., MonomialDegree(0)): body = b[0]; mask_i[0] = a_i[0], mask_i[j] = −a_i[N−j]"""
    ck, ctx = gpu64
    p = ck.params
    k, N = p.glwe_dimension, p.polynomial_size
    rng = np.random.default_rng(15)
    g = rng.integers(0, 2**64, (7, k + 1, N), dtype=np.uint64)
    got = ctx.stage_sample_extract(g)
    want = np.empty((7, k * N + 1), dtype=np.uint64)
    for i in range(k):
        want[:, i * N] = g[:, i, 0]
        want[:, i * N + 1:(i + 1) * N] = (np.uint64(0) - g[:, i, :0:-1])
    want[:, k * N] = g[:, k, 0]
    assert np.array_equal(got, want)
    assert ctx.stage_sample_extract(g[:0]).shape == (0, k * N + 1)


def test_poly_fft_matches_numpy(gpu64):
    """fill_with_forward_fourier alone ([U] Fft::forward_as_torus): size-N/2 complex FFT of the folded, twisted polynomial
    (torus words read as signed fractions), compared through the kernel's slot permutation; relative tolerance 2^-40"""
    ck, ctx = gpu64
    N = ck.params.polynomial_size
    M = N // 2
    rng = np.random.default_rng(16)
    polys = rng.integers(0, 2**64, (40, N), dtype=np.uint64)
    polys[0] = 0
    polys[1] = np.uint64(1 << 63)
    got = ctx.stage_poly_fft(polys)
    freq = ctx.fft_slot_frequencies()
    assert sorted(freq.tolist()) == list(range(M))
    x = polys.astype(np.int64).astype(np.float64) / 2.0**64
    z = (x[:, :M] + 1j * x[:, M:]) * np.exp(1j * np.pi * np.arange(M) / N)
    want = np.fft.fft(z, axis=1)[:, freq] / M                      # the kernels pre-scale by 1/M (the inverse is unnormalised)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 2.0**-40 * scale, np.log2(np.abs(got - want).max() / scale)


def test_extract_bits_chain(gpu64, oracle64):
    """[U] wop_pbs.rs::extract_bits beyond the one-bit case of the AES path: 3-bit messages at delta_log 61 — keyswitch,
    bootstrap of the sign bit, subtract, three times.  The first (least significant) extraction is integer-only and must equal
    the oracle's words.  The later ones follow an f64 bootstrap: two correct blind rotations agree on the phase but not on the
    mask words, so the keyswitch that follows draws an independent noise realisation (σ ≈ 2^57) — they are compared on the
    decoded bit and on the error band.  (63, 1) is the keyswitch itself."""
    ck, ctx = gpu64
    msgs = list(range(8)) + [5, 2]
    cts = ck.encrypt_bits([0] * len(msgs), first_index=4242)
    cts[:, -1] += np.array(msgs, dtype=np.uint64) << np.uint64(61)
    got = ctx.extract_bits(cts, 61, 3)
    ref = oracle64.extract_bits(cts, 61, 3)
    assert np.array_equal(got[:, 2], ref[:, 2])                              # bit 0: shift + keyswitch only
    ph_g = oracle64.phases_small(got.reshape(-1, got.shape[-1])).reshape(len(msgs), 3)
    ph_r = oracle64.phases_small(ref.reshape(-1, ref.shape[-1])).reshape(len(msgs), 3)
    bits = ((ph_g + np.uint64(1 << 62)) >> np.uint64(63)).astype(int)
    assert [int("".join(map(str, b)), 2) for b in bits.tolist()] == msgs
    want = bits.astype(np.uint64) << np.uint64(63)
    e_g, e_r = signed(ph_g - want), signed(ph_r - want)
    assert np.abs(e_g).max() < 2.0**61 and np.abs(e_r).max() < 2.0**61      # decoding margin 2^62
    assert e_g.std() < 2 * e_r.std() + 2.0**55
    one = ctx.extract_bits(cts[:3], 63, 1)
    assert np.array_equal(one[:, 0], ctx.stage_keyswitch(cts[:3]))
    with pytest.raises(RuntimeError, match="delta_log"):
        ctx.extract_bits(cts[:1], 62, 3)


def test_lwe_add_batch_bit_exact(gpu64):
    ck, ctx = gpu64
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**64, (37, ck.params.big_lwe_size), dtype=np.uint64)
    b = rng.integers(0, 2**64, (37, ck.params.big_lwe_size), dtype=np.uint64)
    assert np.array_equal(ctx.lwe_add_batch(a, b), a + b)
    assert ctx.lwe_add_batch(a[:0], b[:0]).shape == (0, ck.params.big_lwe_size)


# ---------------------------------------------------------------------------------------------- f64 core: tolerance
def _rot_diff(a, r, N):
    idx = (np.arange(N) - r) % (2 * N)
    rotd = np.where(idx < N, a[:, idx % N], (-a[:, idx % N].astype(np.int64)).astype(np.uint64))
    return rotd - a


@pytest.mark.parametrize("levels,blog", [(3, 12), (1, 13)])
def test_cmux_rotate_step_tolerance(gpu64, oracle64, levels, blog):
    ck, ctx = gpu64
    p = ck.params
    G, N = p.glwe_dimension + 1, p.polynomial_size
    rng = np.random.default_rng(levels)
    if levels == 3:
        ggsw = np.ascontiguousarray(ck.bsk.reshape(p.lwe_dimension, 3, G, G, N)[5])
    else:
        x = oracle64.pbs(oracle64.keyswitch(ck.encrypt_bits([1])))
        ggsw = oracle64.pfks(x)[0].reshape(1, G, G, N)                      # a circuit-bootstrapped GGSW of the bit 1
    n_acc = 6
    acc = rng.integers(0, 2**64, (n_acc, G, N), dtype=np.uint64)
    rot = rng.integers(0, 2 * N, n_acc).astype(np.int32)
    rot[0], rot[1] = 0, 2 * N - 1
    got = ctx.stage_cmux_rotate(ggsw, levels, blog, acc, rot).reshape(n_acc, G, N)
    for b in range(n_acc):
        want = oracle64.external_product(ggsw, levels, blog, _rot_diff(acc[b], int(rot[b]), N).reshape(-1), acc[b].reshape(-1)).reshape(G, N)
        d = np.abs(signed(got[b] - want))
        assert d.max() < 2.0**35, (b, np.log2(d.max()))
    assert np.array_equal(got[0], acc[0])                                    # rot = 0 ⇒ exact no-op


def test_pbs_phase_and_noise(gpu64, oracle64):
    """homomorphic_shift_boolean: output must encrypt bit·2^51 under the big key; error compared with the oracle's"""
    ck, ctx = gpu64
    bits = np.array([0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0, 0], dtype=np.uint8)
    small = oracle64.keyswitch(ck.encrypt_bits(bits))
    got = ctx.stage_pbs(small)
    ref = oracle64.pbs(small)
    want = bits.astype(np.uint64) << np.uint64(51)
    e_gpu, e_ref = signed(ck.decrypt_phases(got) - want), signed(oracle64.phases(ref) - want)
    # noise after blind rotation is far below the 2^50 half-gap; both implementations sit in the same band
    assert np.abs(e_gpu).max() < 2.0**44 and np.abs(e_ref).max() < 2.0**44
    assert np.abs(e_gpu).std() < 4 * np.abs(e_ref).std() + 2.0**30
    # the two runs are the same computation up to f64 rounding: phases agree far below the noise bound
    assert np.abs(e_gpu - e_ref).max() < 2.0**42


def test_pbs_batch_size_independent(gpu64, oracle64):
    """The throughput kernel (3 ciphertexts per CTA) and the level-parallel small-batch kernel run the same arithmetic in
    the same order: a ciphertext bootstraps to the SAME words whatever batch it travels in."""
    ck, ctx = gpu64
    rng = np.random.default_rng(77)
    small = oracle64.keyswitch(ck.encrypt_bits(rng.integers(0, 2, 12)))
    big_batch = np.concatenate([small, rng.integers(0, 2**64, (600 - 12, small.shape[1]), dtype=np.uint64)])
    alone = ctx.stage_pbs(small)                 # 12 ciphertexts: pbs_wide_kernel
    together = ctx.stage_pbs(big_batch)[:12]     # 600 ciphertexts: pbs_merged_kernel, B = 3 (wave-count cost model)
    assert np.array_equal(alone, together)
    # the throughput kernel is instantiated with the shipped base log folded in at compile time and with a run-time one
    os.environ["TAC_PBS_GENERIC_BASE_LOG"] = "1"
    try:
        generic = ctx.stage_pbs(big_batch)[:12]
    finally:
        del os.environ["TAC_PBS_GENERIC_BASE_LOG"]
    assert np.array_equal(generic, together)
    # ragged tail: a batch that does not fill its last CTA
    ragged = ctx.stage_pbs(big_batch[:598])
    assert np.array_equal(ragged[:12], together) and np.array_equal(ragged[12:], ctx.stage_pbs(big_batch)[12:598])


def test_vertical_packing_from_oracle_ggsws(gpu64, oracle64, ol):
    ck, ctx = gpu64
    f = sbox_gal_mul_fn(ol)
    lut = ctx.generate_lookup_table(8, 24, f)
    vals = [0x00, 0x53]
    ggsw = np.stack([oracle64.pfks(oracle64.pbs(oracle64.keyswitch(ck.encrypt_bytes([v])[0]))) for v in vals])
    got = ctx.stage_vertical_packing(ggsw, lut)
    for i, v in enumerate(vals):
        ref = oracle64.vertical_packing(ggsw[i], 8, lut.table, 24)
        assert ck.decrypt_bytes(got[i]) == f(v).to_bytes(3, "big") == oracle64.decrypt_bytes(ref)
        want = ck.decrypt_bits(ref).astype(np.uint64) << np.uint64(63)
        e_gpu, e_ref = signed(ck.decrypt_phases(got[i]) - want), signed(ck.decrypt_phases(ref) - want)
        assert np.abs(e_gpu).max() < 2.0**59
        # Same computation up to f64 rounding (|Δ| ≈ 2^35 on the phase) — except at balanced-decomposition ties: the CMux
        # operands of vertical packing are differences of LUT values, i.e. 0 or 2^63 plus noise, and 2^63 is exactly the
        # top-digit tie of the l = 1, B = 2^13 decomposer.  tfhe's rule resolves it by the bit below, so the digit is
        # +B/2 or −B/2 depending on the SIGN of that noise; where the noise of a coefficient is itself below the f64
        # error, the two runs may pick opposite signs.  Both are exact decompositions of the same torus value, but the
        # noise terms they multiply differ by B·(GGSW noise) ≈ 2^13·2^41: a handful of outputs differ by up to ≈ 2^55,
        # well inside the noise envelope checked above (decoding margin 2^62).  So: typical outputs agree to rounding
        # error, every output stays inside the envelope.
        d = np.abs(e_gpu - e_ref)
        assert np.median(d) < 2.0**40 and np.count_nonzero(d < 2.0**42) >= 0.7 * d.size, np.log2(d + 1).round(1).tolist()
        assert d.max() < 2.0**57, np.log2(d + 1).round(1).tolist()


def test_vertical_packing_tie_free_all_outputs_tight(gpu64, oracle64, ol):
    """Same stage with a LUT whose CMux operands never sit on a decomposition tie: entries 2^61 (bit 0) and 3·2^61 (bit 1)
    still decode, and every difference of two entries is 0 or ±2^62 — top digit ±2^11, at least 2^61 away from the ±2^12
    tie — so both runs take the same digits everywhere and EVERY output must agree to f64 rounding (< 2^42)."""
    ck, ctx = gpu64
    tac = __import__("importlib").import_module("tfhe-aes-2_b200")
    f = sbox_gal_mul_fn(ol)
    base = ctx.generate_lookup_table(8, 24, f).table
    table = np.where(base != 0, np.uint64(3 << 61), np.uint64(1 << 61)).astype(np.uint64)
    table[:, 256:] = 0                                             # entries beyond 2^n_in stay 0 like the reference's
    lut = tac.LookupTable(table, 8, 24)
    vals = [0x00, 0x53, 0xA7]
    ggsw = np.stack([oracle64.pfks(oracle64.pbs(oracle64.keyswitch(ck.encrypt_bytes([v])[0]))) for v in vals])
    got = ctx.stage_vertical_packing(ggsw, lut)
    for i, v in enumerate(vals):
        ref = oracle64.vertical_packing(ggsw[i], 8, table, 24)
        assert ck.decrypt_bytes(got[i]) == f(v).to_bytes(3, "big") == oracle64.decrypt_bytes(ref)
        d = np.abs(signed(ck.decrypt_phases(got[i]) - ck.decrypt_phases(ref)))
        assert d.max() < 2.0**42, np.log2(d + 1).round(1).tolist()


@pytest.mark.parametrize("n_out", [1, 2, 5])
def test_vertical_packing_small_output_counts(gpu64, oracle64, n_out):
    """the vp_kernel instantiations for 1 and 2 outputs per CTA (and a ragged last CTA: 5 = 3 + 2) at phase level, with the
    tie-free LUT encoding of the test above"""
    ck, ctx = gpu64
    tac = __import__("importlib").import_module("tfhe-aes-2_b200")
    fn = lambda v: (v * 7 + 3) % (1 << n_out)
    base = ctx.generate_lookup_table(8, n_out, fn).table
    table = np.where(base != 0, np.uint64(3 << 61), np.uint64(1 << 61)).astype(np.uint64)
    table[:, 256:] = 0
    lut = tac.LookupTable(table, 8, n_out)
    v = 0xB5
    ggsw = oracle64.pfks(oracle64.pbs(oracle64.keyswitch(ck.encrypt_bytes([v])[0])))[None]
    got = ctx.stage_vertical_packing(ggsw, lut)[0]
    ref = oracle64.vertical_packing(ggsw[0], 8, table, n_out)
    want = [(fn(v) >> (n_out - 1 - o)) & 1 for o in range(n_out)]
    assert ck.decrypt_bits(got).tolist() == want == ck.decrypt_bits(ref).tolist()
    d = np.abs(signed(ck.decrypt_phases(got) - ck.decrypt_phases(ref)))
    assert d.max() < 2.0**42, np.log2(d + 1).round(1).tolist()


# ---------------------------------------------------------------------------------------------- the operator, decrypt-checked
def test_sbox_gal_mul_all_256_bytes(gpu64, ol):
    """BASELINE config 2: the 8-in/24-out SBOX·{1,2,3} WoP-PBS, all byte values in one batch"""
    ck, ctx = gpu64
    f = sbox_gal_mul_fn(ol)
    lut = ctx.generate_lookup_table(8, 24, f)
    cts = ck.encrypt_bytes(bytes(range(256)))
    out = ctx.circuit_bootstrap_batch(cts, lut)
    dec = ck.decrypt_bytes(out.reshape(-1, ck.params.big_lwe_size))
    assert dec == b"".join(f(v).to_bytes(3, "big") for v in range(256))
    # measured phase-error variance must stay inside the model's bound: output noise² = 8·NOMINAL (reference :325) and
    # max_noise_level_squared = 64 must still decode: after the worst-case 33-term sum the error stays below 2^62
    want = ck.decrypt_bits(out.reshape(-1, ck.params.big_lwe_size)).astype(np.uint64) << np.uint64(63)
    err = signed(ck.decrypt_phases(out) - want)
    sigma = err.std()
    assert sigma < 2.0**57, np.log2(sigma)
    assert 8 * np.sqrt(33.0) * sigma < 2.0**62                     # 8σ of the MixColumns+AddRoundKey sum still decodes


@pytest.mark.parametrize("batch", [1, 2, 40])
def test_wopbs_batch_sizes_match_oracle(gpu64, oracle64, ol, batch):
    """small batches take the B=1 / B=2 PBS kernels and the small GEMM tile; decrypted result equals the oracle's"""
    ck, ctx = gpu64
    rng = np.random.default_rng(batch)
    lut = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    vals = rng.integers(0, 256, batch).tolist()
    cts = ck.encrypt_bytes(bytes(vals))
    out = ctx.circuit_bootstrap_batch(cts, lut)
    assert ck.decrypt_bytes(out.reshape(-1, ck.params.big_lwe_size)) == bytes(ol.sbox(v) for v in vals)
    if batch <= 2:
        ref = oracle64.circuit_bootstrap(cts, lut.table, 8)
        assert oracle64.decrypt_bytes(ref.reshape(-1, oracle64.big1)) == bytes(ol.sbox(v) for v in vals)


def test_concurrent_callers_on_one_context(gpu64, ol):
    """The reference calls circuit_bootstrap from rayon workers — 16 SBOX bytes at a time (fhe_sbox_gal_mul_pbs.rs:33-41).
    16 host threads issue single-SBOX calls on ONE context: every result must decrypt correctly, and the coalescing entry
    point must have merged them into far fewer GPU passes than requests.  Plain tac_wopbs_batch calls are serialised by
    the context lock and must stay correct too."""
    import threading
    ck, ctx = gpu64
    f = sbox_gal_mul_fn(ol)
    lut = ctx.generate_lookup_table(8, 24, f)
    lut8 = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    lut.device_id(ctx); lut8.device_id(ctx)
    vals = [(37 * i + 11) % 256 for i in range(16)]
    bits = [[ck.encrypt(b) for b in __import__("importlib").import_module("tfhe-aes-2_b200").u8_to_bits(v)] for v in vals]
    ctx.set_coalescing(window_us=2000)
    before = ctx.coalescing_stats()
    res, errs = [None] * 16, []

    def worker(i):
        try:
            which = lut if i % 4 else lut8                         # two LUTs in flight: grouped per LUT inside a pass
            res[i] = ctx.circuit_bootstrap_coalesced(bits[i], which)
        except Exception as e:                                     # noqa: BLE001
            errs.append(e)

    for rounds in range(2):
        th = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
        [t.start() for t in th]; [t.join() for t in th]
        assert not errs, errs
        for i, v in enumerate(vals):
            got = bytes(__import__("importlib").import_module("tfhe-aes-2_b200").bits_to_u8([ck.decrypt(b) for b in res[i][8 * j:8 * j + 8]])
                        for j in range(len(res[i]) // 8))
            assert got == (f(v).to_bytes(3, "big") if i % 4 else bytes([ol.sbox(v)])), (i, got.hex())
    after = ctx.coalescing_stats()
    assert after["requests"] - before["requests"] == 32
    assert after["passes"] - before["passes"] <= 16, after              # 32 requests, two LUTs: merged, not one pass each
    # the plain batch entry point from many threads: serialised by the context lock
    out = [None] * 8
    def plain(i):
        out[i] = ctx.circuit_bootstrap_batch(np.stack([b.ct for b in bits[i]])[None], lut8)
    th = [threading.Thread(target=plain, args=(i,)) for i in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    for i in range(8):
        assert ck.decrypt_bytes(out[i][0]) == bytes([ol.sbox(vals[i])])
    ctx.set_coalescing(window_us=200)


def test_keys_and_ciphertexts_through_wire_files(tac, ck64, ol, tmp_path):
    """§8 f4 on the device: a fresh context takes its evaluation keys from a key file, inputs and outputs travel as LWE
    list files, and a client rebuilt from the file's secret sections decrypts the result"""
    kf, cf, of = (str(tmp_path / n) for n in ("keys.tac", "in.tac", "out.tac"))
    ck64.save_keys(kf, secret=True)
    ctx = tac.FheContext(tac.key_file_info(kf)[0])
    ctx.load_keys(kf)
    tac.save_lwe_list(cf, ck64.encrypt_bytes(b"\x53\x00"))
    lut = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    out = ctx.circuit_bootstrap_batch(tac.load_lwe_list(cf).reshape(2, 8, -1), lut)
    tac.save_lwe_list(of, out)
    ck2 = tac.ClientKey.load_secret_keys(kf)
    assert ck2.decrypt_bytes(tac.load_lwe_list(of)) == bytes([ol.sbox(0x53), ol.sbox(0x00)])
    other = tac.FheContext(4)
    with pytest.raises(RuntimeError, match="another parameter set"):
        other.load_keys(kf)


def test_two_level_circuit_bootstrap(tac, ck64, oracle64, ol):
    """cbs_level = 2 (base 2^8): one bootstrap + (k+1) PFKS per level, 2-level GGSWs in the vertical packing — decrypt-checked,
    against the oracle with the same parameters on the same keys (evaluation keys do not depend on the cbs decomposition), and
    with output noise no worse than the one-level set's"""
    p = tac.params_preset(64)
    p.cbs_level, p.cbs_base_log = 2, 8
    ctx = tac.FheContext(p)
    ctx.upload_keys(ck64)
    po = ol.preset(64)
    po.cbs_l, po.cbs_b = 2, 8
    orc = ol.Oracle(po, seed=0, raw=(ck64.sk_glwe, ck64.sk_lwe, ck64.bsk, ck64.ksk, ck64.pfpksk))
    lut = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    vals = [0x00, 0x53, 0xCA, 0xFF, 0x3C]
    cts = ck64.encrypt_bytes(bytes(vals))
    out = ctx.circuit_bootstrap_batch(cts, lut)
    assert ck64.decrypt_bytes(out.reshape(-1, p.big_lwe_size)) == bytes(ol.sbox(v) for v in vals)
    ref = orc.circuit_bootstrap(cts[:1], lut.table, 8)
    assert ck64.decrypt_bytes(ref.reshape(-1, p.big_lwe_size)) == bytes([ol.sbox(vals[0])])
    want = ck64.decrypt_bits(out.reshape(-1, p.big_lwe_size)).astype(np.uint64) << np.uint64(63)
    err = signed(ck64.decrypt_phases(out) - want)
    assert np.abs(err).max() < 2.0**60 and err.std() < 2.0**57
    bad = tac.params_preset(64)
    bad.cbs_level = 3
    with pytest.raises(RuntimeError, match="cbs_level"):
        tac.FheContext(bad)


def test_wopbs_empty_batch(gpu64, ol):
    ck, ctx = gpu64
    lut = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    out = ctx.circuit_bootstrap_batch(np.zeros((0, 8, ck.params.big_lwe_size), dtype=np.uint64), lut)
    assert out.shape == (0, 8, ck.params.big_lwe_size)


# ---------------------------------------------------------------------------------------------- reference model tests (shortint_woppbs_1bit.rs:531-617, :792-877)
@pytest.mark.parametrize("bits,words", [(3, [0b001, 0b000, 0b100, 0b101]), (8, [0b11001001, 0b01001001, 0b00101010, 0b11011001])])
def test_multivariate_parity_fn(gpu64, tac, bits, words):
    ck, ctx = gpu64
    parity_fn = lambda val: sum(tac.u16_to_bits(val)) % 2
    tv = ctx.generate_lookup_table(bits, 1, parity_fn)
    for word in words:
        bit_cts = [ck.encrypt(b) for b in tac.u16_to_bits(word)]
        d = ctx.circuit_bootstrap(bit_cts[16 - bits:], tv)[0]
        assert ck.decrypt(d) == parity_fn(word)
        assert d.noise_level.noise_level_squared == bits                  # NOMINAL × input_bit_count (:325)


@pytest.mark.parametrize("bits,words", [(3, [0b101, 0b000, 0b100]), (8, [0b11001001, 0b01001001, 0b00101010, 0b11011001])])
def test_multivariate_multivalued_square_fn(gpu64, tac, bits, words):
    ck, ctx = gpu64
    square_fn = lambda val: (val * val) % (1 << bits)
    tv = ctx.generate_lookup_table(bits, bits, square_fn)
    for byte in words:
        bit_cts = [ck.encrypt(b) for b in tac.u16_to_bits(byte)]
        out = ctx.circuit_bootstrap(bit_cts[16 - bits:], tv)
        got = sum(ck.decrypt(d) << (bits - 1 - i) for i, d in enumerate(out))
        assert got == square_fn(byte)


def test_increment_1bit_adder_with_trivial_carry(gpu64, tac):
    """reference :792-836 — a trivial (noiseless) ciphertext as circuit_bootstrap input"""
    ck, ctx = gpu64
    add_fn = lambda val: ((val >> 1) & 1) + (val & 1)
    lut = ctx.generate_lookup_table(2, 2, add_fn)
    value = [ck.encrypt(0) for _ in range(4)]                               # 4-bit counter, MSB first
    for expect in (1, 2):
        carry = ctx.trivial(1)
        new = []
        for bit in reversed(value):
            carry, nb = ctx.circuit_bootstrap([carry, bit], lut)
            new.append(nb)
        value = list(reversed(new))
        assert sum(ck.decrypt(b) << (3 - i) for i, b in enumerate(value)) == expect


def test_increment_8bit_adder_9_to_9(gpu64, tac):
    """reference test_increment_8bit_adder (:838-877): a 9→9 LUT — n_in = log2 N, nine blind-rotation steps, the first by
    X^-256 — with a trivial carry-in, rippling through all 16 bytes, three increments of 0x…00FF"""
    ck, ctx = gpu64
    add_fn = lambda val: (val & 0xFF) + tac.u16_to_bits(val)[7]
    lut = ctx.generate_lookup_table(9, 9, add_fn)
    value_clear = bytes(15) + b"\xff"
    value = [[ck.encrypt(b) for b in tac.u8_to_bits(v)] for v in value_clear]
    for _ in range(3):
        carry = ctx.trivial(1)
        new = []
        for byte in reversed(value):
            out = ctx.circuit_bootstrap([carry] + byte, lut)
            carry, nb = out[0], out[1:]
            assert nb[0].noise_level.noise_level_squared == 9
            new.append(nb)
        value = list(reversed(new))
    got = bytes(tac.bits_to_u8([ck.decrypt(b) for b in byte]) for byte in value)
    assert got == bytes(14) + b"\x01\x02"


def test_xor_of_bootstrapped_bits_respects_bookkeeping(gpu64, tac):
    ck, ctx = gpu64
    ident = ctx.generate_lookup_table(1, 1, lambda b: b)
    a, b = ck.encrypt(1), ck.encrypt(0)
    a2 = ctx.circuit_bootstrap([a], ident)[0]
    b2 = ctx.circuit_bootstrap([b], ident)[0]
    c = a2 ^ b2
    assert ck.decrypt(c) == 1 and c.noise_level.noise_level_squared == 2
    with pytest.raises(AssertionError, match="noise components not independent"):
        _ = c ^ a2


# ---------------------------------------------------------------------------------------------- AES
def test_aes_light_two_rounds(gpu64, ol):
    """reference test_light_gal_mul (fhe_impls/shortint_woppbs_1bit.rs:185-193 → test_helper.rs:86-120): 2 rounds,
    clear key schedule encrypted directly, one block from the ChaCha20 seed-0 stream; BASELINE config 3."""
    ck, ctx = gpu64
    s = ol.chacha20_stream(32)
    key, blk = s[:16], s[16:32]
    ctx.aes_set_key_schedule(ck.encrypt_bytes(bytes(ol.plain_key_schedule(key))))
    enc = ctx.aes_encrypt_blocks(ck.encrypt_bytes(blk)[None], rounds=2)
    assert ck.decrypt_bytes(enc[0]).hex() == "5c864f984df12113a07c22a99f49f0a1" == ol.plain_encrypt_block(key, blk, 2).hex()
    enc1 = ctx.aes_encrypt_blocks(ck.encrypt_bytes(blk)[None], rounds=1)
    assert ck.decrypt_bytes(enc1[0]).hex() == "de3011192c24fd50c3b199187f869fa4"


def test_aes_light_matches_oracle_phase_band(gpu64, oracle64, ol):
    """same 2-round block through the oracle on the same ciphertexts: same plaintext, comparable output noise"""
    ck, ctx = gpu64
    s = ol.chacha20_stream(32)
    key, blk = s[:16], s[16:32]
    ks_ct = ck.encrypt_bytes(bytes(ol.plain_key_schedule(key)), first_index=10_000)
    b_ct = ck.encrypt_bytes(blk, first_index=50_000)
    ctx.aes_set_key_schedule(ks_ct)
    got = ctx.aes_encrypt_blocks(b_ct[None], rounds=2)[0]
    ref = oracle64.aes_encrypt_blocks(ks_ct, b_ct[None], rounds=2)[0]
    assert ck.decrypt_bytes(got) == ck.decrypt_bytes(ref) == ol.plain_encrypt_block(key, blk, 2)
    want = ck.decrypt_bits(ref).astype(np.uint64) << np.uint64(63)
    e_gpu, e_ref = signed(ck.decrypt_phases(got) - want), signed(ck.decrypt_phases(ref) - want)
    assert e_gpu.std() < 2 * e_ref.std() and np.abs(e_gpu).max() < 2.0**60


def test_aes_cli_stream_10_blocks(gpu64, ol):
    """BASELINE config 4 (reference src/bin/main.rs with --number-of-outputs 10): counter blocks iv ‖ BE64(ctr)"""
    from test_oracle_golden import CLI_IV, CLI_KEY, CLI_OUT
    ck, ctx = gpu64
    ctx.aes_set_key_schedule(ck.encrypt_bytes(bytes(ol.plain_key_schedule(CLI_KEY))))
    blocks = np.stack([ck.encrypt_bytes(CLI_IV + ctr.to_bytes(8, "big")) for ctr in range(1, 11)])
    enc = ctx.aes_encrypt_blocks(blocks)
    assert [ck.decrypt_bytes(e).hex() for e in enc] == CLI_OUT


def test_aes_noise_guard(tac):
    """the fused path refuses a circuit whose squared noise would exceed the parameter set's maximum (NoiseTooBig)"""
    ck = tac.ClientKey(4, seed=1).gen_eval_keys()
    ctx = tac.FheContext(ck.params)
    ctx.upload_keys(ck)
    ctx.aes_set_key_schedule(np.zeros((44, 4, 8, ck.params.big_lwe_size), dtype=np.uint64))
    with pytest.raises(tac.NoiseTooBig):
        ctx.aes_encrypt_blocks(np.zeros((1, 16, 8, ck.params.big_lwe_size), dtype=np.uint64), rounds=2)


def test_fhe_key_schedule_fips197(gpu64, ol):
    """reference test_full_gal_mul / FIPS-197 C.1 (test_helper.rs:61-83): FHE key schedule on the device + 10 rounds"""
    ck, ctx = gpu64
    key = bytes.fromhex("000102030405060708090a0b0c0d0e0f")
    blk = bytes.fromhex("00112233445566778899aabbccddeeff")
    ks = ctx.aes_key_schedule(ck.encrypt_bytes(key))
    assert ck.decrypt_bytes(ks.reshape(-1, ck.params.big_lwe_size)) == bytes(ol.plain_key_schedule(key))
    enc = ctx.aes_encrypt_blocks(ck.encrypt_bytes(blk)[None])
    assert ck.decrypt_bytes(enc[0]).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"


# ---------------------------------------------------------------------------------------------- other parameter sets (N = 1024, k = 2)
def test_lvl1_cmux_tree_16_to_8(tac):
    """reference test_multivariate_multivalues_xor_8bit (:626-659): 16 inputs at N = 1024 — the real CMux tree"""
    ck, ctx = tac.FheContext.generate_keys(1, seed=SEED)
    b1, b2 = 0b11000110, 0b10101010
    xor_fn = lambda v: (v >> 8) ^ (v & 0xFF)
    tv = ctx.generate_lookup_table(16, 8, xor_fn)
    out = ctx.circuit_bootstrap_batch(ck.encrypt_bytes([b1, b2]).reshape(1, 16, -1), tv)
    assert ck.decrypt_bytes(out[0]) == bytes([b1 ^ b2])


def test_lvl1_pfks_integer_gemm_bit_exact(tac, ol):
    """params_sqrd_lvl_1 has pfks_base_log = 24: digits wider than 16 bits take the 64-bit integer-pipe GEMM
    (lwe_gemm_kernel) instead of the tcgen05 path — bit-exact against the oracle on the same keys, ties included"""
    ck = tac.ClientKey(1, seed=SEED + 1).gen_eval_keys()
    ctx = tac.FheContext(ck.params)
    ctx.upload_keys(ck)
    orc = ol.Oracle(1, seed=SEED + 1, raw=(ck.sk_glwe, ck.sk_lwe, ck.bsk, ck.ksk, ck.pfpksk))
    rng = np.random.default_rng(31)
    for n in (1, 19, 270):
        x = rng.integers(0, 2**64, (n, ck.params.big_lwe_size), dtype=np.uint64)
        x[0, :6] = [0, 2**64 - 1, 1 << 63, (1 << 63) - 1, 1 << 39, (1 << 63) + (1 << 39)]
        assert np.array_equal(ctx.stage_pfks(x), orc.pfks(x)), n
    assert np.array_equal(ctx.stage_keyswitch(x[:5]), orc.keyswitch(x[:5]))


@pytest.mark.parametrize("pid", [4, 256])
def test_other_parameter_sets_sbox(tac, ol, pid):
    ck, ctx = tac.FheContext.generate_keys(pid, seed=SEED + pid)
    lut = ctx.generate_lookup_table(8, 8, lambda b: ol.sbox(b))
    vals = [0x00, 0x53, 0xCA]
    out = ctx.circuit_bootstrap_batch(ck.encrypt_bytes(bytes(vals)), lut)
    assert ck.decrypt_bytes(out.reshape(-1, ck.params.big_lwe_size)) == bytes(ol.sbox(v) for v in vals)
