"""CPU-side checks of the product: the client library against the oracle (two independent implementations of the same
key/noise specification must agree bit for bit), the host-side noise bookkeeping, and the C ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(tac):
    header = open(os.path.join(ROOT, "include", "tfhe_aes_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(tac_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 40
    nm = subprocess.check_output(["nm", "-D", "--defined-only", tac.library_path()], text=True)
    exported = set(re.findall(r"\bT (tac_[a-z0-9_]+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/tfhe_aes_cuda.h but not exported: {missing}"
    # the Python binding covers the whole header too
    assert sorted(tac._SIGNATURES) == declared
    tac.load_library()


def test_library_has_sm100a_kernels(tac):
    out = subprocess.run(["cuobjdump", "-lelf", tac.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_presets_match_reference(tac, ol):
    # reference parameters.rs:29-61, :77-109, :125-157, :173-205
    want = {1: (671, 2, 1024, 2, 15, 4, 3, 1, 10, 1, 24, 1), 4: (679, 2, 1024, 2, 15, 4, 3, 1, 11, 2, 16, 4),
            64: (677, 4, 512, 3, 12, 4, 3, 1, 13, 2, 16, 64), 256: (665, 2, 1024, 4, 9, 6, 2, 1, 14, 3, 12, 256)}
    for pid, vals in want.items():
        p, q = tac.params_preset(pid), ol.preset(pid)
        got = (p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_level, p.pbs_base_log, p.ks_level, p.ks_base_log,
               p.cbs_level, p.cbs_base_log, p.pfks_level, p.pfks_base_log, p.max_noise_level_squared)
        assert got == vals
        assert got == (q.n, q.k, q.N, q.pbs_l, q.pbs_b, q.ks_l, q.ks_b, q.cbs_l, q.cbs_b, q.pfks_l, q.pfks_b, q.max_noise_sq)
        assert (p.lwe_noise_std, p.glwe_noise_std, p.pfks_noise_std) == (q.s_lwe, q.s_glwe, q.s_pfks)
    with pytest.raises(ValueError):
        tac.params_preset(7)


def test_client_keys_equal_oracle_keys(ck64, oracle64):
    for name in ("sk_glwe", "sk_lwe", "bsk", "ksk", "pfpksk"):
        assert np.array_equal(getattr(ck64, name), getattr(oracle64, name)), name
    assert set(np.unique(ck64.sk_glwe)) <= {0, 1}
    assert 0.4 < ck64.sk_glwe.mean() < 0.6


def test_client_encrypt_decrypt_equal_oracle(ck64, oracle64):
    bits = [0, 1, 1, 0, 1, 0, 0, 1, 1]
    a = ck64.encrypt_bits(bits, first_index=1234)
    assert np.array_equal(a, oracle64.encrypt_bits(bits, first_index=1234))
    assert ck64.decrypt_bits(a).tolist() == bits
    assert np.array_equal(ck64.decrypt_phases(a), oracle64.phases(a))
    # fresh noise: sigma_lwe·2^64 ≈ 2^49.6
    err = (ck64.decrypt_phases(a) - (np.array(bits, dtype=np.uint64) << np.uint64(63))).astype(np.int64)
    assert 2.0**44 < np.abs(err).astype(float).max() < 2.0**53
    with pytest.raises(AssertionError, match="cleartext out of bounds"):
        ck64.encrypt_bits([2])
    assert ck64.encrypt_bits([]).shape == (0, ck64.params.big_lwe_size)


def test_client_default_seed_is_os_entropy(tac):
    """a default ClientKey must not be derivable from a public constant (the reference seeds from the OS,
    engine.rs:164-168); an explicit seed stays reproducible for tests"""
    a, b, z = tac.ClientKey(64), tac.ClientKey(64), tac.ClientKey(64, seed=0)
    assert a.seed is None
    assert not np.array_equal(a.sk_glwe, b.sk_glwe) and not np.array_equal(a.sk_glwe, z.sk_glwe)
    assert np.array_equal(z.sk_glwe, tac.ClientKey(64, seed=0).sk_glwe)
    bits = [1, 0, 1, 1]
    ca, cb = a.encrypt_bits(bits, first_index=0), a.encrypt_bits(bits, first_index=0)
    assert a.decrypt_bits(ca).tolist() == bits
    assert np.array_equal(ca, cb)                                   # same instance, same index: same stream (documented)
    # a client rebuilt around the same secret keys draws fresh masks: index 0 of the new instance is not a reuse
    a2 = tac.ClientKey(64, secret_keys=(a.sk_glwe, a.sk_lwe))
    c2 = a2.encrypt_bits(bits, first_index=0)
    assert not np.array_equal(c2[:, :8], ca[:, :8])
    assert a.decrypt_bits(c2).tolist() == bits and a2.decrypt_bits(ca).tolist() == bits
    with pytest.raises(RuntimeError):
        tac.ClientKey(64, secret_keys=(a.sk_glwe + np.uint64(2), a.sk_lwe))
    # key views keep their owner alive
    k = tac.ClientKey(64, seed=5).sk_lwe
    import gc; gc.collect()
    assert set(np.unique(k)) <= {0, 1} and k.size == 677


def test_wire_format_round_trip(tac, ck64, tmp_path):
    """§8 f4: flat dump of the raw tfhe-rs containers — keys and ciphertext lists survive a round trip bit for bit, other
    parameter sets and corrupted payloads are refused"""
    import ctypes as C
    path = str(tmp_path / "keys.tac")
    ck64.save_keys(path, secret=True)
    p, present = tac.key_file_info(path)
    assert present == {1, 2, 3, 4, 5} and p.lwe_dimension == 677 and p.polynomial_size == 512
    L = tac.load_library()
    for which, name in enumerate(("sk_glwe", "sk_lwe", "bsk", "ksk", "pfpksk")):
        out = np.empty(L.tac_key_len(C.byref(p), which), dtype=np.uint64)
        args = [None] * 5
        args[which] = out.ctypes.data
        assert L.tac_keys_load(path.encode(), C.byref(p), *args) == 0
        assert np.array_equal(out, getattr(ck64, name)), name
    # header: magic, version, section count, the parameter block
    raw = open(path, "rb").read(96)
    assert raw[:8] == b"TACWIRE\x01" and int.from_bytes(raw[8:12], "little") == 1 and int.from_bytes(raw[12:16], "little") == 5
    assert np.frombuffer(raw[16:64], dtype=np.int32).tolist() == [677, 4, 512, 3, 12, 4, 3, 1, 13, 2, 16, 64]
    # a client rebuilt from the file decrypts what the original encrypted
    ck2 = tac.ClientKey.load_secret_keys(path)
    bits = [1, 0, 0, 1, 1]
    assert ck2.decrypt_bits(ck64.encrypt_bits(bits, first_index=777)).tolist() == bits
    assert ck64.decrypt_bits(ck2.encrypt_bits(bits)).tolist() == bits
    # evaluation-only file: no secret sections
    pub = str(tmp_path / "eval.tac")
    ck64.save_keys(pub)
    assert tac.key_file_info(pub)[1] == {3, 4, 5}
    with pytest.raises(OSError):
        tac.ClientKey.load_secret_keys(pub)
    # wrong parameter set / flipped payload bit are refused
    other = tac.params_preset(4)
    out = np.empty(L.tac_key_len(C.byref(other), 1), dtype=np.uint64)
    assert L.tac_keys_load(path.encode(), C.byref(other), None, out.ctypes.data, None, None, None) != 0
    blob = bytearray(open(path, "rb").read())
    blob[96 + 40 + 1000] ^= 0x01                                    # inside the first payload
    bad = str(tmp_path / "bad.tac")
    open(bad, "wb").write(blob)
    out = np.empty(2048, dtype=np.uint64)
    assert L.tac_keys_load(bad.encode(), C.byref(p), out.ctypes.data, None, None, None, None) != 0
    # ciphertext lists
    cts = ck64.encrypt_bytes(b"\x53\xca")
    lst = str(tmp_path / "cts.tac")
    tac.save_lwe_list(lst, cts)
    back = tac.load_lwe_list(lst)
    assert back.shape == (16, 2049) and np.array_equal(back, cts.reshape(16, 2049))
    with pytest.raises(OSError):
        tac.load_lwe_list(path)                                     # a key file is not an LWE list


def test_context_rejects_unsupported_parameter_sets(tac):
    """constraints the kernels assume beyond (N, k) are validated up front (before any device is touched)"""
    import ctypes as C
    L = tac.load_library()
    def err_for(**kw):
        p = tac.params_preset(64)
        for k, v in kw.items():
            setattr(p, k, v)
        h = L.tac_ctx_create(C.byref(p), 0)
        if h:
            L.tac_ctx_destroy(h)
            return None
        return L.tac_last_error(None).decode()
    assert "base_log must be <= 7" in err_for(ks_base_log=8)
    assert "must be odd" in err_for(lwe_dimension=676)
    assert "16-bit fields" in err_for(pbs_base_log=16)
    assert "cbs_level" in err_for(cbs_level=3)
    assert "below 16000" in err_for(ks_level=8, ks_base_log=2)
    assert "unsupported (polynomial_size" in err_for(polynomial_size=2048)


def test_client_other_parameter_sets(tac, ol):
    ck = tac.ClientKey(4, seed=3)
    assert ck.decrypt_bits(ck.encrypt_bits([1, 0, 1])).tolist() == [1, 0, 1]
    o = ol.Oracle(4, seed=3)
    ck.gen_eval_keys()
    assert np.array_equal(ck.ksk, o.ksk) and np.array_equal(ck.bsk, o.bsk) and np.array_equal(ck.pfpksk, o.pfpksk)


# --- host-side noise bookkeeping: reference shortint_woppbs_1bit.rs:484-529 (KEYS_SQRD_LVL_4)
def test_bit_xor(tac):
    ck = tac.ClientKey(4, seed=11)
    b1, b2, b3, b4 = (ck.encrypt(v) for v in (0, 1, 0, 1))
    assert ck.decrypt(b1 ^ b2) == 1
    assert ck.decrypt(b1 ^ b3) == 0
    assert ck.decrypt(b2 ^ b4) == 0
    t0 = ck.context.trivial(0)
    assert ck.decrypt(b2 ^ t0) == 1
    _ = t0 ^ t0 ^ t0                      # trivial does not accumulate noise
    assert ck.decrypt(ck.context.trivial(1)) == 1


def test_bit_xor_above_max_noise(tac):
    ck = tac.ClientKey(4, seed=11)
    bs = [ck.encrypt(v) for v in (0, 1, 0, 1, 0)]
    with pytest.raises(tac.NoiseTooBig, match="NoiseTooBig"):
        _ = bs[0] ^ bs[1] ^ bs[2] ^ bs[3] ^ bs[4]


def test_bit_xor_not_independent(tac):
    ck = tac.ClientKey(4, seed=11)
    b1 = ck.encrypt(0)
    with pytest.raises(AssertionError, match="noise components not independent"):
        _ = b1 ^ b1


def test_context_creation_fails_loudly_without_gpu(tac):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="tac_ctx_create failed"):
        tac.FheContext(64)
