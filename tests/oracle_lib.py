"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "liboracle.so")


class Params(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("k", C.c_int32), ("N", C.c_int32),
        ("pbs_l", C.c_int32), ("pbs_b", C.c_int32),
        ("ks_l", C.c_int32), ("ks_b", C.c_int32),
        ("cbs_l", C.c_int32), ("cbs_b", C.c_int32),
        ("pfks_l", C.c_int32), ("pfks_b", C.c_int32),
        ("max_noise_sq", C.c_int32),
        ("s_lwe", C.c_double), ("s_glwe", C.c_double), ("s_pfks", C.c_double),
    ]

    @property
    def big(self):
        return self.k * self.N


def build(force=False):
    srcs = [os.path.join(_ORACLE_DIR, f) for f in ("oracle.cpp", "capi.cpp", "oracle.hpp", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = build()
    L = C.CDLL(so)
    L.orc_params_preset.argtypes = [C.c_int, C.POINTER(Params)]
    L.orc_keygen.restype = C.c_void_p
    L.orc_keygen.argtypes = [C.POINTER(Params), C.c_uint64]
    L.orc_keyset_from_raw.restype = C.c_void_p
    L.orc_keyset_from_raw.argtypes = [C.POINTER(Params), C.c_uint64, _u64p, _u64p, _u64p, _u64p, _u64p]
    L.orc_keyset_free.argtypes = [C.c_void_p]
    L.orc_key_ptr.restype = C.POINTER(C.c_uint64)
    L.orc_key_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    L.orc_encrypt_bits.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_uint64, _u64p]
    L.orc_decrypt_phases.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_decrypt_phases_small.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_glwe_phase.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_decrypt_bits.argtypes = [C.c_void_p, _u64p, C.c_int, _u8p]
    L.orc_decompose.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, _i64p]
    L.orc_keyswitch_batch.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_pbs_batch.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_pfks_batch.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_extract_bits_batch.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, C.c_int, _u64p]
    L.orc_external_product.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p, _u64p]
    L.orc_vertical_packing.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p, C.c_int, _u64p]
    L.orc_circuit_bootstrap_batch.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p, C.c_int, _u64p]
    L.orc_lut_len.restype = C.c_size_t
    L.orc_lut_len.argtypes = [C.c_int, C.c_int]
    L.orc_generate_lut.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u64p]
    L.orc_encode_bit.restype = C.c_uint64
    L.orc_encode_bit.argtypes = [C.c_uint64]
    L.orc_decode_bit.restype = C.c_uint64
    L.orc_decode_bit.argtypes = [C.c_uint64]
    L.orc_chacha20_stream.argtypes = [_u8p, C.c_uint64, _u8p, C.c_size_t]
    L.orc_sbox.restype = C.c_uint8
    L.orc_sbox.argtypes = [C.c_int]
    L.orc_gf_256_mul.restype = C.c_uint8
    L.orc_gf_256_mul.argtypes = [C.c_uint8, C.c_uint8]
    L.orc_plain_key_schedule.argtypes = [_u8p, _u8p]
    L.orc_plain_encrypt_block.argtypes = [_u8p, _u8p, C.c_int, _u8p]
    L.orc_aes_encrypt_blocks.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p, _u64p, C.c_int]
    L.orc_aes_key_schedule.argtypes = [C.c_void_p, _u64p, _u64p]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def preset(pid):
    p = Params()
    assert lib().orc_params_preset(pid, C.byref(p)) == 0, f"unknown preset {pid}"
    return p


def u8_to_bits(v):
    """MSB-first bits of a byte (reference src/util.rs:33-35)."""
    return [(v >> (7 - i)) & 1 for i in range(8)]


def bits_to_u8(bits):
    return sum(int(b) << (7 - i) for i, b in enumerate(bits))


class Oracle:
    """Key set + every stage of the reference path on the CPU."""

    def __init__(self, pid=64, seed=0, raw=None):
        self.L = lib()
        self.p = preset(pid) if isinstance(pid, int) else pid
        self.seed = seed
        if raw is None:
            self.h = self.L.orc_keygen(C.byref(self.p), seed)
        else:
            self.h = self.L.orc_keyset_from_raw(C.byref(self.p), seed, *[np.ascontiguousarray(a, dtype=np.uint64) for a in raw])
        self.big1 = self.p.big + 1

    def __del__(self):
        try:
            if self.h:
                self.L.orc_keyset_free(self.h)
                self.h = None
        except Exception:
            pass

    def key(self, which):
        ln = C.c_size_t()
        ptr = self.L.orc_key_ptr(self.h, which, C.byref(ln))
        return np.ctypeslib.as_array(ptr, shape=(ln.value,))

    @property
    def sk_glwe(self): return self.key(0)
    @property
    def sk_lwe(self): return self.key(1)
    @property
    def bsk(self): return self.key(2)
    @property
    def ksk(self): return self.key(3)
    @property
    def pfpksk(self): return self.key(4)

    def encrypt_bits(self, bits, first_index=0):
        bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
        out = np.empty((bits.size, self.big1), dtype=np.uint64)
        self.L.orc_encrypt_bits(self.h, bits, bits.size, first_index, out)
        return out

    def encrypt_bytes(self, data, first_index=0):
        bits = [b for v in data for b in u8_to_bits(v)]
        return self.encrypt_bits(bits, first_index).reshape(len(data), 8, self.big1)

    def phases(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.big1)
        out = np.empty(cts.shape[0], dtype=np.uint64)
        self.L.orc_decrypt_phases(self.h, cts, cts.shape[0], out)
        return out

    def phases_small(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.p.n + 1)
        out = np.empty(cts.shape[0], dtype=np.uint64)
        self.L.orc_decrypt_phases_small(self.h, cts, cts.shape[0], out)
        return out

    def glwe_phases(self, glwes):
        W = (self.p.k + 1) * self.p.N
        g = np.ascontiguousarray(glwes, dtype=np.uint64).reshape(-1, W)
        out = np.empty((g.shape[0], self.p.N), dtype=np.uint64)
        self.L.orc_glwe_phase(self.h, g, g.shape[0], out)
        return out

    def decrypt_bits(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.big1)
        out = np.empty(cts.shape[0], dtype=np.uint8)
        self.L.orc_decrypt_bits(self.h, cts, cts.shape[0], out)
        return out

    def decrypt_bytes(self, cts):
        bits = self.decrypt_bits(cts).reshape(-1, 8)
        return bytes(bits_to_u8(r) for r in bits)

    def keyswitch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.big1)
        out = np.empty((cts.shape[0], self.p.n + 1), dtype=np.uint64)
        self.L.orc_keyswitch_batch(self.h, cts, cts.shape[0], out)
        return out

    def pbs(self, small):
        small = np.ascontiguousarray(small, dtype=np.uint64).reshape(-1, self.p.n + 1)
        out = np.empty((small.shape[0], self.big1), dtype=np.uint64)
        self.L.orc_pbs_batch(self.h, small, small.shape[0], out)
        return out

    def extract_bits(self, bigs, delta_log, n_bits):
        """[U] wop_pbs.rs::extract_bits: [n][big+1] → [n][n_bits][small+1], most significant extracted bit first"""
        bigs = np.ascontiguousarray(bigs, dtype=np.uint64).reshape(-1, self.big1)
        out = np.empty((bigs.shape[0], n_bits, self.p.n + 1), dtype=np.uint64)
        self.L.orc_extract_bits_batch(self.h, bigs, bigs.shape[0], delta_log, n_bits, out)
        return out

    def pfks(self, bigs):
        bigs = np.ascontiguousarray(bigs, dtype=np.uint64).reshape(-1, self.big1)
        G = self.p.k + 1
        out = np.empty((bigs.shape[0], G, G * self.p.N), dtype=np.uint64)
        self.L.orc_pfks_batch(self.h, bigs, bigs.shape[0], out)
        return out

    def external_product(self, ggsw_std, levels, b, glwe_in, acc):
        acc = np.array(acc, dtype=np.uint64, copy=True)
        self.L.orc_external_product(self.h, np.ascontiguousarray(ggsw_std, dtype=np.uint64), levels, b,
                                    np.ascontiguousarray(glwe_in, dtype=np.uint64), acc)
        return acc

    def vertical_packing(self, ggsw_std, n_in, lut, n_out):
        out = np.empty((n_out, self.big1), dtype=np.uint64)
        self.L.orc_vertical_packing(self.h, np.ascontiguousarray(ggsw_std, dtype=np.uint64), n_in,
                                    np.ascontiguousarray(lut, dtype=np.uint64), n_out, out)
        return out

    def generate_lookup_table(self, n_in, n_out, f):
        """reference FheContext::generate_lookup_table (shortint_woppbs_1bit.rs:274-289)."""
        return generate_lut(n_in, n_out, self.p.N, f)

    def circuit_bootstrap(self, bits, lut, n_out):
        """reference FheContext::circuit_bootstrap (:292-336); bits: [batch][n_in][big+1] or [n_in][big+1]."""
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        single = bits.ndim == 2
        if single:
            bits = bits[None]
        batch, n_in = bits.shape[0], bits.shape[1]
        out = np.empty((batch, n_out, self.big1), dtype=np.uint64)
        self.L.orc_circuit_bootstrap_batch(self.h, bits, batch, n_in, np.ascontiguousarray(lut, dtype=np.uint64), n_out, out)
        return out[0] if single else out

    def aes_encrypt_blocks(self, key_sched, blocks, rounds=10, in_noise_sq=1):
        blocks = np.ascontiguousarray(blocks, dtype=np.uint64).reshape(-1, 16, 8, self.big1)
        ksd = np.ascontiguousarray(key_sched, dtype=np.uint64).reshape(44, 4, 8, self.big1)
        out = np.empty_like(blocks)
        rc = self.L.orc_aes_encrypt_blocks(self.h, ksd, blocks.shape[0], rounds, blocks, out, in_noise_sq)
        if rc != 0:
            raise RuntimeError("NoiseTooBig")
        return out

    def aes_key_schedule(self, key_bits):
        kb = np.ascontiguousarray(key_bits, dtype=np.uint64).reshape(16, 8, self.big1)
        out = np.empty((44, 4, 8, self.big1), dtype=np.uint64)
        rc = self.L.orc_aes_key_schedule(self.h, kb, out)
        if rc != 0:
            raise RuntimeError("NoiseTooBig")
        return out


def generate_lut(n_in, n_out, N, f):
    L = lib()
    table = np.array([f(v) for v in range(1 << n_in)], dtype=np.uint64)
    out = np.empty((n_out, L.orc_lut_len(n_in, N)), dtype=np.uint64)
    L.orc_generate_lut(n_in, n_out, N, table, out)
    return out


def chacha20_stream(n, key=bytes(32), nonce=0):
    out = np.empty(n, dtype=np.uint8)
    lib().orc_chacha20_stream(np.frombuffer(key, dtype=np.uint8).copy(), nonce, out, n)
    return bytes(out)


def plain_key_schedule(key):
    ek = np.empty(176, dtype=np.uint8)
    lib().orc_plain_key_schedule(np.frombuffer(bytes(key), dtype=np.uint8).copy(), ek)
    return ek


def plain_encrypt_block(key, block, rounds=10):
    ek = plain_key_schedule(key)
    out = np.empty(16, dtype=np.uint8)
    lib().orc_plain_encrypt_block(ek, np.frombuffer(bytes(block), dtype=np.uint8).copy(), rounds, out)
    return bytes(out)


def sbox(i):
    return lib().orc_sbox(i)


def gf_256_mul(a, b):
    return lib().orc_gf_256_mul(a, b)
