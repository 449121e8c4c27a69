"""Runs the CUDA kernels' phase functions (csrc/ep_core.cuh, ep_step.cuh) on the CPU, thread by thread in the kernels'
phase order (tests/cpu/ep_emul.cpp), and checks them against numpy and the oracle: FFT passes, swizzles, slot order,
rotation, decomposition, MAC and inverse.  Same code the GPU executes, so index bugs are caught without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpu"), "-s"])
    E = C.CDLL(os.path.join(HERE, "cpu", "libep_emul.so"))
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    up = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
    E.emul_fft_fwd_inv.argtypes = [C.c_int, dp, dp, dp, ip]
    E.emul_cmux_step.argtypes = [C.c_int] * 5 + [up, C.c_int, ip, up]
    E.emul_cmux_step_merged.argtypes = [C.c_int] * 5 + [up, C.c_int, ip, up]
    E.emul_digits.argtypes = [C.c_uint64, C.c_int, C.c_int, dp]
    E.emul_f64_to_torus.restype = C.c_uint64
    E.emul_f64_to_torus.argtypes = [C.c_double]
    return E


@pytest.mark.parametrize("N", [512, 1024])
def test_fft_matches_numpy(emul, N):
    rng = np.random.default_rng(N)
    M = N // 2
    x = rng.integers(-2048, 2048, N).astype(np.float64)
    ore, oim, sf = np.zeros(M + N), np.zeros(M + N), np.zeros(M, dtype=np.int32)
    assert emul.emul_fft_fwd_inv(N, x, ore, oim, sf) == 0
    z = (x[:M] + 1j * x[M:]) * np.exp(1j * np.pi * np.arange(M) / N)
    X = np.fft.fft(z)
    assert sorted(sf.tolist()) == list(range(M))                       # slot order is a permutation of the frequencies
    assert np.abs((ore[:M] + 1j * oim[:M]) - X[sf]).max() < 1e-9
    assert np.abs(ore[M:] - x).max() < 1e-10                            # inverse(forward) == identity


@pytest.mark.parametrize("b,l", [(12, 3), (13, 1), (15, 2), (9, 4), (10, 1), (11, 1), (14, 1)])
def test_f64_digits_equal_iterator(emul, ol, b, l):
    """the closed-form + tie-replay digits of the f64 path must equal the bit-exact iterator, ties included"""
    rng = np.random.default_rng(b * 10 + l)
    xs = [int(v) for v in rng.integers(0, 2**64, 3000, dtype=np.uint64)]
    nonrep = 64 - b * l
    # force ties: some raw digit == B/2 with everything below it zero / rounding up into it
    for lev in range(1, l + 1):
        base = (1 << (b - 1)) << (64 - b * lev)
        for extra in (0, 1 << (nonrep - 1), (1 << (nonrep - 1)) - 1, rng.integers(0, 2**63)):
            hi = int(rng.integers(0, 2**63)) >> (b * (l - lev) + nonrep) << (b * (l - lev) + nonrep) if lev > 1 else 0
            xs.append((base + int(extra) % (1 << (64 - b * lev)) + (hi << 0)) % 2**64)
            xs.append((base + int(extra) % (1 << nonrep)) % 2**64)
    xs += [0, 2**63, 2**64 - 1, 2**63 - 1, (1 << 63) + (1 << (nonrep - 1))]
    for x in xs:
        d = np.zeros(l)
        emul.emul_digits(x, b, l, d)
        want = np.zeros(l, dtype=np.int64)
        ol.lib().orc_decompose(x, b, l, 0, want)
        assert d.tolist() == want.astype(float).tolist(), hex(x)


def test_f64_to_torus(emul):
    assert emul.emul_f64_to_torus(0.25) == 1 << 62
    assert emul.emul_f64_to_torus(-0.25) == 3 << 62
    assert emul.emul_f64_to_torus(12345.5) == 1 << 63
    assert emul.emul_f64_to_torus(-7.0) == 0
    assert emul.emul_f64_to_torus(3.0 + 2.0**-20) == 1 << 44


def _rot_diff(a, r, N):
    idx = (np.arange(N) - r) % (2 * N)
    rotd = np.where(idx < N, a[:, idx % N], (-a[:, idx % N].astype(np.int64)).astype(np.uint64))
    return rotd - a


CASES = [  # (preset, N, K, L, base_log, B, NT) — the kernel instantiations of capi.cu
    (64, 512, 4, 3, 12, 3, 256), (64, 512, 4, 3, 12, 4, 320), (64, 512, 4, 3, 12, 2, 256), (64, 512, 4, 1, 13, 3, 256), (64, 512, 4, 1, 13, 4, 320),
    (64, 512, 4, 1, 13, 1, 256),
    (4, 1024, 2, 2, 15, 2, 256), (256, 1024, 2, 4, 9, 2, 256), (4, 1024, 2, 1, 11, 2, 256), (4, 1024, 2, 1, 11, 1, 96)]


@pytest.mark.parametrize("case", CASES)
def test_cmux_step_matches_oracle(emul, ol, oracle64, case):
    pid, N, K, L, blog, B, NT = case
    G = K + 1
    rng = np.random.default_rng(sum(case))
    if pid == 64 and L == 3:
        o = oracle64
        ggsw = np.ascontiguousarray(o.bsk.reshape(o.p.n, L, G, G, N)[11])          # a real GGSW of a key bit
    else:
        # external_product only needs the FFT plan and (k, N): a key-less oracle over random "GGSW" words
        p = ol.preset(pid)
        o = _plan_only(ol, p)
        ggsw = rng.integers(0, 2**64, (L, G, G, N), dtype=np.uint64)
    acc = rng.integers(0, 2**64, (B, G, N), dtype=np.uint64)
    rot = rng.integers(0, 2 * N, B).astype(np.int32)
    rot[0] = 2 * N - 1
    want = np.stack([o.external_product(ggsw, L, blog, _rot_diff(acc[b], int(rot[b]), N).reshape(-1), acc[b].reshape(-1)).reshape(G, N)
                     for b in range(B)])
    got = acc.copy()
    assert emul.emul_cmux_step(N, K, L, B, NT, ggsw, blog, rot, got) == 0
    diff = np.abs((got - want).astype(np.int64)).astype(np.float64)
    # f64 rounding tolerance: both sides are unnormalised f64 FFTs in different operation order; |x| <= 2^23 ⇒ 2^-30 abs
    assert diff.max() < 2.0**35, np.log2(diff.max())


@pytest.mark.parametrize("L,blog", [(3, 12), (2, 15)])
def test_merged_step_equals_level_by_level_step(emul, ol, oracle64, L, blog):
    """pbs_merged_kernel's schedule (accumulator coefficients in registers, rotation copy aliased onto FFT buffer 0, all levels in
    one barrier interval) performs the same arithmetic in the same order as the level-by-level step: the SAME words, two steps
    in a row (the second consumes the refreshed rotation copy)."""
    N, K, B, NT = 512, 4, 3, 256
    G = K + 1
    rng = np.random.default_rng(100 + L)
    if L == 3:
        ggsws = [np.ascontiguousarray(oracle64.bsk.reshape(oracle64.p.n, L, G, G, N)[i]) for i in (11, 12)]
    else:
        ggsws = [rng.integers(0, 2**64, (L, G, G, N), dtype=np.uint64) for _ in range(2)]
    acc = rng.integers(0, 2**64, (B, G, N), dtype=np.uint64)
    a, b = acc.copy(), acc.copy()
    for step, ggsw in enumerate(ggsws):
        rot = rng.integers(0, 2 * N, B).astype(np.int32)
        rot[step] = (0, 2 * N - 1)[step]
        assert emul.emul_cmux_step(N, K, L, B, NT, ggsw, blog, rot, a) == 0
        assert emul.emul_cmux_step_merged(N, K, L, B, NT, ggsw, blog, rot, b) == 0
        assert np.array_equal(a, b), f"step {step}"
    assert not np.array_equal(a, acc)


_plans = {}


def _plan_only(ol, p):
    """an Oracle handle whose keys are all-zero (cheap) — enough for external_product, which only uses the FFT plan"""
    key = (p.N, p.k)
    if key not in _plans:
        z = lambda n: np.zeros(n, dtype=np.uint64)
        G = p.k + 1
        raw = (z(p.big), z(p.n), z(p.n * p.pbs_l * G * G * p.N), z(p.big * p.ks_l * (p.n + 1)), z(G * (p.big + 1) * p.pfks_l * G * p.N))
        _plans[key] = ol.Oracle(p, seed=0, raw=raw)
    return _plans[key]
