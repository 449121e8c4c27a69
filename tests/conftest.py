import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SEED = 20261018


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tac():
    """the product package (host mirror + C ABI loader)"""
    so = os.path.join(ROOT, "tfhe-aes-2_b200", "csrc", "libtfhe_aes_cuda.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("tfhe-aes-2_b200")


@pytest.fixture(scope="session")
def ol():
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def oracle64(ol):
    return ol.Oracle(64, seed=SEED)


@pytest.fixture(scope="session")
def ck64(tac):
    return tac.ClientKey(64, seed=SEED).gen_eval_keys()


@pytest.fixture(scope="session")
def gpu64(tac, ck64):
    """(ClientKey, FheContext) for params_sqrd_lvl_64 with keys resident on cuda:0 — like the reference's KEYS_SQRD_LVL_64"""
    ctx = tac.FheContext(ck64.params, device=0)
    ctx.upload_keys(ck64)
    ck64.context = ctx
    return ck64, ctx


def sbox_gal_mul_fn(ol):
    S = [ol.sbox(i) for i in range(256)]
    return lambda b: (ol.gf_256_mul(S[b], 1) << 16) | (ol.gf_256_mul(S[b], 2) << 8) | ol.gf_256_mul(S[b], 3)
