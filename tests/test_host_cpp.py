"""The C++ host mirror (tfhe-aes-2_b200/host/): the reference's generic AES code written against the ByteT policy, the model
API with the reference's panic messages, and the CLI binary of src/bin/main.rs — built here, run on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tfhe-aes-2_b200", "host")


@pytest.fixture(scope="module")
def host_bins(tac):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return os.path.join(HOST, "tfhe_aes_cli"), os.path.join(HOST, "host_tests")


def test_cli_usage_and_loud_failure_without_gpu(host_bins):
    cli, _ = host_bins
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([cli, "--key", "00", "--iv", "bdd219b8a08ded1a", "--number-of-outputs", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "invalid key length, must be 16 bytes" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([cli, "--key", "76b8e0ada0f13d90405d6ae55386bd28", "--iv", "bdd219b8a08ded1a", "--number-of-outputs", "1"],
                           capture_output=True, text=True)
        assert r.returncode == 101 and "panicked" in r.stderr and "tac_ctx_create failed" in r.stderr       # no CPU fallback


@pytest.mark.gpu
def test_host_tests_binary(host_bins):
    _, tests = host_bins
    r = subprocess.run([tests], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "all passed" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--generic"]])
def test_cli_reference_config(host_bins, extra):
    """BASELINE config 1: --key 76b8… --iv bdd2… (here 2 outputs; full FHE key expansion + 10 rounds, like the reference binary)"""
    from test_oracle_golden import CLI_OUT
    cli, _ = host_bins
    n = 2 if not extra else 1
    r = subprocess.run([cli, "--key", "76b8e0ada0f13d90405d6ae55386bd28", "--iv", "bdd219b8a08ded1a", "--number-of-outputs", str(n)] + extra,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = r.stdout.strip().splitlines()
    assert lines[-n:] == CLI_OUT[:n]
    assert any(l.startswith("AES key expansion took") for l in lines) and any(l.startswith(f"AES of #{n} outputs computed in") for l in lines)
