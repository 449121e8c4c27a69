#!/usr/bin/env python
"""The reference's CLI scenario (src/bin/main.rs:41-159) on the B200 path:

    python tools/aes_ctr.py --key 76b8e0ada0f13d90405d6ae55386bd28 --iv bdd219b8a08ded1a --number-of-outputs 10

client: FHE-encrypt the AES key and the counter blocks iv ‖ BE64(ctr), ctr = 1..N  →  server: FHE key expansion, FHE AES of all
blocks (two timers, like main.rs:137 and :153-157)  →  client: decrypt  →  assert equality with clear AES.
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--number-of-outputs", type=int, required=True)
    ap.add_argument("--iv", required=True)
    ap.add_argument("--key", required=True)
    ap.add_argument("--implementation", default="cuda-woppbs-1bit", choices=["cuda-woppbs-1bit"])
    ap.add_argument("--seed", type=int, default=None, help="TEST ONLY: reproducible (publicly computable) keys; default = OS entropy like the reference")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    print(f"using implementation: {args.implementation}")
    key = bytes.fromhex(args.key)
    iv = bytes.fromhex(args.iv)
    if len(key) != 16:
        raise SystemExit("invalid key length, must be 16 bytes")
    if len(iv) != 8:
        raise SystemExit("invalid iv length, must be 8 bytes")

    tac = importlib.import_module("tfhe-aes-2_b200")
    client_key, ctx = tac.FheContext.generate_keys(64, seed=args.seed, device=args.device)        # generate_keys_sqrd_lvl_64
    # client side: FHE encrypt AES key and blocks (main.rs:107-116)
    key_ct = client_key.encrypt_bytes(key)
    blocks_clear = [iv + ctr.to_bytes(8, "big") for ctr in range(1, args.number_of_outputs + 1)]
    blocks = np.stack([client_key.encrypt_bytes(b) for b in blocks_clear]) if blocks_clear else np.zeros((0, 16, 8, client_key.params.big_lwe_size), np.uint64)
    # server side
    t0 = time.perf_counter()
    ctx.aes_key_schedule(key_ct)                                                                   # leaves the expanded key on the device
    print(f"AES key expansion took: {time.perf_counter() - t0:.3f}s")
    t0 = time.perf_counter()
    enc = ctx.aes_encrypt_blocks(blocks)
    print(f"AES of #{len(blocks_clear)} outputs computed in: {time.perf_counter() - t0:.3f}s")
    # client side: decrypt and compare with clear AES (main.rs:123-127)
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    for i, b in enumerate(blocks_clear):
        got = client_key.decrypt_bytes(enc[i])
        want = Cipher(algorithms.AES(key), modes.ECB()).encryptor().update(b)
        assert got == want, (i, got.hex(), want.hex())
        print(got.hex())


if __name__ == "__main__":
    main()
