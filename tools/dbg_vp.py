"""development helper: vertical-packing phase differences GPU vs oracle (same data as tests/test_gpu_parity.py)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as ol
from conftest import SEED, sbox_gal_mul_fn
tac = importlib.import_module("tfhe-aes-2_b200")
ck = tac.ClientKey(64, seed=SEED).gen_eval_keys()
ctx = tac.FheContext(ck.params, device=0); ctx.upload_keys(ck)
orc = ol.Oracle(64, seed=SEED)
f = sbox_gal_mul_fn(ol)
lut = ctx.generate_lookup_table(8, 24, f)
vals = [0x00, 0x53]
ggsw = np.stack([orc.pfks(orc.pbs(orc.keyswitch(ck.encrypt_bytes([v])[0]))) for v in vals])
got = ctx.stage_vertical_packing(ggsw, lut)
sg = lambda x: np.asarray(x, dtype=np.uint64).astype(np.int64).astype(np.float64)
for i, v in enumerate(vals):
    ref = orc.vertical_packing(ggsw[i], 8, lut.table, 24)
    want = ck.decrypt_bits(ref).astype(np.uint64) << np.uint64(63)
    e_gpu, e_ref = sg(ck.decrypt_phases(got[i]) - want), sg(ck.decrypt_phases(ref) - want)
    print(hex(v), "log2 max|e_gpu|", np.log2(np.abs(e_gpu).max()), "log2 max|e_ref|", np.log2(np.abs(e_ref).max()), "log2 max|diff|", np.log2(np.abs(e_gpu - e_ref).max() + 1))
    print("   per-output log2|diff|:", np.round(np.log2(np.abs(e_gpu - e_ref) + 1), 1).tolist())
