"""Throughput of Lindell-2017 two-party signing through the C ABI (host buffers): party two's `PartialSig::compute` (one job modulo N^2 per
signature) and party one's `Signature::compute_with_recid` (one CRT decrypt per signature), then `verify`, for `count` signatures
over the 24 Paillier keys of the fixture key sets.  Every signature is checked on the device with the reference's `verify`.
    python tools/l17_throughput.py [count]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
from tests.golden import fixtures
pkg = entry.load_package()
from mpecdsa_b200 import gg20, lindell17
from oracle.gg20_oracle import Q
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = pkg.Engine(0)
lindell17._bind(eng.lib); pkg._bind_l01(eng.lib)
keysets = fixtures.load_all_keysets()
ks = gg20.KeySets(eng, keysets)
rows = [lk for s in keysets for lk in s]
n_list = [lk.dk.p * lk.dk.q for lk in rows]
rng = np.random.default_rng(0x117)
def scalars(n):
    a = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint32); a[:, 7] &= 0x7FFFFFFF; a[:, 0] |= 1
    return a
P = lambda a: a.ctypes.data
key_idx = (np.arange(count) % len(rows)).astype(np.uint32)
N = pkg.ints_to_limbs(n_list, 64)
x1, x2, k1, k2, msg = (scalars(count) for _ in range(5))
rho = rng.integers(0, 2**32, size=(count, 16), dtype=np.uint32); rho[:, 15] &= 0x3FFFFFFF
r_enc = rng.integers(0, 2**32, size=(count, 64), dtype=np.uint32); r_enc[:, 63] &= 0x3FFFFFFF; r_enc[:, 0] |= 1
r_key = rng.integers(0, 2**32, size=(count, 64), dtype=np.uint32); r_key[:, 63] &= 0x3FFFFFFF; r_key[:, 0] |= 1
x1w = np.zeros((count, 64), np.uint32); x1w[:, :8] = x1
c_key = np.zeros((count, 128), np.uint32)
eng._ck(eng.lib.tecdsa_paillier_encrypt_batch(eng._ctx, P(N), P(key_idx), len(rows), P(x1w), P(r_key), P(c_key), count, pkg.HOST), "encrypt")
R1, R2, pub = (np.zeros((count, 16), np.uint32) for _ in range(3))
eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, None, P(k1), P(R1), count, pkg.HOST), "R1")
eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, None, P(k2), P(R2), count, pkg.HOST), "R2")
x12 = np.zeros((count, 8), np.uint32)
eng._ck(eng.lib.tecdsa_secp_scalar_mul_batch(eng._ctx, P(x1), P(x2), P(x12), count, pkg.HOST), "x1 x2")
eng._ck(eng.lib.tecdsa_secp_mul_batch(eng._ctx, None, P(x12), P(pub), count, pkg.HOST), "pub")
c3 = np.zeros((count, 128), np.uint32)
st = np.full(count, 255, np.uint8)
sr, ss = np.zeros((count, 8), np.uint32), np.zeros((count, 8), np.uint32)
rec = np.zeros(count, np.uint8)
res = {"count": count, "keys": len(rows)}
for rep in range(2):
    t0 = time.perf_counter()
    eng._ck(eng.lib.tecdsa_l17_partial_sig_batch(eng._ctx, P(N), P(key_idx), len(rows), P(c_key), P(x2), P(k2), P(R1), P(msg), P(rho), P(r_enc), P(c3), P(st), count, pkg.HOST), "partial_sig")
    t1 = time.perf_counter()
    assert not st.any()
    eng._ck(eng.lib.tecdsa_l17_sign_batch(eng._ctx, ks.handle, P(key_idx), P(c3), P(k1), P(R2), P(sr), P(ss), P(rec), P(st), count, pkg.HOST), "sign")
    t2 = time.perf_counter()
    assert not st.any()
    eng._ck(eng.lib.tecdsa_l17_verify_batch(eng._ctx, P(sr), P(ss), P(pub), P(msg), P(st), count, pkg.HOST), "verify")
    t3 = time.perf_counter()
    res = {"count": count, "keys": len(rows), "partial_sig_per_s": count / (t1 - t0), "sign_per_s": count / (t2 - t1), "verify_per_s": count / (t3 - t2),
           "signatures_per_s_both_parties": count / (t2 - t0), "all_signatures_verify": bool(not st.any())}
print(json.dumps(res))
