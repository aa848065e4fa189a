"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library at sizes
that finish under the tool — an offline batch over 8 key sets through the records call, a SPLIT batch (two host threads, two
streams), the online step, modexp with both squaring paths, modinv, the Scalar/Point surface, keygen prove+verify, Paillier
open + ECDDH (blame), fixed-base table build, N-adic setup.
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py [split]
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
from tests.golden import fixtures
pkg = entry.load_package()
from mpecdsa_b200 import blame, gg20, keygen
eng = pkg.Engine(0)
keysets = fixtures.load_all_keysets()
ks = gg20.KeySets(eng, keysets)
n = 2100 if "split" in sys.argv else 6                  # >= 2048 sessions take the two-stream / two-thread path
sess, rnd = gg20.synthetic_batch(keysets, n, 7)
rec = np.zeros((1, 2 * n, 256), np.uint8)
eng.offline_records(ks, None, sess, n, rnd, rec, pkg.HOST)
print("offline records ok:", not rec[0, :, 0].any(), flush=True)
if "split" in sys.argv:
    ks.free(); eng.close(); print("done", flush=True); sys.exit(0)
res = gg20.offline_batch(eng, ks, sess, rnd)
msg = np.tile(np.arange(8, dtype=np.uint32) + 1, (n, 1))
sig = gg20.sign_batch(eng, ks, sess, msg, res.R, res.sigma, np.ascontiguousarray(rnd[:, 8:16]))
print("sign", sig["status"].tolist(), flush=True)
r = random.Random(2)
m = [r.getrandbits(2048) | 1 | (1 << 2047) for _ in range(5)]
for sqr in (0, 1):
    eng.set_option("sqr", sqr)
    print("modexp sqr=%d" % sqr, eng.mod_pow([r.getrandbits(2048) for _ in m], [r.getrandbits(300) for _ in m], m)[1].tolist(), flush=True)
eng.set_option("sqr", 0)
print("modinv", [x is not None for x in eng.mod_inv([r.getrandbits(2048) for _ in m], m)], flush=True)
P = eng.secp_mul(None, [5, 7])
print("secp", eng.secp_mul(P, [11, 13])[0] is not None, eng.point_add(P, P[::-1])[0] is not None, len(eng.point_compress(P)), eng.point_decompress(eng.point_compress(P)) == P, flush=True)
print("scalar", eng.scalar_op("inv", [5, 0]), eng.unit_mod_check([5, m[1]], m[:2]), eng.sha256([b"abc"])[0][:2].hex(), flush=True)
k0 = keysets[0][0]
sigv, st = keygen.correct_key_prove(eng, [(k0.dk.p, k0.dk.q)])
print("correct key", st.tolist(), keygen.correct_key_verify(eng, [k0.dk.p * k0.dk.q], sigv).tolist(), flush=True)
sh, cm = keygen.vss_share(eng, 1, 3, [[5, 7]])
print("vss", keygen.vss_validate_share(eng, [cm[0]] * 3, sh[0], [1, 2, 3]).tolist(), flush=True)
c = eng.paillier_encrypt([k0.dk.p * k0.dk.q], [0], [12345], [987654321])
print("open", blame.paillier_open(eng, ks, [0], c), flush=True)
pf = blame.ecddh_prove(eng, [9], [P[0]], eng.secp_mul([P[0]], [9]), [P[1]], eng.secp_mul([P[1]], [9]), [77])
print("ecddh", blame.ecddh_verify(eng, pf, [P[0]], eng.secp_mul([P[0]], [9]), [P[1]], eng.secp_mul([P[1]], [9])).tolist(), flush=True)
ks.free(); eng.close()
print("done", flush=True)
