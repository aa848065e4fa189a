"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family of the library at sizes
that finish under the tool — an offline batch over 8 key sets through the records call, a SPLIT batch (two host threads, two
streams), the online step, modexp with both squaring paths, modinv, the Scalar/Point surface, keygen prove+verify, Paillier
open + ECDDH (blame), Lindell-2017 signing, the interactive PDL proof, GG18 whole signing (three signers), the size-generic GG20 driver
(three signers), fixed-base table build, N-adic setup.
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
# ---- section 8(f) rank 4: Lindell-2017, zk_pdl, GG18 (whole signing), and the size-generic GG20 driver on three signers
from mpecdsa_b200 import gg18, gg20_general, lindell17
from oracle import gg20_oracle as o
Q = o.Q
n0 = k0.dk.p * k0.dk.q
x1, x2, k1, k2 = 11, 13, 17, 19
c_key = eng.paillier_encrypt([n0], [0], [x1], [5])
e1 = lindell17.eph_create(eng, [k1], [23])
e2 = lindell17.eph_create(eng, [k2], [29], [31], [37])
print("l17 eph", lindell17.eph_verify(eng, e1["public_share"], e1["c"], e1["proof"]).tolist(),
      lindell17.eph_verify(eng, e2["public_share"], e2["c"], e2["proof"], [31], [37], e2["pk_commitment"], e2["zk_pok_commitment"]).tolist(), flush=True)
c3, st = lindell17.p2_partial_sig(eng, [n0], [0], c_key, [x2], [k2], e1["public_share"], [1234], [3 * Q + 1], [7])
sr, ss, rec, st2 = lindell17.p1_sign(eng, ks, [0], c3, [k1], e2["public_share"])
print("l17 sign", st.tolist(), st2.tolist(), lindell17.verify(eng, sr, ss, eng.secp_mul(None, [x1 * x2]), [1234]).tolist(), flush=True)
ct, ctt, qt, st = lindell17.pdl_verifier_message1(eng, [n0], [0], c_key, eng.secp_mul(None, [x1]), [41], [43 * Q + 3], [9], [47])
ch, qh, al, st2 = lindell17.pdl_prover_message1(eng, ks, [0], ct, [53])
print("zk_pdl", st.tolist(), st2.tolist(), lindell17.pdl_prover_message2(eng, [x1], al, ctt, [41], [43 * Q + 3], [47]).tolist(),
      lindell17.pdl_verifier_finalize(eng, ch, qh, [53], qt).tolist(), flush=True)
rr = random.Random(3)
parties, rows = 3, [0, 1, 2]
U, P1 = 3, 2
w = [o.lagrange_at_zero(p_, rows) * keysets[0][p_].x_i % Q for p_ in rows]
sc = lambda m_: [rr.randrange(1, Q) for _ in range(m_)]
nm = lambda elems: [rr.randrange(1, keysets[0][rows[u]].dk.p * keysets[0][rows[u]].dk.q >> 1) for u in elems]
alice = [u for u in range(U) for _ in range(P1)]
rnd18 = dict(k=sc(U), gamma=sc(U), blind=sc(U), r_a=nm(range(U)), l=sc(U), rho=sc(U), blind5=sc(U), blind5c=sc(U), heg_s1=sc(U), heg_s2=sc(U), dlog_nonce=sc(U),
             r_b_gamma=nm(alice), r_b_w=nm(alice), nb_gamma=sc(U * P1), nbt_gamma=sc(U * P1), nb_w=sc(U * P1), nbt_w=sc(U * P1),
             beta_tag_gamma=nm(alice), beta_tag_w=nm(alice))
out = gg18.sign_batch(eng, ks, parties, rows, w, [keysets[0][0].y_sum_s] * U, [99] * U, rnd18)
print("gg18 sign", out["status"].tolist(), flush=True)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gg20_general as tg
keys3, rnd3 = tg._session(rr, keysets[0], [1, 2, 3])
args = tg._flatten([(keys3, [1, 2, 3], rnd3)], lambda lk, j: j)
print("gg20 general", gg20_general.offline_batch(eng, ks, *args)["status"].tolist(), flush=True)
ks.free(); eng.close()
print("done", flush=True)
