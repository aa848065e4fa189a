"""Generates tests/golden/vectors_r01.json — known-answer vectors for the hot path.

The reference ships none; these are produced by the oracle (oracle/gg20_oracle.py, itself pinned
to GMP / OpenSSL / hashlib in tests/test_oracle.py) from fixed seeds over the committed key fixtures, so that (a) the
oracle cannot drift silently and (b) the CUDA path is checked against frozen bytes, not only against a live oracle.
When a build of the real Rust reference becomes available, running it on these inputs is what pins the [R] encodings.
    python -m tests.golden.make_vectors
"""
import json
import os
import random

from oracle import gg20_oracle as o
from oracle.sampling import Drbg, sample_unit
from tests.golden import fixtures

HERE = os.path.dirname(os.path.abspath(__file__))
H = hex


def build():
    keyset = fixtures.load_keyset(0)
    v = {"about": "oracle-generated known answers, see make_vectors.py", "modexp": [], "paillier": [], "alice": [], "pdl": [], "bob": [],
         "sigma": {}, "offline": []}
    rnd = random.Random(0xB2000001)
    for bits in (1024, 2048, 4096):
        for _ in range(3):
            m = rnd.getrandbits(bits) | 1 | (1 << (bits - 1))
            b, e = rnd.getrandbits(bits) % m, rnd.getrandbits(bits)
            v["modexp"].append({"bits": bits, "base": H(b), "exp": H(e), "mod": H(m), "out": H(pow(b, e, m))})
    rng = Drbg(0xB2000001, "vectors")
    q3 = o.Q ** 3
    for row in range(3):
        lk = keyset[row]
        ek = lk.paillier_key_vec[row]
        m, r, k = rng.below(ek.n), rng.unit_mod(ek.n), rng.scalar()
        c = o.paillier_encrypt(ek, m, r)
        ck = o.paillier_mul(ek, c, k)
        v["paillier"].append({"row": row, "m": H(m), "r": H(r), "k": H(k), "c": H(c), "c_mul_k": H(ck), "c_add": H(o.paillier_add(ek, c, ck)),
                              "dec_c_add": H(o.paillier_decrypt(lk.dk, o.paillier_add(ek, c, ck)))})
        st_row = (row + 1) % 3
        st = lk.h1_h2_n_tilde_vec[st_row]
        a, ra = rng.scalar(), rng.unit_mod(ek.n)
        ca = o.paillier_encrypt(ek, a, ra)
        rr = (rng.below(q3), rng.unit_mod(ek.n), rng.below(q3 * st.N), rng.below(o.Q * st.N))
        pf = o.alice_proof_generate(a, ca, ek, st, ra, *rr)
        assert o.alice_proof_verify(pf, ca, ek, st)
        v["alice"].append({"ek_row": row, "st_row": st_row, "a": H(a), "r": H(ra), "cipher": H(ca), "rand": [H(x) for x in rr],
                           "z": H(pf.z), "e": H(pf.e), "s": H(pf.s), "s1": H(pf.s1), "s2": H(pf.s2)})
        x, rx = rng.scalar(), rng.unit_mod(ek.n)
        cx = o.paillier_encrypt(ek, x, rx)
        Gp = o.pt_mul(o.G, rng.scalar()); Qp = o.pt_mul(Gp, x)
        pr = (rng.below(q3), 1 + rng.below(ek.n - 2), rng.below(o.Q * st.N), rng.below(q3 * st.N))
        pp = o.pdl_prove(x, rx, cx, ek, Qp, Gp, st.g, st.ni, st.N, *pr)
        assert o.pdl_verify(pp, cx, ek, Qp, Gp, st.g, st.ni, st.N)
        v["pdl"].append({"ek_row": row, "st_row": st_row, "x": H(x), "r": H(rx), "cipher": H(cx), "G": [H(Gp[0]), H(Gp[1])], "Q": [H(Qp[0]), H(Qp[1])],
                         "rand": [H(t) for t in pr], "z": H(pp.z), "u1": [H(pp.u1[0]), H(pp.u1[1])], "u2": H(pp.u2), "u3": H(pp.u3),
                         "s1": H(pp.s1), "s2": H(pp.s2), "s3": H(pp.s3)})
        b_, bp_, rb = rng.scalar(), rng.below(ek.n), rng.unit_mod(ek.n)
        mta = o.paillier_add(ek, o.paillier_mul(ek, ca, b_), o.paillier_encrypt(ek, bp_, rb))
        br = (rng.below(q3), rng.unit_mod(ek.n), rng.below(o.Q ** 2 * ek.n), rng.below(o.Q * st.N), rng.below(q3 * st.N), rng.below(o.Q * st.N), rng.below(q3 * st.N))
        bpf, u = o.bob_proof_generate(ca, mta, b_, bp_, ek, st, rb, True, *br)
        assert o.bob_proof_ext_verify(bpf, u, ca, mta, ek, st, o.pt_mul(o.G, b_))
        v["bob"].append({"ek_row": row, "st_row": st_row, "a_enc": H(ca), "mta": H(mta), "b": H(b_), "beta_prim": H(bp_), "r": H(rb), "rand": [H(t) for t in br],
                         "t": H(bpf.t), "z": H(bpf.z), "e": H(bpf.e), "s": H(bpf.s), "s1": H(bpf.s1), "s2": H(bpf.s2), "t1": H(bpf.t1), "t2": H(bpf.t2),
                         "u": [H(u[0]), H(u[1])]})
    sk, nonce = rng.scalar(), rng.scalar()
    d = o.dlog_prove(sk, nonce)
    v["sigma"]["dlog"] = {"sk": H(sk), "nonce": H(nonce), "pk": [H(d.pk[0]), H(d.pk[1])], "T": [H(d.pk_t_rand_commitment[0]), H(d.pk_t_rand_commitment[1])],
                          "response": H(d.challenge_response)}
    m, r, s1, s2 = (rng.scalar() for _ in range(4))
    pd = o.pedersen_prove(m, r, s1, s2)
    v["sigma"]["pedersen"] = {"m": H(m), "r": H(r), "s1": H(s1), "s2": H(s2), "com": [H(pd.com[0]), H(pd.com[1])], "e": H(pd.e), "z1": H(pd.z1), "z2": H(pd.z2)}
    for s_l in ([1, 2], [1, 3], [2, 3], [3, 1]):
        keys = [keyset[i - 1] for i in s_l]
        drbg = Drbg(0xB2000005, f"vector-session{s_l}")
        rnds = [sample_unit(drbg, keys, s_l, p) for p in range(2)]
        res = o.offline_session(keys, s_l, rnds)
        assert [x.status for x in res] == [0, 0]
        v["offline"].append({"s_l": s_l, "seed": "Drbg(0xB2000005, 'vector-session%s')" % s_l, "R": [H(res[0].R[0]), H(res[0].R[1])],
                             "sigma": [H(x.sigma_i) for x in res], "digest": [x.transcript.hex() for x in res]})
    return v


if __name__ == "__main__":
    with open(os.path.join(HERE, "vectors_r01.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote vectors_r01.json")
