"""Frozen outputs of the section-8(f) rank-4 oracles (oracle/lindell17_oracle.py, oracle/gg18_oracle.py, oracle/gg20_general_oracle.py)
on seeded inputs: tests/golden/vectors_other_protocols.json.  They pin the ORACLE against accidental drift (the oracle is our
restatement: parity with the reference stays unpinned).  The inputs are regenerated from the seeds by the
same helpers the tests use (tests/test_other_protocols.py, tests/test_gg20_general.py); only outputs are stored.
    python -m tests.golden.make_other_vectors"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import gg18_oracle as e18                                     # noqa: E402
from oracle import gg20_general_oracle as gen                             # noqa: E402
from oracle import lindell17_oracle as l17                                # noqa: E402
from oracle import gg20_oracle as o                                       # noqa: E402
from tests.golden import fixtures                                         # noqa: E402
import test_gg20_general as tg                                            # noqa: E402
import test_other_protocols as tp                                         # noqa: E402

PATH = os.path.join(HERE, "vectors_other_protocols.json")
hx = lambda v: format(v, "x")
pt = lambda p: [hx(p[0]), hx(p[1])]


def compute():
    keyset = fixtures.load_keyset()
    out = {}
    # Lindell-2017 + zk_pdl
    rng = random.Random(0x117601)
    c = tp._l17_case(keyset, rng, 2)
    rows = []
    for i in range(2):
        e1 = l17.eph_create(c["k1"][i], c["n1"][i])
        e2 = l17.eph_create(c["k2"][i], c["n2"][i], c["b1"][i], c["b2"][i])
        c3 = l17.p2_partial_sig(c["eks"][i], c["c_key"][i], c["x2"][i], c["k2"][i], e1.public_share, c["msg"][i], c["rho"][i], c["r_enc"][i])
        r, s, rec = l17.p1_sign(c["dks"][i], c3, c["k1"][i], e2.public_share)
        a, b = rng.randrange(1, o.Q), rng.randrange(o.Q * o.Q)
        st = l17.pdl_verifier_message1(c["eks"][i], c["c_key"][i], o.pt_mul(o.G, c["x1"][i]), a, b, c["r_enc"][i], c["b1"][i] % o.Q)
        c_hat, q_hat, alpha = l17.pdl_prover_message1(c["dks"][i], st.c_tag, c["b2"][i] % o.Q)
        first = l17.p1_keygen_first(c["x1"][i], c["n1"][i], c["b1"][i], c["b2"][i])
        rows.append({"pk_commitment": hx(e2.pk_commitment), "zk_pok_commitment": hx(e2.zk_pok_commitment), "ecddh_z": [hx(e1.proof.z), hx(e2.proof.z)],
                     "c3_sha256": hashlib.sha256(o.bn_bytes(c3)).hexdigest(), "sig": [hx(r), hx(s), rec],
                     "c_tag_sha256": hashlib.sha256(o.bn_bytes(st.c_tag)).hexdigest(), "c_tag_tag": hx(st.c_tag_tag), "q_tag": pt(st.q_tag),
                     "c_hat": hx(c_hat), "alpha": hx(alpha), "keygen_commitments": [hx(first.pk_commitment), hx(first.zk_pok_commitment)]})
    out["lindell17_seed_0x117601"] = rows
    # GG18 phases 5a-5d
    rng = random.Random(0x18601)
    g = tp._gg18_case(rng, 1, 3)
    a5, c5, d5 = tp._gg18_oracle_run(g)
    out["gg18_seed_0x18601"] = {"com": [hx(x.com) for x in a5], "V": [pt(x.V) for x in a5], "heg_z1": [hx(x.heg.z1) for x in a5],
                                "dlog_response": [hx(x.dlog.challenge_response) for x in a5], "com2": [hx(x[1][0]) for x in c5],
                                "u": [pt(x[1][1]) for x in c5], "t": [pt(x[1][2]) for x in c5], "phase5d": d5,
                                "signature": [hx(v) if isinstance(v, int) and v > 1 else v for v in e18.output_signature(g["R"][0], g["y"][0], g["msg"][0], g["s"])[1]]}
    # size-generic GG20 offline stage, three signers
    rng = random.Random(0x6E601)
    keys, rnd = tg._session(rng, keyset, [2, 3, 1])
    res = gen.offline_session(keys, [2, 3, 1], rnd)
    out["gg20_general_seed_0x6E601_signers_2_3_1"] = [{"status": x.status, "R": pt(x.R), "sigma_i": hx(x.sigma_i), "T_own": pt(x.t_vec[p])} for p, x in enumerate(res)]
    return out


if __name__ == "__main__":
    with open(PATH, "w") as f:
        json.dump(compute(), f, indent=1)
    print("wrote", PATH)
