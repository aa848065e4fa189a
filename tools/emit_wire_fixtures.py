"""Writes bindings/rust/tests/data/*.json: proofs, messages and a LocalKey PRODUCED BY THE GPU ENGINE, serialised in the
reference's serde-JSON format (multi-party-ecdsa_b200/wire.py), for bindings/rust/tests/reference_accepts_gpu_proofs.rs — the
test a maintainer with a Rust toolchain runs to pin the parity of every [R] encoding at once: the reference deserialises the
documents and its OWN `verify` functions must accept them.

Needs a GPU (run it through gpurun and copy gpurun_out/wire/ to bindings/rust/tests/data/):
    python tools/emit_wire_fixtures.py gpurun_out/wire
Inputs are fixed (committed key fixtures + a seeded sampler), so the documents are reproducible byte for byte.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_documents(eng, keyset, seed: int = 0xB2F2):
    """-> {name: document}; every proof / message below comes out of the engine's C ABI"""
    import hashlib
    import __graft_entry__ as entry
    entry.load_package()
    import numpy as np
    from mpecdsa_b200 import gg20, keygen, wire
    from oracle.sampling import Drbg, sample_unit          # seeded sampling of the values the reference draws at random (test infrastructure)
    from oracle.gg20_oracle import Q as o_Q
    E = wire.DEFAULT
    ks = gg20.KeySets(eng, [keyset])
    lk_a, lk_b = keyset[0], keyset[2]                       # Alice = party 1, Bob = party 3 (s_l = [1, 3])
    rng = Drbg(seed, "wire")
    r = [sample_unit(rng, [lk_a, lk_b], [1, 3], p) for p in range(2)]
    n_a = lk_a.paillier_key_vec[0].n
    stmts = [(s.N, s.g, s.ni) for s in lk_a.h1_h2_n_tilde_vec]
    docs = {}
    # MessageA::a (mta/mod.rs:52-87): Alice encrypts k_i and proves its range against every statement
    c, proofs = gg20.mta_message_a(eng, ks, [0], [[0, 1, 2]], [r[0].k_i], [r[0].r_k], [r[0].alice])
    pf = [{k: proofs[k][0][x] for k in ("z", "e", "s", "s1", "s2")} for x in range(3)]
    docs["message_a"] = {"message": wire.message_a(c[0], pf), "ek": wire.encryption_key(n_a), "dlog_statements": [wire.dlog_statement(*s) for s in stmts],
                         "expect": "every range_proofs[x].verify(&message.c, &ek, &dlog_statements[x]) == true (range_proofs.rs:105)"}
    docs["alice_proof"] = {"cipher": E.bigint(c[0]), "ek": wire.encryption_key(n_a), "dlog_statement": wire.dlog_statement(*stmts[2]), "proof": wire.alice_proof(pf[2]),
                           "expect": "proof.verify(&cipher, &ek, &dlog_statement) == true"}
    # MessageB::b (mta/mod.rs:91-158): Bob answers with b = gamma_i
    b = r[1].gamma_i
    c_b, bp, btp, beta, st = gg20.mta_message_b(eng, ks, [0], [[0, 1, 2]], [b], c, proofs, [r[1].r_gamma], [r[1].beta_tag_gamma],
                                                [r[1].nonce_gamma_b], [r[1].nonce_gamma_beta])
    assert list(st) == [0]
    alpha, _, st = gg20.mta_get_alpha(eng, ks, [0], [r[0].k_i], c_b, bp, btp)
    assert list(st) == [0]
    ab = eng.scalar_op("add", alpha, beta)[0]
    docs["message_b"] = {"message": wire.message_b(c_b[0], bp[0], btp[0]), "dk": {"p": E.bigint(lk_a.dk.p), "q": E.bigint(lk_a.dk.q)},
                         "a": E.scalar(r[0].k_i), "beta": E.scalar(beta[0]), "expected_alpha": E.scalar(alpha[0]), "expected_alpha_plus_beta": E.scalar(ab),
                         "expect": "message.verify_proofs_get_alpha(&dk, &a) == Ok((expected_alpha, _)); alpha + beta == a * b (mta/mod.rs:160, mta/test.rs:12-18)"}
    docs["dlog_proof"] = {"proof": wire.dlog_proof(bp[0]), "expect": "DLogProof::verify(&proof).is_ok()"}
    # PDLwSlackProof (zk_pdl_with_slack/mod.rs:68-179): statement (c, ek_a, Q = x*G', G', (h1, h2, N~) of Bob)
    Gp = eng.secp_mul(None, [r[1].l])[0]                     # an arbitrary second generator, as R is in the protocol
    Qp = eng.secp_mul([Gp], [r[0].k_i])[0]
    al, be, rh, ga = r[0].pdl
    stb = lk_a.h1_h2_n_tilde_vec[2]
    pd = gg20.pdl_prove(eng, ks, [0], [2], [r[0].k_i], [r[0].r_k], c, [Qp], [Gp], [al], [be], [rh], [ga])
    one = {k: pd[k][0] for k in pd}
    assert list(gg20.pdl_verify(eng, ks, [0], [2], c, [Qp], [Gp], *[[one[k]] for k in ("z", "u1", "u2", "u3", "s1", "s2", "s3")])) == [0]
    docs["pdl"] = {"statement": wire.pdl_statement(c[0], n_a, Qp, Gp, stb.g, stb.ni, stb.N), "proof": wire.pdl_proof(one),
                   "expect": "PDLwSlackProof::verify(&proof, &statement).is_ok() (zk_pdl_with_slack/mod.rs:127)"}
    # LocalKey (keygen/rounds.rs:309-322)
    comm = [lk_a.y_sum_s, eng.point_add([lk_a.pk_vec[1]], [lk_a.pk_vec[0]], subtract=True)[0]]     # f(0)*G and a_1*G = X_2 - X_1 for the degree-1 fixture polynomial
    docs["local_key"] = wire.local_key(lk_a, comm)
    # KeyGenBroadcastMessage1 (party_i.rs:219-258) for party 1's Paillier key with a fresh (N~, h1, h2)
    from tests.test_keygen_oracle import _setup
    import random
    setup = _setup(random.Random(seed), bits=1024)
    params, st = keygen.h1_h2_n_tilde(eng, [setup])
    nt, h1, h2, xn, xin = params[0]
    sig, _ = keygen.correct_key_prove(eng, [(lk_a.dk.p, lk_a.dk.q)])
    cd = keygen.composite_dlog_prove(eng, [(nt, h1, h2), (nt, h2, h1)], [xn, xin], [rng.bits(512), rng.bits(512)])
    blind = rng.bits(256)
    y_i = eng.secp_mul(None, [r[0].gamma_i])[0]
    com = gg20.hash_commitment(eng, [y_i], [blind])[0]
    docs["keygen_broadcast1"] = {"message": wire.keygen_broadcast1(n_a, (nt, h1, h2), com, sig[0], cd[0], cd[1]),
                                 "decommit": {"blind_factor": E.bigint(blind), "y_i": E.point(y_i)},
                                 "expect": "correct_key_proof.verify(&e, SALT_STRING).is_ok(); both composite dlog proofs verify; the commitment opens (party_i.rs:272-305)"}
    # a full offline session + online step: SignatureRecid verifies under the reference's `verify` (party_i.rs:913-936)
    rnd = gg20.pack_randomness(r)
    res = gg20.offline_batch(eng, ks, [(0, 0, 2)], rnd)
    assert not res.status.any()
    m = int.from_bytes(hashlib.sha256(b"ZenGo").digest(), "big")
    msg = np.frombuffer(m.to_bytes(32, "little"), dtype="<u4").reshape(1, 8).copy()
    kk = np.ascontiguousarray(rnd[:, 8:16])
    sg = gg20.sign_batch(eng, ks, np.array([[0, 0, 2]], dtype=np.uint32), msg, res.R, res.sigma, kk)
    assert list(sg["status"]) == [0]
    I = lambda row: int.from_bytes(np.ascontiguousarray(row).tobytes(), "little")
    I2pt = lambda row: (I(row[:8]), I(row[8:16]))
    unpack_res_point = lambda row: I2pt(row)
    docs["signature"] = {"sig": wire.signature_recid(I(sg["r"][0]), I(sg["s"][0]), int(sg["recid"][0])), "y": E.point(lk_a.y_sum_s), "message": E.bigint(m),
                         "expect": "verify(&sig, &y, &message).is_ok() (gg_2020/party_i.rs:913); message = Sha256(b\"ZenGo\") as in sign.rs:693-696"}
    # Lindell-2017 (two_party_ecdsa/lindell_2017): party one's ephemeral first message (ECDDHProof) and a whole signature; the
    # reference recomputes the signature from party two's GPU-made c3 with ITS Paillier decrypt and must get the engine's (r, s)
    from mpecdsa_b200 import lindell17
    x1, x2, k1, k2 = rng.below(o_Q // 3), rng.below(o_Q), rng.below(o_Q - 1) + 1, rng.below(o_Q - 1) + 1
    r_key, r_enc, rho = rng.below(n_a - 1) + 1, rng.below(n_a - 1) + 1, rng.below(o_Q * o_Q)
    c_key = eng.paillier_encrypt([n_a], [0], [x1], [r_key])
    e1 = lindell17.eph_create(eng, [k1], [rng.below(o_Q - 1) + 1])
    e2 = lindell17.eph_create(eng, [k2], [rng.below(o_Q - 1) + 1], [rng.bits(256)], [rng.bits(256)])
    assert list(lindell17.eph_verify(eng, e1["public_share"], e1["c"], e1["proof"])) == [0]
    c3, st = lindell17.p2_partial_sig(eng, [n_a], [0], c_key, [x2], [k2], e1["public_share"], [m], [rho], [r_enc])
    assert list(st) == [0]
    sr, ss, rec, st = lindell17.p1_sign(eng, ks, [0], c3, [k1], e2["public_share"])
    pub = eng.secp_mul(None, [x1 * x2 % o_Q])[0]
    assert list(st) == [0] and list(lindell17.verify(eng, sr, ss, [pub], [m])) == [0]
    pf = e1["proof"][0]
    docs["lindell17_eph_first_message"] = {"message": {"d_log_proof": {"a1": E.point(I2pt(pf[:16])), "a2": E.point(I2pt(pf[16:32])), "z": E.scalar(I(pf[32:40]))},
                                                       "public_share": E.point(e1["public_share"][0]), "c": E.point(e1["c"][0])},
                                           "expect": "ECDDHProof::verify over (G, public_share, base_point2, c) is Ok (party_two.rs:374-387)"}
    docs["lindell17_signature"] = {"party_one_ec_key": {"public_share": E.point(eng.secp_mul(None, [x1])[0]), "secret_share": E.scalar(x1)},
                                   "party_one_paillier": {"ek": wire.encryption_key(n_a), "dk": {"p": E.bigint(lk_a.dk.p), "q": E.bigint(lk_a.dk.q)},
                                                          "encrypted_share": E.bigint(c_key[0]), "randomness": E.bigint(r_key)},
                                   "party_one_eph": {"public_share": E.point(e1["public_share"][0]), "secret_share": E.scalar(k1)},
                                   "party_two_eph_public": E.point(e2["public_share"][0]), "c3": E.bigint(c3[0]), "pubkey": E.point(pub), "message": E.bigint(m),
                                   "signature": {"s": E.bigint(ss[0]), "r": E.bigint(sr[0])}, "recid": int(rec[0]),
                                   "expect": "party_one::Signature::compute(..) == signature and party_one::verify(&signature, &pubkey, &message).is_ok() "
                                             "(party_one.rs:486-517, 567-592)"}
    # the six round messages of one signer of a two-signer offline stage, as `Msg<OfflineProtocolMessage>` (state_machine/sign.rs:478-490)
    from mpecdsa_b200 import gg20_general
    from oracle.gg20_oracle import lagrange_at_zero, pt_mul
    s_l = [1, 3]
    keys2 = [lk_a, lk_b]
    g_rnd = {f: [] for f in ("gamma", "k", "blind", "r_k", "l", "ped_s1", "ped_s2", "heg_s1", "heg_s2", "alice", "beta_tag_gamma", "r_gamma", "nonce_gamma_b",
                             "nonce_gamma_beta", "beta_tag_w", "r_w", "nonce_w_b", "nonce_w_beta", "pdl")}
    for u_ in r:
        for f, v in (("gamma", u_.gamma_i), ("k", u_.k_i), ("blind", u_.blind), ("r_k", u_.r_k), ("l", u_.l), ("ped_s1", u_.ped_s1), ("ped_s2", u_.ped_s2),
                     ("heg_s1", u_.heg_s1), ("heg_s2", u_.heg_s2), ("alice", list(u_.alice)), ("beta_tag_gamma", u_.beta_tag_gamma), ("r_gamma", u_.r_gamma),
                     ("nonce_gamma_b", u_.nonce_gamma_b), ("nonce_gamma_beta", u_.nonce_gamma_beta), ("beta_tag_w", u_.beta_tag_w), ("r_w", u_.r_w),
                     ("nonce_w_b", u_.nonce_w_b), ("nonce_w_beta", u_.nonce_w_beta), ("pdl", u_.pdl)):
            g_rnd[f].append(v)
    lam = [lagrange_at_zero(i_ - 1, [0, 2]) for i_ in s_l]
    gout = gg20_general.offline_batch(eng, ks, 2, [0, 2], [[0, 1, 2], [0, 1, 2]], [lam[p_] * keys2[p_].x_i % o_Q for p_ in range(2)],
                                      [pt_mul(lk_a.pk_vec[s_l[p_] - 1], lam[p_]) for p_ in range(2)], [lk_a.y_sum_s] * 2, g_rnd, messages=True)
    assert not gout["status"].any() and gout["R"][0] == unpack_res_point(res.R[0])
    docs["offline_messages"] = {"messages": gout["messages"][0], "R": E.point(gout["R"][0]), "T_i": E.point(gout["T"][0]),
                                "pdl_statement": wire.pdl_statement(c[0], n_a, eng.secp_mul([gout["R"][0]], [r[0].k_i])[0], gout["R"][0], stb.g, stb.ni, stb.N),
                                "expect": "the array deserialises as Vec<Msg<OfflineProtocolMessage>> and re-serialises to the same JSON; M3's PedersenProof, M5's "
                                          "PDLwSlackProof (against pdl_statement) and M6's HomoELGamalProof (statement G = R, H = base_point2, Y = generator, D = T_i, "
                                          "E = M6[0]) verify; the same session through the fused driver gave the same R"}
    ks.free()
    return docs


def main(out_dir: str):
    import __graft_entry__ as entry
    pkg = entry.load_package()
    from mpecdsa_b200 import wire
    from tests.golden import fixtures
    eng = pkg.Engine(0)
    docs = build_documents(eng, fixtures.load_keyset(0))
    os.makedirs(out_dir, exist_ok=True)
    for name, doc in docs.items():
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            f.write(wire.dumps(doc) + "\n")
    print("wrote", len(docs), "documents to", out_dir)
    eng.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "wire"))
