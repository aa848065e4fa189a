"""Wire-format module: structure of the serde-JSON documents, LocalKey round trip, and (GPU)
documents built from engine outputs equal those built from the oracle's values."""
import json

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle import keygen_oracle as kg
from oracle.sampling import Drbg, sample_unit


def _decompress(b: bytes):
    x = int.from_bytes(b[1:], "big")
    y = pow((x ** 3 + 7) % o.P, (o.P + 1) // 4, o.P)
    if y & 1 != b[0] & 1:
        y = o.P - y
    return (x, y)


@pytest.mark.parametrize("bytes_as", ["hex", "array"])
def test_local_key_roundtrip_and_shapes(pkg, keyset, bytes_as):
    from mpecdsa_b200 import wire
    enc = wire.Encoding(bytes_as=bytes_as)
    lk = keyset[1]
    comm = [o.pt_mul(o.G, 5), o.pt_mul(o.G, 7)]
    doc = json.loads(wire.dumps(wire.local_key(lk, comm, enc)))
    assert list(doc) == ["paillier_dk", "pk_vec", "keys_linear", "paillier_key_vec", "y_sum_s", "h1_h2_n_tilde_vec", "vss_scheme", "i", "t", "n"]   # keygen/rounds.rs:310-322
    assert doc["i"] == 2 and doc["t"] == 1 and doc["n"] == 3 and doc["vss_scheme"]["parameters"] == {"threshold": 1, "share_count": 3}
    assert doc["paillier_dk"]["p"] == "%x" % lk.dk.p and len(doc["pk_vec"]) == 3
    assert doc["pk_vec"][0]["curve"] == "secp256k1"
    raw = doc["pk_vec"][0]["point"]
    assert (bytes.fromhex(raw) if bytes_as == "hex" else bytes(raw)) == o.pt_compress(lk.pk_vec[0])
    back = wire.load_local_key(doc, _decompress, enc)
    assert (back.i, back.x_i, back.dk.p, back.dk.q, back.pk_vec, back.y_sum_s) == (lk.i, lk.x_i, lk.dk.p, lk.dk.q, lk.pk_vec, lk.y_sum_s)
    assert [s.N for s in back.h1_h2_n_tilde_vec] == [s.N for s in lk.h1_h2_n_tilde_vec] and back.vss_commitments == comm
    assert [e.n for e in back.paillier_key_vec] == [e.n for e in lk.paillier_key_vec]
    # BigInt leaf: lower-case hex of to_bytes(); zero is one 0x00 byte
    assert enc.bigint(0) == "00" and enc.bigint(255) == "ff" and enc.bigint(256) == "0100" and enc.bigint_from("0100") == 256
    assert enc.scalar_from(enc.scalar(12345)) == 12345


def test_message_documents_follow_the_struct_definitions(pkg, keyset):
    from mpecdsa_b200 import wire
    pf = {"z": 1, "e": 2, "s": 3, "s1": 4, "s2": 5}
    assert list(wire.alice_proof(pf)) == ["z", "e", "s", "s1", "s2"]                                   # range_proofs.rs:95-101
    d = o.dlog_prove(7, 9)
    row = np.frombuffer(b"".join(int(v).to_bytes(n, "little") for v, n in ((d.pk[0] | d.pk[1] << 256, 64), (d.pk_t_rand_commitment[0] | d.pk_t_rand_commitment[1] << 256, 64),
                                                                          (d.challenge_response, 32))), dtype="<u4")
    m = wire.message_b(10, row, row)
    assert list(m) == ["c", "b_proof", "beta_tag_proof"] and list(m["b_proof"]) == ["pk", "pk_t_rand_commitment", "challenge_response"]   # mta/mod.rs:41-46
    assert bytes.fromhex(m["b_proof"]["pk"]["point"]) == o.pt_compress(d.pk)
    assert int(m["b_proof"]["challenge_response"]["scalar"], 16) == d.challenge_response
    assert list(wire.message_a(10, [pf, pf])) == ["c", "range_proofs"]
    pd = wire.pdl_proof({"z": 1, "u1": o.G, "u2": 2, "u3": 3, "s1": 4, "s2": 5, "s3": 6})
    assert list(pd) == ["z", "u1", "u2", "u3", "s1", "s2", "s3"]                                     # zk_pdl_with_slack/mod.rs:57-66
    st = wire.pdl_statement(1, 15, o.G, o.H2, 2, 3, 35)
    assert list(st) == ["ciphertext", "ek", "Q", "G", "h1", "h2", "N_tilde"] and st["ek"] == {"n": "0f"}
    bc = wire.keygen_broadcast1(15, (35, 2, 3), 99, [1] * 11, (4, 5), (6, 7))
    assert list(bc) == ["e", "dlog_statement", "com", "correct_key_proof", "composite_dlog_proof_base_h1", "composite_dlog_proof_base_h2"]   # party_i.rs:96-104
    assert wire.signature_recid(1, 2, 1) == {"r": {"curve": "secp256k1", "scalar": "%064x" % 1}, "s": {"curve": "secp256k1", "scalar": "%064x" % 2}, "recid": 1}


@pytest.mark.gpu
def test_fixture_documents_from_engine_outputs(engine, pkg, keyset):
    """tools/emit_wire_fixtures.py: the documents handed to the Rust test are built from ENGINE outputs; they must equal the
    documents built from the oracle's values for the same inputs, and reload into a working key set."""
    from mpecdsa_b200 import gg20, wire
    from tools.emit_wire_fixtures import build_documents
    docs = build_documents(engine, keyset, seed=0xB2F2)
    lk_a, lk_b = keyset[0], keyset[2]
    ek_a = lk_a.paillier_key_vec[0]
    rng = Drbg(0xB2F2, "wire")
    r = [sample_unit(rng, [lk_a, lk_b], [1, 3], p) for p in range(2)]
    m_a = o.message_a(r[0].k_i, ek_a, r[0].r_k, lk_a.h1_h2_n_tilde_vec, r[0].alice)
    want = wire.message_a(m_a.c, [{"z": p.z, "e": p.e, "s": p.s, "s1": p.s1, "s2": p.s2} for p in m_a.range_proofs])
    assert docs["message_a"]["message"] == want
    mb, beta = o.message_b(r[1].gamma_i, ek_a, m_a, r[1].r_gamma, r[1].beta_tag_gamma, lk_b.h1_h2_n_tilde_vec, r[1].nonce_gamma_b, r[1].nonce_gamma_beta)
    assert docs["message_b"]["message"]["c"] == wire.DEFAULT.bigint(mb.c)
    assert docs["message_b"]["message"]["b_proof"]["challenge_response"] == wire.DEFAULT.scalar(mb.b_proof.challenge_response)
    assert docs["message_b"]["expected_alpha_plus_beta"] == wire.DEFAULT.scalar(r[0].k_i * r[1].gamma_i % o.Q)
    back = wire.load_local_key(docs["local_key"], lambda b: engine.point_decompress([b])[0])
    assert back.pk_vec == lk_a.pk_vec and back.x_i == lk_a.x_i
    sig, _ = kg.correct_key_proof(lk_a.dk), None
    assert docs["keygen_broadcast1"]["message"]["correct_key_proof"]["sigma_vec"] == [wire.DEFAULT.bigint(x) for x in sig]


def test_round_message_wrappers_follow_the_reference_definitions(pkg):
    """state_machine/sign.rs:478-490 (`OfflineProtocolMessage(OfflineM)`, externally tagged, tuple bodies as arrays), sign/rounds.rs:33-49
    (transparent newtypes), gg_2020/party_i.rs:113-122 (`SignBroadcastPhase1`, `SignDecommitPhase1`), round_based `Msg`"""
    from mpecdsa_b200 import wire
    E = wire.DEFAULT
    ped = np.arange(64, dtype=np.uint32)
    d = wire.pedersen_proof(ped)
    assert list(d) == ["e", "a1", "a2", "com", "z1", "z2"] and d["e"] == E.scalar(int.from_bytes(ped[:8].tobytes(), "little"))
    heg = wire.heg_proof(np.arange(48, dtype=np.uint32))
    assert list(heg) == ["T", "A3", "z1", "z2"]
    assert wire.sign_broadcast_phase1(255) == {"com": "ff"}
    assert list(wire.sign_decommit_phase1(1, (o.G[0], o.G[1]))) == ["blind_factor", "g_gamma_i"]
    m = wire.msg(2, None, wire.offline_message("M4", wire.sign_decommit_phase1(1, (o.G[0], o.G[1]))))
    assert list(m) == ["sender", "receiver", "body"] and m["receiver"] is None and list(m["body"]) == ["M4"]
    assert wire.msg(1, 2, wire.offline_message("M2", [{}, {}]))["receiver"] == 2
    with pytest.raises(AssertionError):
        wire.offline_message("M7", {})


def test_offline_messages_parse_back(pkg):
    """wire.parse_offline_message inverts the emitters on the engine-produced document committed for the Rust test
    (bindings/rust/tests/data/offline_messages.json), with the oracle's decompression standing in for `Point::from_bytes`"""
    import json, os
    from mpecdsa_b200 import wire
    E = wire.DEFAULT

    def decompress(b):
        x = int.from_bytes(b[1:], "big")
        y = pow((x ** 3 + 7) % o.P, (o.P + 1) // 4, o.P)
        return (x, y if (y & 1) == (b[0] & 1) else o.P - y)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings", "rust", "tests", "data", "offline_messages.json")
    doc = json.load(open(path))
    parsed = [wire.parse_offline_message(m, decompress) for m in doc["messages"]]
    assert [p["kind"] for p in parsed] == ["M1", "M2", "M3", "M4", "M5", "M6"] and [p["receiver"] for p in parsed] == [None, 2, None, None, None, None]
    m1, m2, m3, m4, m5, m6 = parsed
    # re-emitting the parsed values gives the document back, byte for byte
    assert wire.message_a(m1["c"], m1["range_proofs"]) == doc["messages"][0]["body"]["M1"][0] and wire.sign_broadcast_phase1(m1["com"]) == doc["messages"][0]["body"]["M1"][1]
    assert wire.sign_decommit_phase1(m4["blind_factor"], m4["g_gamma"]) == doc["messages"][3]["body"]["M4"]
    assert [wire.pdl_proof(p) for p in m5["pdl"]] == doc["messages"][4]["body"]["M5"][1] and E.point(m5["R_dash"]) == doc["messages"][4]["body"]["M5"][0]
    # and the values are the protocol's: the commitment opens, T is the Pedersen commitment, the sigma proofs verify under the oracle
    assert o.hash_commitment(o.bn_from_bytes(o.pt_compress(m4["g_gamma"])), m4["blind_factor"]) == m1["com"]
    ped = m3["pedersen"]
    assert ped["com"] == m3["T"] and o.pedersen_verify(o.PedersenProof(ped["e"], ped["a1"], ped["a2"], ped["com"], ped["z1"], ped["z2"]))
    R = decompress(E.raw_from(doc["R"]["point"]))
    heg = m6["heg"]
    assert o.heg_verify(o.HomoElGamalProof(heg["T"], heg["A3"], heg["z1"], heg["z2"]), R, o.H2, o.G, m3["T"], m6["S"])
    assert all(o.dlog_verify(o.DLogProof(d["pk"], d["pk_t_rand_commitment"], d["challenge_response"])) for mb in (m2["gamma"], m2["w"]) for d in (mb["b_proof"], mb["beta_tag_proof"]))


def test_committed_engine_documents_verify_under_the_oracle(pkg):
    """Every document under bindings/rust/tests/data/ was produced on a B200 (tools/emit_wire_fixtures.py); here, without a GPU, the
    oracle's verifiers judge them — the same judgement the Rust test asks of the reference."""
    import json, os
    from oracle import lindell17_oracle as l17
    from mpecdsa_b200 import wire
    E = wire.DEFAULT
    base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings", "rust", "tests", "data")
    load = lambda name: json.load(open(os.path.join(base, name + ".json")))
    B, S = E.bigint_from, E.scalar_from

    def pt(d):
        b = E.raw_from(d["point"])
        x = int.from_bytes(b[1:], "big")
        y = pow((x ** 3 + 7) % o.P, (o.P + 1) // 4, o.P)
        return (x, y if (y & 1) == (b[0] & 1) else o.P - y)
    ek = lambda d: o.EncryptionKey(B(d["n"]), B(d["n"]) ** 2)
    st = lambda d: o.DLogStatement(B(d["N"]), B(d["g"]), B(d["ni"]))
    ap = lambda d: o.AliceProof(*(B(d[k]) for k in ("z", "e", "s", "s1", "s2")))
    d = load("alice_proof")
    assert o.alice_proof_verify(ap(d["proof"]), B(d["cipher"]), ek(d["ek"]), st(d["dlog_statement"]))
    d = load("message_a")
    assert all(o.alice_proof_verify(ap(pf), B(d["message"]["c"]), ek(d["ek"]), st(s_)) for pf, s_ in zip(d["message"]["range_proofs"], d["dlog_statements"]))
    d = load("dlog_proof")["proof"]
    assert o.dlog_verify(o.DLogProof(pt(d["pk"]), pt(d["pk_t_rand_commitment"]), S(d["challenge_response"])))
    d = load("pdl")
    s_, p_ = d["statement"], d["proof"]
    pf = o.PDLwSlackProof(B(p_["z"]), pt(p_["u1"]), B(p_["u2"]), B(p_["u3"]), B(p_["s1"]), B(p_["s2"]), B(p_["s3"]))
    assert o.pdl_verify(pf, B(s_["ciphertext"]), ek(s_["ek"]), pt(s_["Q"]), pt(s_["G"]), B(s_["h1"]), B(s_["h2"]), B(s_["N_tilde"]))
    d = load("signature")
    assert o.ecdsa_verify(S(d["sig"]["r"]), S(d["sig"]["s"]), pt(d["y"]), B(d["message"]))
    d = load("lindell17_eph_first_message")["message"]
    proof = o.ECDDHProof(pt(d["d_log_proof"]["a1"]), pt(d["d_log_proof"]["a2"]), S(d["d_log_proof"]["z"]))
    assert o.ecddh_verify(proof, o.G, pt(d["public_share"]), o.H2, pt(d["c"]))
    d = load("lindell17_signature")
    dk = o.DecryptionKey(B(d["party_one_paillier"]["dk"]["p"]), B(d["party_one_paillier"]["dk"]["q"]))
    r, s, recid = l17.p1_sign(dk, B(d["c3"]), S(d["party_one_eph"]["secret_share"]), pt(d["party_two_eph_public"]))
    assert (r, s, recid) == (B(d["signature"]["r"]), B(d["signature"]["s"]), d["recid"]) and l17.verify(r, s, pt(d["pubkey"]), B(d["message"]))
    d = load("message_b")
    mb = d["message"]
    dl = lambda x: o.DLogProof(pt(x["pk"]), pt(x["pk_t_rand_commitment"]), S(x["challenge_response"]))
    got = o.verify_proofs_get_alpha(o.MessageB(B(mb["c"]), dl(mb["b_proof"]), dl(mb["beta_tag_proof"])), o.DecryptionKey(B(d["dk"]["p"]), B(d["dk"]["q"])), S(d["a"]))
    assert got is not None and got[0] == S(d["expected_alpha"]) and (got[0] + S(d["beta"])) % o.Q == S(d["expected_alpha_plus_beta"])
