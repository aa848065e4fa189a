"""GPU parity tests of the key-generation entry points against oracle/keygen_oracle.py and the
frozen vectors of tests/golden/vectors_keygen.json, through the C ABI."""
import json
import os
import random

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle import keygen_oracle as kg

pytestmark = pytest.mark.gpu


def _vectors():
    with open(os.path.join(os.path.dirname(__file__), "golden", "vectors_keygen.json")) as f:
        return json.load(f)

OK, PROOF = 0, 10


def test_correct_key_verify_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    ns, sig = [], []
    for k in keyset:
        ns.append(k.dk.p * k.dk.q)
        sig.append(kg.correct_key_proof(k.dk))
    # accepted proofs, then one tampered sigma, a proof against the wrong key, a wrong salt (separate call)
    bad = [list(s) for s in sig]
    bad[0][3] = (bad[0][3] + 1) % ns[0]
    bad[1] = sig[2]
    st = keygen.correct_key_verify(engine, ns + ns, sig + bad)
    want = [kg.correct_key_verify(s, o.EncryptionKey(n, n * n)) for n, s in zip(ns + ns, sig + bad)]
    assert want == [True, True, True, False, False, True]
    assert list(st) == [OK if w else PROOF for w in want]
    st = keygen.correct_key_verify(engine, ns, sig, salt=b"other")
    assert list(st) == [PROOF] * 3
    # frozen proofs
    v = _vectors()
    st = keygen.correct_key_verify(engine, [ns[e["row"]] for e in v["correct_key"]], [[int(x, 16) for x in e["sigma"]] for e in v["correct_key"]])
    assert list(st) == [OK] * len(v["correct_key"])
    # gcd(P, N) == 1 for the primorial of all primes <= 6379: N = f * q with every sigma^N == rho check passing
    q_big = int(v["small_factor"]["q"], 16)
    cases = v["small_factor"]["cases"]
    st = keygen.correct_key_verify(engine, [e["p"] * q_big for e in cases], [[int(x, 16) for x in e["sigma"]] for e in cases])
    assert [e["accept"] for e in cases] == [False, False, False, False, True]
    assert list(st) == [OK if e["accept"] else PROOF for e in cases]
    # an even modulus, and over-wide / negative proof fields, reject that proof only
    st = keygen.correct_key_verify(engine, [ns[0] - 1, ns[0], ns[1]], [sig[0], [sig[0][0] + (1 << 2048)] + sig[0][1:], sig[1]])
    assert list(st) == [PROOF, PROOF, OK]


def test_composite_dlog_verify_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    from tests.test_keygen_oracle import _setup
    rng = random.Random(0xB2F3)
    stmts, proofs, want = [], [], []
    for i in range(6):
        p_t, q_t, h1, xhi = _setup(rng, bits=1024 if i < 2 else 512)
        nt, h1, h2, xn, xin = kg.h1_h2_n_tilde(p_t, q_t, h1, xhi)
        st1, st2 = o.DLogStatement(nt, h1, h2), o.DLogStatement(nt, h2, h1)
        pf1 = kg.composite_dlog_prove(st1, xn, rng.getrandbits(512))
        pf2 = kg.composite_dlog_prove(st2, xin, rng.getrandbits(512))
        stmts += [(nt, h1, h2), (nt, h2, h1), (nt, h1, h2), (nt, h1, h2)]
        proofs += [(pf1.x, pf1.y), (pf2.x, pf2.y), (pf1.x, pf1.y + 1), (pf2.x, pf2.y)]
        want += [True, True, False, kg.composite_dlog_verify(pf2, st1)]
    stmts.append((stmts[0][0], stmts[0][1] * 3, stmts[0][2])); proofs.append(proofs[0]); want.append(False)
    st = keygen.composite_dlog_verify(engine, stmts, proofs)
    assert [kg.composite_dlog_verify(kg.CompositeDLogProof(*p), o.DLogStatement(*s)) for s, p in zip(stmts, proofs)] == want
    assert list(st) == [OK if w else PROOF for w in want]


def test_vss_validate_share_batch(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    rng = random.Random(0xB2F4)
    comms, shares, idx, want = [], [], [], []
    for _ in range(3):                                 # one batch = one polynomial degree (t = 1, as in the fixtures)
        vss, sh = kg.vss_share(1, 5, rng.randrange(1, o.Q), [rng.randrange(1, o.Q)])
        for i, s in enumerate(sh, start=1):
            comms.append(vss.commitments); shares.append(s); idx.append(i); want.append(True)
        comms.append(vss.commitments); shares.append(sh[0]); idx.append(2); want.append(False)
        comms.append(vss.commitments); shares.append((sh[0] + 1) % o.Q); idx.append(1); want.append(False)
    st = keygen.vss_validate_share(engine, comms, shares, idx)
    assert list(st) == [OK if w else PROOF for w in want]
    vss, sh = kg.vss_share(3, 4, 12345, [7, 8, 9])
    st = keygen.vss_validate_share(engine, [vss.commitments] * 4, sh, [1, 2, 3, 4])
    assert list(st) == [OK] * 4


def test_correct_key_prove_matches_oracle_and_verifies(engine, pkg, keyset):
    from mpecdsa_b200 import keygen
    pq = [(k.dk.p, k.dk.q) for k in keyset]
    sig, st = keygen.correct_key_prove(engine, pq)
    assert list(st) == [OK] * 3
    assert sig == [kg.correct_key_proof(k.dk) for k in keyset]
    v = _vectors()
    for e in v["correct_key"]:
        assert sig[e["row"]] == [int(x, 16) for x in e["sigma"]]                      # frozen vectors
    assert list(keygen.correct_key_verify(engine, [p * q for p, q in pq], sig)) == [OK] * 3
    sig2, _ = keygen.correct_key_prove(engine, pq[:1], salt=b"other")
    assert sig2[0] == kg.correct_key_proof(keyset[0].dk, b"other") and sig2[0] != sig[0]


def test_h1_h2_n_tilde_and_composite_dlog_prove(engine, pkg):
    from mpecdsa_b200 import keygen
    from tests.test_keygen_oracle import _setup
    rng = random.Random(0xB2F5)
    setups = [_setup(rng, bits=1024) for _ in range(3)] + [_setup(rng, bits=512)]
    got, st = keygen.h1_h2_n_tilde(engine, setups)
    assert list(st) == [OK] * 4
    assert got == [kg.h1_h2_n_tilde(*s) for s in setups]
    # an even xhi has no inverse modulo phi: the reference's sampling loop would draw again
    p_t, q_t, h1, xhi = setups[0]
    _, st = keygen.h1_h2_n_tilde(engine, [(p_t, q_t, h1, xhi + 1 if xhi % 2 else xhi), (p_t, q_t, h1, (p_t - 1) // 2 * 3)])
    assert list(st) == [4, 4]                                                          # TECDSA_ST_NOT_INVERTIBLE
    stmts, secrets, nonces, want = [], [], [], []
    for nt, h1, h2, xn, xin in got:
        for stmt, sec in (((nt, h1, h2), xn), ((nt, h2, h1), xin)):
            r = rng.getrandbits(512)
            stmts.append(stmt); secrets.append(sec); nonces.append(r)
            pf = kg.composite_dlog_prove(o.DLogStatement(*stmt), sec, r)
            want.append((pf.x, pf.y))
    proofs = keygen.composite_dlog_prove(engine, stmts, secrets, nonces)
    assert proofs == want
    assert list(keygen.composite_dlog_verify(engine, stmts, proofs)) == [OK] * len(stmts)
    v = _vectors()
    for e in v["composite_dlog"]:
        I = lambda s: int(s, 16)
        pf = keygen.composite_dlog_prove(engine, [(I(e["n_tilde"]), I(e["h1"]), I(e["h2"])), (I(e["n_tilde"]), I(e["h2"]), I(e["h1"]))],
                                         [I(e["xhi_neg"]), I(e["xhi_inv_neg"])], [I(e["r1"]), I(e["r2"])])
        assert [list(p) for p in pf] == [[I(x) for x in e["proof_h1"]], [I(x) for x in e["proof_h2"]]]


def test_vss_share_batch(engine, pkg):
    from mpecdsa_b200 import keygen
    rng = random.Random(0xB2F6)
    for t, n in ((1, 3), (2, 5), (3, 4)):
        polys = [[rng.randrange(1, o.Q) for _ in range(t + 1)] for _ in range(4)]
        shares, comms = keygen.vss_share(engine, t, n, polys)
        for p, sh, cm in zip(polys, shares, comms):
            vss, want = kg.vss_share(t, n, p[0], p[1:])
            assert sh == want and cm == vss.commitments
        st = keygen.vss_validate_share(engine, [c for c in comms for _ in range(n)], [s for sh in shares for s in sh], list(range(1, n + 1)) * 4)
        assert list(st) == [OK] * (4 * n)
    v = _vectors()
    for e in v["vss"]:
        I = lambda s: int(s, 16)
        shares, comms = keygen.vss_share(engine, e["t"], e["n"], [[I(e["secret"])] + [I(c) for c in e["coefficients"]]])
        assert shares[0] == [I(s) for s in e["shares"]] and comms[0] == [(I(p[0]), I(p[1])) for p in e["commitments"]]


def test_keygen_phases_three_parties(engine, pkg, keyset):
    """Keys::phase1_broadcast.. / phase1_verify.. / phase2_verify_vss.. / verify_dlog_proofs_check_against_vss for n = 3, t = 1:
    every message produced by the engine equals the oracle's, every check agrees, one corrupted sender is singled out"""
    from mpecdsa_b200 import gg20, keygen
    from tests.test_keygen_oracle import _setup
    rng = random.Random(0xB2F7)
    n, t = 3, 1
    u = [rng.randrange(1, o.Q) for _ in range(n)]
    y_i = [o.pt_mul(o.G, x) for x in u]
    setups = [_setup(rng, bits=1024) for _ in range(n)]
    blinds = [rng.getrandbits(256) for _ in range(n)]
    r1 = [rng.getrandbits(512) for _ in range(n)]
    r2 = [rng.getrandbits(512) for _ in range(n)]
    want_bc, want_dec = [], []
    for i in range(n):
        nt, h1, h2, xn, xin = kg.h1_h2_n_tilde(*setups[i])
        bc, dec = kg.phase1_broadcast(keyset[i].dk, nt, h1, h2, xn, xin, y_i[i], blinds[i], r1[i], r2[i])
        want_bc.append(bc); want_dec.append(dec)
    # engine side of phase 1
    params, st = keygen.h1_h2_n_tilde(engine, setups)
    sig, _ = keygen.correct_key_prove(engine, [(k.dk.p, k.dk.q) for k in keyset])
    pf = keygen.composite_dlog_prove(engine, [(p[0], p[1], p[2]) for p in params] + [(p[0], p[2], p[1]) for p in params],
                                     [p[3] for p in params] + [p[4] for p in params], r1 + r2)
    com = gg20.hash_commitment(engine, y_i, blinds)
    for i in range(n):
        bc = want_bc[i]
        assert (params[i][0], params[i][1], params[i][2]) == (bc.dlog_statement.N, bc.dlog_statement.g, bc.dlog_statement.ni)
        assert sig[i] == bc.correct_key_proof and com[i] == bc.com
        assert pf[i] == (bc.composite_dlog_proof_base_h1.x, bc.composite_dlog_proof_base_h1.y)
        assert pf[n + i] == (bc.composite_dlog_proof_base_h2.x, bc.composite_dlog_proof_base_h2.y)
    ok = keygen.phase1_verify(engine, want_bc, want_dec)
    assert list(ok) == [kg.phase1_verify(b, d) for b, d in zip(want_bc, want_dec)] == [True] * n
    import dataclasses
    bad_bc = list(want_bc)
    bad_bc[1] = dataclasses.replace(bad_bc[1], com=bad_bc[1].com ^ 1)
    bad_bc[2] = dataclasses.replace(bad_bc[2], correct_key_proof=want_bc[0].correct_key_proof)
    assert list(keygen.phase1_verify(engine, bad_bc, want_dec)) == [True, False, False]
    # phase 2: Feldman shares
    polys = [[u[i], rng.randrange(1, o.Q)] for i in range(n)]
    shares, comms = keygen.vss_share(engine, t, n, polys)
    vss = [kg.VerifiableSS(t, n, comms[i]) for i in range(n)]
    for me in range(1, n + 1):
        mine = [shares[j][me - 1] for j in range(n)]
        assert list(keygen.phase2_verify_vss(engine, y_i, mine, vss, me)) == [True] * n
        want = kg.phase2_verify_vss(y_i, mine, vss, me, 5)
        assert want is not None
    mine = [shares[j][0] for j in range(n)]; mine[2] = (mine[2] + 1) % o.Q
    assert list(keygen.phase2_verify_vss(engine, y_i, mine, vss, 1)) == [True, True, False]
    # phase 3: DLogProof of x_i against the summed commitments
    x = [sum(shares[j][i] for j in range(n)) % o.Q for i in range(n)]
    nonces = [rng.randrange(1, o.Q) for _ in range(n)]
    proofs = gg20.dlog_prove(engine, x, nonces)
    assert list(gg20.dlog_verify(engine, proofs)) == [OK] * n
    xi = kg.commitments_to_xi(vss)
    for i in range(n):
        pk = gg20.unpack_point(int.from_bytes(proofs[i, :16].tobytes(), "little"))
        assert pk == xi[i] == o.dlog_prove(x[i], nonces[i]).pk
