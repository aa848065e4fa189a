"""CPU tests of oracle/keygen_oracle.py (the key-generation path).  They pin the restatement
the CUDA entry points (tests/test_keygen_gpu.py) are checked against: every proof verifies, every tampered field rejects, the Feldman arithmetic agrees with the committed key fixtures."""
import dataclasses
import random

import pytest

from oracle import gg20_oracle as o
from oracle import keygen_oracle as kg
from tests.golden import fixtures


@pytest.fixture(scope="module")
def keyset():
    return fixtures.load_keyset(0)


def test_correct_key_proof_roundtrip_and_tamper(keyset):
    dk = keyset[0].dk
    n = dk.p * dk.q
    ek = o.EncryptionKey(n, n * n)
    sigma = kg.correct_key_proof(dk)
    assert len(sigma) == 11 and all(0 < s < n for s in sigma)
    assert kg.correct_key_verify(sigma, ek)
    bad = list(sigma); bad[5] = (bad[5] + 1) % n
    assert not kg.correct_key_verify(bad, ek)
    assert not kg.correct_key_verify(sigma[:-1], ek)
    assert not kg.correct_key_verify(sigma, ek, salt=b"other")
    other = keyset[1].dk.p * keyset[1].dk.q
    assert not kg.correct_key_verify(sigma, o.EncryptionKey(other, other * other))
    # a modulus sharing a factor with the primorial P (all primes <= 6379) fails the gcd test before any exponentiation
    for f in (2, 3, 11, 13, 6361, 6379):
        assert not kg.correct_key_verify(sigma, o.EncryptionKey(f * n, (f * n) ** 2))
    assert kg.PRIMORIAL.bit_length() == 9120 and kg.PRIMORIAL % 6379 == 0 and kg.PRIMORIAL % 6389 != 0
    # mask_generation covers the key length: 2048-bit N -> 9 digests -> a 2304-bit mask, reduced mod N
    assert kg.mask_generation(2048, 1).bit_length() > 2048


def _setup(rng, bits=256):
    """a small (p~, q~, h1, xhi) setup; the algebra does not depend on the size"""
    def prime(b):
        while True:
            c = rng.getrandbits(b) | 1 | (1 << (b - 1))
            if all(c % s for s in (3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)) and pow(2, c - 1, c) == 1:
                return c
    p_t, q_t = prime(bits), prime(bits)
    phi = (p_t - 1) * (q_t - 1)
    h1 = rng.randrange(2, p_t * q_t)
    while True:
        xhi = rng.randrange(2, phi)
        try:
            pow(xhi, -1, phi)
            break
        except ValueError:
            continue
    return p_t, q_t, h1, xhi


def test_composite_dlog_proofs_of_h1_h2():
    rng = random.Random(0xB2F1)
    p_t, q_t, h1, xhi = _setup(rng)
    nt, h1, h2, xhi_neg, xhi_inv_neg = kg.h1_h2_n_tilde(p_t, q_t, h1, xhi)
    assert h2 == pow(h1, xhi, nt) and pow(h2, pow(xhi, -1, (p_t - 1) * (q_t - 1)), nt) == h1
    st1, st2 = o.DLogStatement(nt, h1, h2), o.DLogStatement(nt, h2, h1)
    r1, r2 = rng.getrandbits(512), rng.getrandbits(512)
    pf1 = kg.composite_dlog_prove(st1, xhi_neg, r1)
    pf2 = kg.composite_dlog_prove(st2, xhi_inv_neg, r2)
    assert kg.composite_dlog_verify(pf1, st1) and kg.composite_dlog_verify(pf2, st2)
    assert pf1.y == r1 + kg.compute_digest([pf1.x, h1, nt, h2]) * xhi_neg              # over the integers, no reduction
    assert not kg.composite_dlog_verify(pf1, st2)                                        # wrong statement
    assert not kg.composite_dlog_verify(dataclasses.replace(pf1, y=pf1.y + 1), st1)
    assert not kg.composite_dlog_verify(dataclasses.replace(pf1, x=(pf1.x + 1) % nt), st1)
    assert not kg.composite_dlog_verify(pf1, o.DLogStatement(nt, h1 * p_t % nt, h2))     # gcd(g, N) != 1
    assert not kg.composite_dlog_verify(pf1, o.DLogStatement(1 << 100, h1, h2))          # N too small


def test_feldman_vss_matches_fixture_shares(keyset):
    # the fixtures' x_i are f(i) of a degree-1 polynomial with f(0) = secret: rebuild it from two shares
    x1, x2, x3 = (k.x_i for k in keyset)
    a1 = (x2 - x1) % o.Q
    secret = (x1 - a1) % o.Q
    assert o.pt_mul(o.G, secret) == keyset[0].y_sum_s
    vss, shares = kg.vss_share(1, 3, secret, [a1])
    assert shares == [x1, x2, x3]
    for i, s in enumerate(shares, start=1):
        assert kg.vss_validate_share(vss, s, i)
        assert kg.vss_point_commitment(vss, i) == keyset[0].pk_vec[i - 1]
    assert not kg.vss_validate_share(vss, shares[0], 2)
    assert not kg.vss_validate_share(vss, (shares[0] + 1) % o.Q, 1)
    # Lagrange at zero over any two parties recovers the secret (what map_share_to_new_params feeds the signers)
    for a, b in ((0, 1), (0, 2), (1, 2)):
        s_l = [a, b]
        rec = sum(o.lagrange_at_zero(i, s_l) * shares[i] for i in s_l) % o.Q
        assert rec == secret


def test_keygen_three_parties_end_to_end(keyset):
    """KeyGen rounds 1-4 of three parties through the oracle (keygen/rounds.rs:26-300): every check passes, all parties
    derive the same y, the x_i commitments equal the DLogProof keys; one corrupted share / proof is caught."""
    rng = random.Random(0xB2F2)
    n = 3
    u = [rng.randrange(1, o.Q) for _ in range(n)]
    y_i = [o.pt_mul(o.G, x) for x in u]
    bcs, decs = [], []
    for i in range(n):
        p_t, q_t, h1, xhi = _setup(rng, bits=1024 if i == 0 else 256)        # one full-size setup, two small ones for speed
        nt, h1, h2, xhi_neg, xhi_inv_neg = kg.h1_h2_n_tilde(p_t, q_t, h1, xhi)
        bc, dec = kg.phase1_broadcast(keyset[i].dk, nt, h1, h2, xhi_neg, xhi_inv_neg, y_i[i], rng.getrandbits(256),
                                      rng.getrandbits(512), rng.getrandbits(512))
        bcs.append(bc); decs.append(dec)
    assert kg.phase1_verify(bcs[0], decs[0])
    # the small setups fail ONLY the N_tilde bit-length window (party_i.rs:294-295): everything else holds
    for i in (1, 2):
        assert not kg.phase1_verify(bcs[i], decs[i])
        assert kg.composite_dlog_verify(bcs[i].composite_dlog_proof_base_h1, bcs[i].dlog_statement)
        assert kg.correct_key_verify(bcs[i].correct_key_proof, bcs[i].e)
    assert not kg.phase1_verify(bcs[0], dataclasses.replace(decs[0], blind_factor=decs[0].blind_factor + 1))
    vss, shares = zip(*(kg.vss_share(1, n, u[i], [rng.randrange(1, o.Q)]) for i in range(n)))
    out = []
    for me in range(n):
        mine = [shares[j][me] for j in range(n)]
        res = kg.phase2_verify_vss(y_i, mine, vss, me + 1, rng.randrange(1, o.Q))
        assert res is not None
        out.append(res)
    assert len({r[0] for r in out}) == 1 and out[0][0] == o.pt_mul(o.G, sum(u) % o.Q)
    proofs = [r[2] for r in out]
    assert kg.verify_dlog_proofs_check_against_vss(proofs, vss)
    assert kg.commitments_to_xi(vss) == [o.pt_mul(o.G, r[1]) for r in out]
    bad_mine = [shares[j][0] for j in range(n)]; bad_mine[1] = (bad_mine[1] + 1) % o.Q
    assert kg.phase2_verify_vss(y_i, bad_mine, vss, 1, 5) is None
    assert not kg.verify_dlog_proofs_check_against_vss([proofs[1], proofs[0], proofs[2]], vss)


def test_keygen_golden_vectors(keyset):
    """tests/golden/vectors_keygen.json (make_keygen_vectors.py): the restatement reproduces its frozen outputs"""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "vectors_keygen.json")) as f:
        v = json.load(f)
    I = lambda s: int(s, 16)
    for e in v["correct_key"]:
        dk = keyset[e["row"]].dk
        n = dk.p * dk.q
        salt = bytes.fromhex(e["salt"])
        assert [I(x) for x in e["rho"]] == kg._rho_vec(n, salt)
        sigma = [I(x) for x in e["sigma"]]
        assert sigma == kg.correct_key_proof(dk, salt)
        assert kg.correct_key_verify(sigma, o.EncryptionKey(n, n * n), salt)
    for e in v["composite_dlog"]:
        st1 = o.DLogStatement(I(e["n_tilde"]), I(e["h1"]), I(e["h2"]))
        st2 = o.DLogStatement(st1.N, st1.ni, st1.g)
        pf1 = kg.composite_dlog_prove(st1, I(e["xhi_neg"]), I(e["r1"]))
        pf2 = kg.composite_dlog_prove(st2, I(e["xhi_inv_neg"]), I(e["r2"]))
        assert [pf1.x, pf1.y] == [I(x) for x in e["proof_h1"]] and [pf2.x, pf2.y] == [I(x) for x in e["proof_h2"]]
        assert kg.composite_dlog_verify(pf1, st1) and kg.composite_dlog_verify(pf2, st2)
    # N = f * q with gcd(N, phi(N)) = 1: all eleven sigma^N == rho checks pass, only gcd(P, N) decides
    q_big = I(v["small_factor"]["q"])
    for e in v["small_factor"]["cases"]:
        n = e["p"] * q_big
        sigma = [I(x) for x in e["sigma"]]
        assert all(pow(s_, n, n) == r for s_, r in zip(sigma, kg._rho_vec(n, kg.SALT_STRING)))
        assert kg.correct_key_verify(sigma, o.EncryptionKey(n, n * n)) == e["accept"] == (e["p"] > 6379)
    for e in v["vss"]:
        vss, shares = kg.vss_share(e["t"], e["n"], I(e["secret"]), [I(c) for c in e["coefficients"]])
        assert shares == [I(s) for s in e["shares"]]
        assert vss.commitments == [(I(p[0]), I(p[1])) for p in e["commitments"]]
