"""CPU-only tests that pin the oracle (oracle/gg20_oracle.py) to independent implementations and
to the invariants the reference's own tests assert.  The reference holds no golden vectors for
this path, so the anchors are: GMP (the reference's BigInt backend), hashlib,
OpenSSL via `cryptography`, published secp256k1 constants, and protocol self-consistency."""
import ctypes
import hashlib
import random

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle.sampling import Drbg, sample_unit


def _limbs(vals, k):
    return np.frombuffer(b"".join(int(v).to_bytes(4 * k, "little") for v in vals), dtype=np.uint32).reshape(len(vals), k).copy()


def _ints(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


def test_python_pow_equals_gmp_powm(oracle_lib):
    """BigInt::mod_pow == mpz_powm: the Python restatement's pow() against GMP itself."""
    rng = random.Random(1)
    for bits in (1024, 2048, 4096):
        k = bits // 32
        mods = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(16)]
        bases = [rng.getrandbits(bits) % m for m in mods]
        exps = [rng.getrandbits(bits) for _ in mods]
        B, E, M = _limbs(bases, k), _limbs(exps, k), _limbs(mods, k)
        out = np.zeros_like(B)
        oracle_lib.oracle_modexp_batch(B.ctypes.data_as(ctypes.c_void_p), E.ctypes.data_as(ctypes.c_void_p),
                                       M.ctypes.data_as(ctypes.c_void_p), None, out.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_size_t(len(mods)), k, k, 2)
        assert _ints(out) == [pow(b, e, m) for b, e, m in zip(bases, exps, mods)]


def test_config0_paillier_roundtrip_python_vs_gmp(oracle_lib, keyset):
    """BASELINE.json configs[0]: one 2048-bit Paillier encrypt + CRT decrypt on CPU; the Python
    restatement and the GMP twin must agree on the ciphertext bytes and recover m."""
    rng = random.Random(2)
    lk = keyset[0]
    ek = lk.paillier_key_vec[0]
    for _ in range(3):
        m = rng.randrange(o.Q)
        r = rng.randrange(1, ek.n)
        c = o.paillier_encrypt(ek, m, r)
        assert o.paillier_decrypt(lk.dk, c) == m
        P, Qq = _limbs([lk.dk.p], 32), _limbs([lk.dk.q], 32)
        Ml, Rl = _limbs([m], 64), _limbs([r], 64)
        c_out, m_out = np.zeros((1, 128), np.uint32), np.zeros((1, 64), np.uint32)
        oracle_lib.oracle_paillier_roundtrip(*(x.ctypes.data_as(ctypes.c_void_p) for x in (P, Qq, Ml, Rl, c_out, m_out)), 64)
        assert _ints(c_out) == [c] and _ints(m_out) == [m]
    # homomorphic identities used by MtA (mta/mod.rs:140-145)
    a, b, bt = rng.randrange(o.Q), rng.randrange(o.Q), rng.randrange(ek.n)
    ca = o.paillier_encrypt(ek, a, rng.randrange(1, ek.n))
    cb = o.paillier_add(ek, o.paillier_mul(ek, ca, b), o.paillier_encrypt(ek, bt, rng.randrange(1, ek.n)))
    assert o.paillier_decrypt(lk.dk, cb) == (a * b + bt) % ek.n


def test_secp256k1_against_openssl():
    from cryptography.hazmat.primitives.asymmetric import ec
    rng = random.Random(3)
    for k in [1, 2, 3, o.Q - 1] + [rng.randrange(1, o.Q) for _ in range(8)]:
        pub = ec.derive_private_key(k, ec.SECP256K1()).public_key().public_numbers()
        assert o.pt_mul(o.G, k) == (pub.x, pub.y)
    # group law sanity on variable base points
    a, b = rng.randrange(1, o.Q), rng.randrange(1, o.Q)
    A = o.pt_mul(o.G, a)
    assert o.pt_mul(A, b) == o.pt_mul(o.G, a * b % o.Q)
    assert o.pt_add(A, o.pt_mul(o.G, b)) == o.pt_mul(o.G, (a + b) % o.Q)
    assert o.pt_sub(A, A) is None


def test_base_point2_known_answer():
    """curv base_point2: x = SHA256^3(compressed G), on the curve (SURVEY.md §8c known answer)."""
    x = o.pt_compress(o.G)
    for _ in range(3):
        x = hashlib.sha256(x).digest()
    assert int.from_bytes(x, "big") == o.H2X
    assert (o.H2Y * o.H2Y - o.H2X ** 3 - 7) % o.P == 0


def test_sha256_chain_and_encodings():
    assert o.bn_bytes(0) == b"\x00" and o.bn_bytes(255) == b"\xff" and o.bn_bytes(256) == b"\x01\x00"
    assert o.sha256_bigints([1, 2]) == int.from_bytes(hashlib.sha256(b"\x01\x02").digest(), "big")
    # FIPS 180-4 "abc"
    assert o.sha256_bigints([0x616263]) == 0xBA7816BF8F01CFEA414140DE5DAE2223B00361A396177A9CB410FF61F20015AD
    assert len(o.pt_compress(o.G)) == 33 and len(o.pt_uncompressed(o.G)) == 65


def test_alice_proof_generate_verify(keyset):
    """mirrors range_proofs.rs `alice_zkp` (:615-634) + a tampered negative."""
    rng = Drbg(11, "alice")
    lk = keyset[0]
    ek, st = lk.paillier_key_vec[0], lk.h1_h2_n_tilde_vec[1]
    a = rng.scalar()
    r = rng.unit_mod(ek.n)
    c = o.paillier_encrypt(ek, a, r)
    q3 = o.Q ** 3
    pf = o.alice_proof_generate(a, c, ek, st, r, rng.below(q3), rng.unit_mod(ek.n), rng.below(q3 * st.N), rng.below(o.Q * st.N))
    assert o.alice_proof_verify(pf, c, ek, st)
    assert not o.alice_proof_verify(pf, c + 1, ek, st)
    bad = o.AliceProof(pf.z, pf.e, pf.s, q3 + 1, pf.s2)
    assert not o.alice_proof_verify(bad, c, ek, st)                 # :118 range check


def test_bob_proofs(keyset):
    """mirrors range_proofs.rs `bob_zkp` (:636-709) for BobProof and BobProofExt."""
    rng = Drbg(12, "bob")
    lk = keyset[0]
    ek, st = lk.paillier_key_vec[0], lk.h1_h2_n_tilde_vec[2]
    q3 = o.Q ** 3
    for check in (False, True):
        a, b = rng.scalar(), rng.scalar()
        enc_a = o.paillier_encrypt(ek, a, rng.unit_mod(ek.n))
        beta_prim, r = rng.below(ek.n), rng.unit_mod(ek.n)
        mta_out = o.paillier_add(ek, o.paillier_mul(ek, enc_a, b), o.paillier_encrypt(ek, beta_prim, r))
        pf, u = o.bob_proof_generate(enc_a, mta_out, b, beta_prim, ek, st, r, check, rng.below(q3), rng.unit_mod(ek.n),
                                     rng.below(o.Q ** 2 * ek.n), rng.below(o.Q * st.N), rng.below(q3 * st.N),
                                     rng.below(o.Q * st.N), rng.below(q3 * st.N))
        if check:
            assert o.bob_proof_ext_verify(pf, u, enc_a, mta_out, ek, st, o.pt_mul(o.G, b))
            assert not o.bob_proof_ext_verify(pf, u, enc_a, mta_out, ek, st, o.pt_mul(o.G, b + 1))
        else:
            assert o.bob_proof_verify(pf, enc_a, mta_out, ek, st)
            assert not o.bob_proof_verify(pf, enc_a, mta_out + 1, ek, st)


def test_mta_end_to_end(keyset):
    """mirrors mta/test.rs:5-18: alpha + beta == a*b (mod q)."""
    rng = Drbg(13, "mta")
    lk = keyset[0]
    ek = lk.paillier_key_vec[0]
    stmts = lk.h1_h2_n_tilde_vec
    q3 = o.Q ** 3
    a, b = rng.scalar(), rng.scalar()
    r = rng.below(ek.n)
    pr = [(rng.below(q3), rng.unit_mod(ek.n), rng.below(q3 * st.N), rng.below(o.Q * st.N)) for st in stmts]
    m_a = o.message_a(a, ek, r, stmts, pr)
    res = o.message_b(b, ek, m_a, rng.below(ek.n), rng.below(ek.n), stmts, rng.scalar(), rng.scalar())
    assert res is not None
    m_b, beta = res
    alpha, _ = o.verify_proofs_get_alpha(m_b, lk.dk, a)
    assert (alpha + beta) % o.Q == a * b % o.Q
    # wrong statement count -> InvalidKey (mta/mod.rs:119)
    assert o.message_b(b, ek, m_a, 1, 1, stmts[:2], 1, 1) is None


def test_pdl_with_slack(keyset):
    """mirrors zk_pdl_with_slack/test.rs: accept (:11-68) and the x+1 soundness negative (:70-129)."""
    rng = Drbg(14, "pdl")
    lk = keyset[1]
    ek, st = lk.paillier_key_vec[1], lk.h1_h2_n_tilde_vec[0]
    q3 = o.Q ** 3
    x, r = rng.scalar(), rng.unit_mod(ek.n)
    c = o.paillier_encrypt(ek, x, r)
    Gp = o.pt_mul(o.G, rng.scalar())
    Qp = o.pt_mul(Gp, x)
    rnd = (rng.below(q3), 1 + rng.below(ek.n - 2), rng.below(o.Q * st.N), rng.below(q3 * st.N))
    pf = o.pdl_prove(x, r, c, ek, Qp, Gp, st.g, st.ni, st.N, *rnd)
    assert o.pdl_verify(pf, c, ek, Qp, Gp, st.g, st.ni, st.N)
    c_bad = o.paillier_encrypt(ek, x + 1, r)
    pf_bad = o.pdl_prove(x, r, c_bad, ek, Qp, Gp, st.g, st.ni, st.N, *rnd)
    assert not o.pdl_verify(pf_bad, c_bad, ek, Qp, Gp, st.g, st.ni, st.N)


def test_sigma_proofs():
    rng = Drbg(15, "sigma")
    x = rng.scalar()
    pf = o.dlog_prove(x, rng.scalar())
    assert o.dlog_verify(pf)
    assert not o.dlog_verify(o.DLogProof(pf.pk, pf.pk_t_rand_commitment, (pf.challenge_response + 1) % o.Q))
    m, r = rng.scalar(), rng.scalar()
    pp = o.pedersen_prove(m, r, rng.scalar(), rng.scalar())
    assert o.pedersen_verify(pp) and pp.com == o.pt_add(o.pt_mul(o.G, m), o.pt_mul(o.H2, r))
    R = o.pt_mul(o.G, rng.scalar())
    l, sigma = rng.scalar(), rng.scalar()
    T = o.pt_add(o.pt_mul(o.G, sigma), o.pt_mul(o.H2, l))
    S = o.pt_mul(R, sigma)
    hp = o.heg_prove(l, sigma, R, o.H2, o.G, T, S, rng.scalar(), rng.scalar())
    assert o.heg_verify(hp, R, o.H2, o.G, T, S)
    assert not o.heg_verify(hp, R, o.H2, o.G, T, o.pt_add(S, o.G))


@pytest.mark.parametrize("s_l", [[1, 2], [1, 3], [2, 3]])
def test_offline_stage_and_signature(keyset, s_l):
    """mirrors state_machine/sign.rs tests (:728-764): the offline stage completes for every
    signer subset of (t=1,n=3) and the online signature verifies — here additionally under an
    independent ECDSA implementation (`cryptography`/OpenSSL), as gg_2020/test.rs:711-748 does
    with libsecp256k1."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec, utils
    keys = [keyset[i - 1] for i in s_l]
    rng = Drbg(0xB2000005, f"session{s_l}")
    rnd = [sample_unit(rng, keys, s_l, p) for p in range(2)]
    res = o.offline_session(keys, s_l, rnd)
    assert [r.status for r in res] == [0, 0]
    assert res[0].R == res[1].R and res[0].t_vec == res[1].t_vec
    msg = o.sha256_bigints([o.bn_from_bytes(b"ZenGo")])                     # sign.rs:693-696
    parts = [o.local_sig(r.k_i, msg, r.R, r.sigma_i) for r in res]
    r_, s_, recid = o.output_signature(res[0].R, parts)
    y = keys[0].y_sum_s
    assert o.ecdsa_verify(r_, s_, y, msg)
    pub = ec.EllipticCurvePublicNumbers(y[0], y[1], ec.SECP256K1()).public_key()
    pub.verify(utils.encode_dss_signature(r_, s_), msg.to_bytes(32, "big"), ec.ECDSA(utils.Prehashed(hashes.SHA256())))
    assert recid in (0, 1)


def test_offline_stage_detects_corruption(keyset):
    """fault injection in the spirit of gg_2020/test.rs:69-148: a corrupted Paillier randomness
    makes the PDL proof of that party fail for both verifiers."""
    s_l = [1, 2]
    keys = [keyset[0], keyset[1]]
    rng = Drbg(77, "corrupt")
    rnd = [sample_unit(rng, keys, s_l, p) for p in range(2)]
    a, b, rho, g = rnd[0].pdl
    rnd[0].pdl = (a, b, rho, g)
    rnd[0].alice[1] = (rnd[0].alice[1][0] + o.Q ** 3, *rnd[0].alice[1][1:])     # alpha out of range -> s1 > q^3
    res = o.offline_session(keys, s_l, rnd)
    assert res[1].status == o.ST_INVALID_KEY


def test_c_twin_equals_python_restatement(keyset):
    """oracle/gg20_twin.c (GMP + OpenSSL, the reference's scalar call sequence) and oracle/gg20_oracle.py agree on every output
    of an offline session: two independent implementations of the restatement, one in C over the reference's own bignum
    backend, pin each other."""
    import numpy as np
    from oracle import twin
    from oracle.sampling import Drbg, sample_unit
    from tests.golden import fixtures
    keysets = [keyset, fixtures.load_keyset(5)]
    drbg = Drbg(0xB2C7, "twin")
    sess, rnds, want = [], [], []
    for ks_i, a, b in ((0, 0, 2), (1, 1, 0)):
        keys, s_l = [keysets[ks_i][a], keysets[ks_i][b]], [a + 1, b + 1]
        r = [sample_unit(drbg, keys, s_l, p) for p in range(2)]
        sess.append((ks_i, a, b)); rnds += r; want += o.offline_session(keys, s_l, r)
    # the packing of the randomness record (include/tecdsa_b200.h TECDSA_RND_*), restated here so that the CPU suite does not
    # depend on the product package
    rnd = np.zeros((len(rnds), 1408), dtype=np.uint32)

    def put(row, off, limbs, val):
        rnd[row, off:off + limbs] = np.frombuffer(int(val).to_bytes(4 * limbs, "little"), dtype="<u4")

    for u, r in enumerate(rnds):
        for name, (off, limbs) in {"gamma_i": (0, 8), "k_i": (8, 8), "blind": (16, 8), "r_k": (24, 64), "beta_tag_gamma": (832, 64), "r_gamma": (896, 64),
                                   "nonce_gamma_b": (960, 8), "nonce_gamma_beta": (968, 8), "beta_tag_w": (976, 64), "r_w": (1040, 64), "nonce_w_b": (1104, 8),
                                   "nonce_w_beta": (1112, 8), "l": (1120, 8), "ped_s1": (1128, 8), "ped_s2": (1136, 8), "heg_s1": (1392, 8), "heg_s2": (1400, 8)}.items():
            put(u, off, limbs, getattr(r, name))
        for x in range(3):
            for (o_, l_), v in zip(((0, 24), (24, 64), (88, 88), (176, 72)), r.alice[x]):
                put(u, 88 + x * 248 + o_, l_, v)
        for (o_, l_), v in zip(((1144, 24), (1168, 64), (1232, 72), (1304, 88)), r.pdl):
            put(u, o_, l_, v)
    res = twin.offline_batch(twin.KeyTables(keysets), np.array(sess, dtype=np.uint32), rnd, 2)
    I = lambda row: int.from_bytes(row.tobytes(), "little")
    for u, w in enumerate(want):
        assert res.status[u] == w.status == 0
        assert I(res.digest[u]).to_bytes(32, "big") == w.transcript
        assert (I(res.R[u, :8]), I(res.R[u, 8:])) == w.R and I(res.sigma[u]) == w.sigma_i and I(res.k[u]) == w.k_i
        assert [(I(res.t_vec[u, :8]), I(res.t_vec[u, 8:16])), (I(res.t_vec[u, 16:24]), I(res.t_vec[u, 24:]))] == w.t_vec
    # a corrupted range-proof nonce: both implementations reject at the same place
    bad = rnd.copy(); bad[0, 88 + 248:88 + 248 + 24] = 0xFFFFFFFF
    r0 = list(rnds[:2]); import copy; r0[0] = copy.deepcopy(r0[0]); a = r0[0].alice[1]; r0[0].alice[1] = ((1 << 768) - 1, a[1], a[2], a[3])
    w = o.offline_session([keysets[0][0], keysets[0][2]], [1, 3], r0)
    res = twin.offline_batch(twin.KeyTables(keysets), np.array(sess[:1], dtype=np.uint32), bad[:2], 1)
    assert [int(x) for x in res.status] == [x.status for x in w] == [0, 2]
