"""GPU parity tests of the batched GG20 offline stage (tecdsa_gg20_offline_batch) against the
oracle's restatement of sign/rounds.rs Round0..Round6: per-unit status, R, sigma_i, t_vec and the
SHA-256 digest over every emitted message (ciphertexts, range proofs, PDL proofs, sigma proofs)."""
import hashlib

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle.sampling import Drbg, sample_unit

pytestmark = pytest.mark.gpu


def _session_inputs(keyset, pairs, seed):
    """pairs: list of (party_a, party_b) 0-based -> oracle inputs + packed randomness"""
    rng = Drbg(seed, "gg20-gpu")
    sess, rnds, oracle_in = [], [], []
    for a, b in pairs:
        s_l = [a + 1, b + 1]
        keys = [keyset[a], keyset[b]]
        r = [sample_unit(rng, keys, s_l, p) for p in range(2)]
        sess.append((0, a, b)); rnds += r; oracle_in.append((keys, s_l, r))
    return sess, rnds, oracle_in


def _digest_int(row):
    return int.from_bytes(row.tobytes(), "little")


def test_key_tables_derived_on_device(engine, pkg, keyset):
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    assert ks.table(0, 64) == [k.dk.p * k.dk.q for k in keyset]
    assert ks.table(1, 128) == [(k.dk.p * k.dk.q) ** 2 for k in keyset]
    assert ks.table(5, 64) == [k.dk.p ** 2 for k in keyset]
    R = 1 << 1024
    assert ks.table(11, 32) == [pow(k.dk.p, -1, R) for k in keyset]                       # p^-1 mod 2^1024
    assert ks.table(13, 32) == [(-pow(k.dk.q, -1, k.dk.p)) % k.dk.p * R % k.dk.p for k in keyset]   # hp * R mod p
    assert ks.table(15, 32) == [pow(k.dk.p, -1, k.dk.q) * R % k.dk.q for k in keyset]     # (p^-1 mod q) * R mod q
    assert ks.table(16, 64) == [pow(k.dk.p ** 2, -1, k.dk.q ** 2) * (1 << 2048) % k.dk.q ** 2 for k in keyset]
    ks.free()


def test_offline_stage_matches_oracle(engine, pkg, keyset):
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    pairs = [(0, 1), (0, 2), (1, 2), (1, 0), (2, 0)]
    sess, rnds, oracle_in = _session_inputs(keyset, pairs, 0xB2000005)
    res = gg20.offline_batch(engine, ks, sess, gg20.pack_randomness(rnds))
    assert list(res.status) == [0] * (2 * len(pairs))
    for s, (keys, s_l, r) in enumerate(oracle_in):
        want = o.offline_session(keys, s_l, r)
        for p in range(2):
            u = 2 * s + p
            assert want[p].status == 0
            assert gg20.unpack_point(pkg.limbs_to_ints(res.R[u:u + 1])[0]) == want[p].R
            assert pkg.limbs_to_ints(res.sigma[u:u + 1])[0] == want[p].sigma_i
            tv = [gg20.unpack_point(v) for v in pkg.limbs_to_ints(res.t_vec[u].reshape(2, 16))]
            assert tv == want[p].t_vec
            assert _digest_int(res.digest[u]).to_bytes(32, "big") == want[p].transcript, (s, p)
    # intermediate messages, byte for byte: MessageA.c and the first AliceProof of unit 0
    keys, s_l, r = oracle_in[0]
    m_a = o.message_a(r[0].k_i, keys[0].paillier_key_vec[keys[0].i - 1], r[0].r_k, keys[0].h1_h2_n_tilde_vec, r[0].alice)
    U = 2 * len(pairs)
    assert pkg.limbs_to_ints(gg20.debug_field(engine, "CK", U)[:1]) == [m_a.c]
    pf = m_a.range_proofs[0]
    for name, val in (("Z0", pf.z), ("E0", pf.e), ("S0", pf.s), ("S10", pf.s1), ("S20", pf.s2)):
        assert pkg.limbs_to_ints(gg20.debug_field(engine, name, U)[:1]) == [val], name
    ks.free()


def test_offline_signature_verifies_independently(engine, pkg, keyset):
    """end-to-end validity: the online step on the GPU outputs gives an ECDSA signature that an
    independent implementation (OpenSSL via `cryptography`) accepts (cf. gg_2020/test.rs:711-748)."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec, utils
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    sess, rnds, _ = _session_inputs(keyset, [(0, 2), (2, 1)], 4242)
    res = gg20.offline_batch(engine, ks, sess, gg20.pack_randomness(rnds))
    assert not res.status.any()
    msg = o.sha256_bigints([o.bn_from_bytes(b"ZenGo")])
    y = keyset[0].y_sum_s
    pub = ec.EllipticCurvePublicNumbers(y[0], y[1], ec.SECP256K1()).public_key()
    for s in range(2):
        R = gg20.unpack_point(pkg.limbs_to_ints(res.R[2 * s:2 * s + 1])[0])
        parts = [o.local_sig(rnds[2 * s + p].k_i, msg, R, pkg.limbs_to_ints(res.sigma[2 * s + p:2 * s + p + 1])[0]) for p in range(2)]
        r_, s_, _ = o.output_signature(R, parts)
        pub.verify(utils.encode_dss_signature(r_, s_), msg.to_bytes(32, "big"), ec.ECDSA(utils.Prehashed(hashes.SHA256())))
    ks.free()


def test_offline_detects_tampering(engine, pkg, keyset):
    """fault injection (gg_2020/test.rs:69-148 style): out-of-range alpha in one range proof of
    unit 0 -> its peer rejects MessageA with InvalidKey; the untouched session stays OK."""
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    sess, rnds, oracle_in = _session_inputs(keyset, [(0, 1), (1, 2)], 77)
    a = rnds[0].alice[1]
    rnds[0].alice[1] = ((1 << 768) - 1, a[1], a[2], a[3])          # alpha > q^3 -> s1 > q^3
    res = gg20.offline_batch(engine, ks, sess, gg20.pack_randomness(rnds))
    want = o.offline_session(*oracle_in[0])
    assert res.status[1] == pkg.ST_INVALID_KEY == want[1].status
    assert res.status[0] == 0 and list(res.status[2:]) == [0, 0]
    ks.free()


def test_offline_two_keysets_mixed_batch(engine, pkg):
    """several key sets in one batch (BASELINE.json configs[4] reuses 8): every unit OK, and the digests of a sample of
    sessions equal the oracle's."""
    from mpecdsa_b200 import gg20
    from tests.golden import fixtures
    keysets = [fixtures.load_keyset(0), fixtures.load_keyset(1)]
    ks = gg20.KeySets(engine, keysets)
    rng = Drbg(99, "two-keysets")
    sess, rnds, oracle_in = [], [], []
    for s_i, (kidx, a, b) in enumerate([(1, 0, 1), (0, 2, 1), (1, 2, 0), (1, 1, 2)]):
        keys = [keysets[kidx][a], keysets[kidx][b]]
        s_l = [a + 1, b + 1]
        r = [sample_unit(rng, keys, s_l, p) for p in range(2)]
        sess.append((kidx, a, b)); rnds += r; oracle_in.append((keys, s_l, r))
    res = gg20.offline_batch(engine, ks, sess, gg20.pack_randomness(rnds))
    assert not res.status.any()
    for s_i in (0, 3):
        want = o.offline_session(*oracle_in[s_i])
        for p in range(2):
            assert _digest_int(res.digest[2 * s_i + p]).to_bytes(32, "big") == want[p].transcript
    # a larger synthetic batch over both key sets: all units must complete (sum R_dash = G, sum S = y checks inside)
    sessions, rnd = gg20.synthetic_batch(keysets, 96, 5)
    res = gg20.offline_batch(engine, ks, sessions, rnd)
    assert not res.status.any()
    ks.free()


def test_offline_split_batch_equals_plain_batches(engine, pkg):
    """batches of >= 2048 sessions run as two half-batches on two streams / host threads (csrc/gg20.cu): every output must
    equal what the same sessions give in plain single-stream calls (units are independent)."""
    from mpecdsa_b200 import gg20
    from tests.golden import fixtures
    keysets = [fixtures.load_keyset(0), fixtures.load_keyset(1)]
    ks = gg20.KeySets(engine, keysets)
    n = 2049                                        # odd: unequal halves
    sessions, rnd = gg20.synthetic_batch(keysets, n, 11)
    rnd[7, :] = 0                                   # one unit with degenerate randomness: a failing session inside a half
    big = gg20.offline_batch(engine, ks, sessions, rnd)
    for lo, hi in ((0, 700), (700, 1500), (1500, n)):
        part = gg20.offline_batch(engine, ks, sessions[lo:hi], rnd[2 * lo:2 * hi])
        assert np.array_equal(part.status, big.status[2 * lo:2 * hi])
        assert np.array_equal(part.R, big.R[2 * lo:2 * hi])
        assert np.array_equal(part.sigma, big.sigma[2 * lo:2 * hi])
        assert np.array_equal(part.digest, big.digest[2 * lo:2 * hi])
        assert np.array_equal(part.t_vec, big.t_vec[2 * lo:2 * hi])
    assert big.status[6] != 0 or big.status[7] != 0
    assert not big.status[8:].any() and not big.status[:6].any()
    ks.free()


def _be(limbs):
    """[n][8] little-endian limbs -> [n][32] big-endian bytes"""
    return np.ascontiguousarray(limbs[:, ::-1]).astype(">u4").view(np.uint8).reshape(limbs.shape[0], 32)


def test_records_e2e_call_and_online_step(engine, pkg):
    """tecdsa_gg20_offline_records (host buffers in, 256-byte records out) == the C twin of the oracle on every field of the
    record, for 8 key sets; then the device online step (phase7_local_sig / output_signature / verify, party_i.rs:850-936)
    against the oracle and OpenSSL; the kernels' own work counter and the per-launch profile are live."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec, utils
    from mpecdsa_b200 import gg20
    from oracle import twin
    from tests.golden import fixtures
    keysets = fixtures.load_all_keysets()
    assert len(keysets) == 8
    ks = gg20.KeySets(engine, keysets)
    n = 48
    sessions, rnd = gg20.synthetic_batch(keysets, n, 0xB2000005)
    assert len(set(sessions[:, 0].tolist())) == 8
    engine.work(reset=True)
    rec = np.zeros((1, 2 * n, pkg.REC_BYTES), dtype=np.uint8)
    engine.offline_records(ks, None, sessions, n, rnd, rec, pkg.HOST)
    macs = engine.work()
    assert 0.3e9 < macs / (2 * n) < 1.0e9                    # executed multiply-accumulates per unit, counted by the kernels
    rec = rec[0]
    want = twin.offline_batch(twin.KeyTables(keysets), sessions, rnd, 8)
    assert not want.status.any() and not rec[:, 0].any()
    assert np.array_equal(rec[:, 164:196], _be(want.digest))
    assert np.array_equal(rec[:, 34:66], _be(want.sigma)) and np.array_equal(rec[:, 66:98], _be(want.k))
    assert np.array_equal(rec[:, 2:34], _be(want.R[:, :8])) and np.array_equal(rec[:, 1], 2 + (want.R[:, 8] & 1).astype(np.uint8))
    assert np.array_equal(rec[:, 99:131], _be(want.t_vec[:, :8])) and np.array_equal(rec[:, 132:164], _be(want.t_vec[:, 16:24]))
    assert np.array_equal(rec[:, 98], 2 + (want.t_vec[:, 8] & 1).astype(np.uint8)) and not rec[:, 196:].any()
    # the plain batch call gives the same per-unit outputs
    res = gg20.offline_batch(engine, ks, sessions, rnd)
    assert np.array_equal(res.R, want.R) and np.array_equal(res.sigma, want.sigma) and np.array_equal(res.digest, want.digest)
    # online step
    m = o.sha256_bigints([o.bn_from_bytes(b"ZenGo")])
    msg = np.tile(np.frombuffer(m.to_bytes(32, "little"), dtype="<u4"), (n, 1))
    sig = gg20.sign_batch(engine, ks, sessions, msg, res.R, res.sigma, want.k)
    assert not sig["status"].any()
    for s in range(n):
        R = gg20.unpack_point(pkg.limbs_to_ints(res.R[2 * s:2 * s + 1])[0])
        parts = [o.local_sig(pkg.limbs_to_ints(want.k[2 * s + p:2 * s + p + 1])[0], m, R, pkg.limbs_to_ints(res.sigma[2 * s + p:2 * s + p + 1])[0]) for p in range(2)]
        assert pkg.limbs_to_ints(sig["s_i"][2 * s:2 * s + 2]) == parts
        r_, s_, recid = o.output_signature(R, parts)
        assert (pkg.limbs_to_ints(sig["r"][s:s + 1])[0], pkg.limbs_to_ints(sig["s"][s:s + 1])[0], int(sig["recid"][s])) == (r_, s_, recid)
        y = keysets[int(sessions[s, 0])][0].y_sum_s
        pub = ec.EllipticCurvePublicNumbers(y[0], y[1], ec.SECP256K1()).public_key()
        pub.verify(utils.encode_dss_signature(r_, s_), m.to_bytes(32, "big"), ec.ECDSA(utils.Prehashed(hashes.SHA256())))
    # a wrong sigma_i makes the in-tree verify fail for that session only (Error::InvalidSig)
    bad = res.sigma.copy(); bad[5, 0] ^= 1
    sig2 = gg20.sign_batch(engine, ks, sessions, msg, res.R, bad, want.k)
    assert sig2["status"][2] == pkg.ST_INVALID_SIG and not np.delete(sig2["status"], 2).any()
    # per-launch profile: every job-list kernel shows up with its counted work
    prof = engine.profile_step(lambda: engine.offline_records(ks, None, sessions, n, rnd, np.zeros((1, 2 * n, 256), np.uint8), pkg.HOST))
    assert any(k.startswith("nadic_jobs_kernel<64") for k in prof) and any(k.startswith("exp_jobs_kernel<64") for k in prof)
    assert abs(sum(v["mac32"] for v in prof.values()) - macs) <= 1e-6 * macs
    assert all(v["ms"] > 0 for v in prof.values())
    ks.free()


def test_config4_share_digest_sample_vs_cpu_twin(engine, pkg):
    """BASELINE.json configs[4], the per-GPU share at full size: 8 192 sessions = 16 384 units over 8 key sets; every unit
    completes, 2^8 sampled units are bit-compared with the oracle's C twin."""
    from mpecdsa_b200 import gg20
    from oracle import twin
    from tests.golden import fixtures
    keysets = fixtures.load_all_keysets()
    ks = gg20.KeySets(engine, keysets)
    n = 8192
    sessions, rnd = gg20.synthetic_batch(keysets, n, 0xB2000005)
    rec = np.zeros((1, 2 * n, pkg.REC_BYTES), dtype=np.uint8)
    engine.offline_records(ks, None, sessions, n, rnd, rec, pkg.HOST)
    rec = rec[0]
    assert not rec[:, 0].any()
    pick = np.unique(np.linspace(0, n - 1, 128).astype(np.int64))
    units = np.stack([2 * pick, 2 * pick + 1], axis=1).reshape(-1)
    want = twin.offline_batch(twin.KeyTables(keysets), sessions[pick], rnd[units], 32)
    got = rec[units]
    assert not want.status.any()
    assert np.array_equal(got[:, 164:196], _be(want.digest)) and np.array_equal(got[:, 34:66], _be(want.sigma))
    assert np.array_equal(got[:, 2:34], _be(want.R[:, :8]))
    ks.free()
