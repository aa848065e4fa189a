"""BASELINE.json configs[2] and configs[3] at FULL size on one GPU, as tests: every rate printed here
is also a parity check — configs[2]: 2^20 generator multiplications with 2^16 of them compared with OpenSSL, plus a
variable-base run checked through (k2 * (k1 G)) == (k1 k2) G; configs[3]: 16 384 AliceProof generate + verify with 2^10 proofs
byte-compared with the oracle and a 1 % tampered subset that must be rejected (and nothing else).  configs[4]'s per-GPU share
is tests/test_gg20_gpu.py::test_config4_share_digest_sample_vs_cpu_twin; configs[1] runs inside bench.py with all 65 536
outputs compared with GMP."""
import json
import multiprocessing as mp
import time

import numpy as np
import pytest

from oracle import gg20_oracle as o

pytestmark = pytest.mark.gpu
I = lambda row: int.from_bytes(row.tobytes(), "little")


def _openssl_pub(k: int):
    from cryptography.hazmat.primitives.asymmetric import ec
    pub = ec.derive_private_key(k, ec.SECP256K1()).public_key().public_numbers()
    return pub.x, pub.y


def test_config2_one_million_scalar_muls(engine, pkg):
    pkg._bind_l01(engine.lib)
    rng = np.random.default_rng(0xB2000003)
    n = 1 << 20
    k = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint32)
    k[:, 7] &= 0x7FFFFFFF                                     # below q
    k[:, 0] |= 1
    out = np.zeros((n, 16), dtype=np.uint32)
    dt = None
    for _ in range(2):
        t0 = time.perf_counter()
        engine._ck(engine.lib.tecdsa_secp_mul_batch(engine._ctx, None, k.ctypes.data, out.ctypes.data, n, 0), "secp_mul")
        dt = time.perf_counter() - t0
    idx = np.linspace(0, n - 1, 1 << 16).astype(np.int64)
    for i in idx:
        v = I(out[i])
        assert (v & ((1 << 256) - 1), v >> 256) == _openssl_pub(I(k[i])), i
    comp = engine.point_compress([(I(out[i, :8]), I(out[i, 8:])) for i in idx[:64]])
    assert comp == [o.pt_compress((I(out[i, :8]), I(out[i, 8:]))) for i in idx[:64]]
    # variable base: k2 * (k1 G) == (k1 k2) G
    nv = 1 << 18
    pts, out2 = out[:nv].copy(), np.zeros((nv, 16), dtype=np.uint32)
    t0 = time.perf_counter()
    engine._ck(engine.lib.tecdsa_secp_mul_batch(engine._ctx, pts.ctypes.data, k[nv:2 * nv].ctypes.data, out2.ctypes.data, nv, 0), "secp_mul")
    dtv = time.perf_counter() - t0
    for i in range(0, nv, nv // 1024):
        v = I(out2[i])
        assert (v & ((1 << 256) - 1), v >> 256) == _openssl_pub(I(k[i]) * I(k[nv + i]) % o.Q), i
    print("\nCONFIG2 " + json.dumps({"generator_base_mul_per_s_host_buffers": n / dt, "variable_base_mul_per_s_host_buffers": nv / dtv, "count": n,
                                     "checked_vs_openssl": int(len(idx)) + 1024}))


def _oracle_proof(args):
    a, c, n, st, r, al, be, ga, ro = args
    w = o.alice_proof_generate(a, c, o.EncryptionKey(n, n * n), o.DLogStatement(*st), r, al, be, ga, ro)
    return (w.z, w.e, w.s, w.s1, w.s2)


def test_config3_16k_range_proofs(engine, pkg):
    from mpecdsa_b200 import gg20
    from tests.golden import fixtures
    pkg._bind_l01(engine.lib); gg20._bind_l2(engine.lib)
    keysets = fixtures.load_all_keysets()
    rows = [lk for ks_ in keysets for lk in ks_]               # 16 Paillier keys / (N~, h1, h2) setups are asked for: the first 16 rows
    ks = gg20.KeySets(engine, keysets)
    n_rows = 16
    rng = np.random.default_rng(0xB2000004)
    m = 16384
    ek_row = rng.integers(0, n_rows, size=m).astype(np.uint32)
    st_row = rng.integers(0, n_rows, size=m).astype(np.uint32)
    ns = [lk.dk.p * lk.dk.q for lk in rows]
    sts = [(lk.h1_h2_n_tilde_vec[lk.i - 1].N, lk.h1_h2_n_tilde_vec[lk.i - 1].g, lk.h1_h2_n_tilde_vec[lk.i - 1].ni) for lk in rows]

    def rnd_bits(bits_per_row, limbs):
        a = rng.integers(0, 2**32, size=(m, limbs), dtype=np.uint32)
        bits = np.asarray(bits_per_row)
        col = np.arange(limbs)[None, :]
        full, rem = bits[:, None] // 32, bits[:, None] % 32
        a[col > full] = 0
        top = (col == full)
        a = np.where(top, a & ((np.uint64(1) << rem.astype(np.uint64)) - np.uint64(1)).astype(np.uint32), a)
        a[:, 0] |= 1
        return np.ascontiguousarray(a.astype(np.uint32))

    q3b = (o.Q ** 3).bit_length() - 1
    a_ = rnd_bits([255] * m, 8)
    r_ = rnd_bits([ns[e].bit_length() - 1 for e in ek_row], 64)
    al, be = rnd_bits([q3b] * m, 24), rnd_bits([ns[e].bit_length() - 1 for e in ek_row], 64)
    ga = rnd_bits([(o.Q ** 3 * sts[s][0]).bit_length() - 1 for s in st_row], 88)
    ro = rnd_bits([(o.Q * sts[s][0]).bit_length() - 1 for s in st_row], 72)
    N = pkg.ints_to_limbs(ns, 64)
    a64 = np.zeros((m, 64), np.uint32); a64[:, :8] = a_
    c = np.zeros((m, 128), np.uint32)
    P = lambda x: x.ctypes.data
    engine._ck(engine.lib.tecdsa_paillier_encrypt_batch(engine._ctx, P(N), P(ek_row), len(ns), P(a64), P(r_), P(c), m, 0), "enc")
    z, e, s = np.zeros((m, 64), np.uint32), np.zeros((m, 8), np.uint32), np.zeros((m, 64), np.uint32)
    s1, s2 = np.zeros((m, 28), np.uint32), np.zeros((m, 92), np.uint32)
    t0 = time.perf_counter()
    engine._ck(engine.lib.tecdsa_alice_proof_generate_batch(engine._ctx, ks.handle, P(ek_row), P(st_row), P(a_), P(c), P(r_), P(al), P(be), P(ga), P(ro),
                                                            P(z), P(e), P(s), P(s1), P(s2), m, 0), "gen")
    dt_g = time.perf_counter() - t0
    st = np.full(m, 255, np.uint8)
    t0 = time.perf_counter()
    engine._ck(engine.lib.tecdsa_alice_proof_verify_batch(engine._ctx, ks.handle, P(ek_row), P(st_row), P(c), P(z), P(e), P(s), P(s1), P(s2), P(st), m, 0), "ver")
    dt_v = time.perf_counter() - t0
    assert not st.any()
    # 1 % tampered: exactly those reject
    bad = np.sort(rng.choice(m, size=m // 100, replace=False))
    c2 = c.copy(); c2[bad, 5] ^= 1
    st2 = np.full(m, 255, np.uint8)
    engine._ck(engine.lib.tecdsa_alice_proof_verify_batch(engine._ctx, ks.handle, P(ek_row), P(st_row), P(c2), P(z), P(e), P(s), P(s1), P(s2), P(st2), m, 0), "ver")
    assert np.array_equal(np.nonzero(st2)[0], bad)
    # 2^10 proofs byte-compared with the oracle (Python integers; a process pool keeps it to seconds)
    pick = np.linspace(0, m - 1, 1 << 10).astype(np.int64)
    jobs = [(I(a_[i]), I(c[i]), ns[ek_row[i]], sts[st_row[i]], I(r_[i]), I(al[i]), I(be[i]), I(ga[i]), I(ro[i])) for i in pick]
    with mp.get_context("fork").Pool(8) as pool:
        want = pool.map(_oracle_proof, jobs, chunksize=16)
    got = [(I(z[i]), I(e[i]), I(s[i]), I(s1[i]), I(s2[i])) for i in pick]
    assert got == want
    ks.free()
    print("\nCONFIG3 " + json.dumps({"proofs": m, "generate_s": dt_g, "verify_s": dt_v, "pairs_per_s_host_buffers": m / (dt_g + dt_v),
                                     "achieved_tmac32_reference_oplist": 2.73e8 * m / (dt_g + dt_v) / 1e12, "byte_compared_with_oracle": len(pick), "tampered": len(bad)}))
