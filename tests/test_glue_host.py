"""CPU-only: the product's single-thread glue device code (csrc/secp256k1.cuh, sha256.cuh,
st_bigint.cuh, gg20_glue.cuh) compiled for the host through tests/host_harness (CUDA qualifier
shims) and checked against the oracle.  This covers the EC / hashing / CRT / sigma-proof logic of
the GPU path without a GPU; the lane-group big-integer kernels are covered by the `-m gpu` tests."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import gg20_oracle as o
from oracle.sampling import Drbg

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_harness")
KEY_SIZE = [64, 128, 64, 64, 64, 64, 64, 32, 32, 32, 32, 32, 32, 32, 32, 32, 64, 32, 32, 8, 16]


@pytest.fixture(scope="module")
def h():
    so = os.path.join(HERE, "libglue_host.so")
    srcs = [os.path.join(HERE, "harness.cpp")] + [os.path.join(entry.CSRC, f) for f in os.listdir(entry.CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-w", "-I", HERE, "-I", entry.CSRC, "-shared", "-fPIC", "-o", so, os.path.join(HERE, "harness.cpp")])
    lib = ctypes.CDLL(so)
    lib.h_init()
    return lib


def L(v, k):
    return np.frombuffer(int(v).to_bytes(4 * k, "little"), dtype=np.uint32).copy()


def I(a):
    return int.from_bytes(np.ascontiguousarray(a).tobytes(), "little")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def PT(p):
    return L(0 if p is None else p[0] | (p[1] << 256), 16)


def UNPT(a):
    v = I(a)
    return None if v == 0 else (v & ((1 << 256) - 1), v >> 256)


def test_field_and_scalar_arithmetic(h):
    rng = random.Random(1)
    o8 = np.zeros(8, np.uint32)
    for _ in range(200):
        a, b = rng.randrange(o.P), rng.randrange(o.P)
        h.h_fe_mul(P(o8), P(L(a, 8)), P(L(b, 8))); assert I(o8) == a * b % o.P
        a, b = rng.randrange(o.Q), rng.randrange(o.Q)
        h.h_sc_mul(P(o8), P(L(a, 8)), P(L(b, 8))); assert I(o8) == a * b % o.Q
    for a in (1, 2, o.Q - 1, rng.randrange(1, o.Q)):
        h.h_sc_inv(P(o8), P(L(a, 8))); assert I(o8) == pow(a, -1, o.Q)
    for bits in (1, 255, 256, 257, 768, 2048, 2817):
        x = rng.getrandbits(bits) | (1 << (bits - 1))
        n = (bits + 31) // 32
        h.h_sc_from_limbs(P(o8), P(L(x, n)), n); assert I(o8) == x % o.Q
    edge = [(o.P - 1, o.P - 1), (0, 5), (o.P - 1, 1), (2**256 - 1 - o.P, 3)]
    for a, b in edge:
        h.h_fe_mul(P(o8), P(L(a, 8)), P(L(b, 8))); assert I(o8) == a * b % o.P


def test_point_arithmetic(h):
    rng = random.Random(2)
    o16 = np.zeros(16, np.uint32)
    for k in [1, 2, 3, o.Q - 1, 0] + [rng.randrange(o.Q) for _ in range(10)]:
        h.h_pt_mul(P(o16), P(PT(o.G)), P(L(k, 8))); assert UNPT(o16) == o.pt_mul(o.G, k)
    A, B = o.pt_mul(o.G, 1234567), o.pt_mul(o.H2, 987654321)
    for x, y in ((A, B), (A, A), (A, o.pt_neg(A)), (None, B), (A, None)):
        h.h_pt_add(P(o16), P(PT(x)), P(PT(y))); assert UNPT(o16) == o.pt_add(x, y)
    for k in [1, 15, 16, o.Q - 1, 0] + [rng.randrange(o.Q) for _ in range(20)]:      # fixed-base tables (G and base_point2)
        h.h_mul_fixed(P(o16), 0, P(L(k, 8))); assert UNPT(o16) == o.pt_mul(o.G, k)
        h.h_mul_fixed(P(o16), 1, P(L(k, 8))); assert UNPT(o16) == o.pt_mul(o.H2, k)
    o8 = np.zeros(8, np.uint32)
    for own, peer in ((0, 1), (1, 0), (0, 2), (2, 1)):
        h.h_lagrange2(P(o8), own, peer); assert I(o8) == o.lagrange_at_zero(own, [own, peer])


def test_transcript_hashes(h, keyset):
    rng = random.Random(3)
    ek = keyset[0].paillier_key_vec[0]
    e8 = np.zeros(8, np.uint32)
    for _ in range(3):
        c, z, u, w = rng.randrange(ek.nn), rng.randrange(ek.n), rng.randrange(ek.nn), rng.randrange(ek.n)
        if _ == 1:
            z >>= 40; u >>= 100                     # leading zero bytes: minimal-length encoding matters
        h.h_alice_hash(P(e8), P(L(ek.n, 64)), P(L(c, 128)), P(L(z, 64)), P(L(u, 128)), P(L(w, 64)))
        assert I(e8) == o.sha256_bigints([ek.n, ek.n + 1, c, z, u, w])
    Gp, Qp, u1 = o.pt_mul(o.G, 5), o.pt_mul(o.G, 77), o.pt_mul(o.G, 991)
    c, z, u2, u3 = rng.randrange(ek.nn), rng.randrange(ek.n), rng.randrange(ek.nn), rng.randrange(ek.n)
    h.h_pdl_hash(P(e8), P(PT(Gp)), P(PT(Qp)), P(L(c, 128)), P(L(z, 64)), P(PT(u1)), P(L(u2, 128)), P(L(u3, 64)))
    assert I(e8) == o.sha256_bigints([o.bn_from_bytes(o.pt_compress(Gp)), o.bn_from_bytes(o.pt_compress(Qp)), c, z,
                                      o.bn_from_bytes(o.pt_compress(u1)), u2, u3])
    for blind in (rng.getrandbits(256), rng.getrandbits(200), 0):
        h.h_hash_commit(P(e8), P(PT(Gp)), P(L(blind, 8)))
        assert I(e8) == o.hash_commitment(o.bn_from_bytes(o.pt_compress(Gp)), blind)


def test_sigma_proofs(h):
    rng = Drbg(4, "host-sigma")
    out = np.zeros(40, np.uint32)
    sk, nonce = rng.scalar(), rng.scalar()
    h.h_dlog_prove(P(out), P(L(sk, 8)), P(L(nonce, 8)))
    want = o.dlog_prove(sk, nonce)
    assert (UNPT(out[:16]), UNPT(out[16:32]), I(out[32:])) == (want.pk, want.pk_t_rand_commitment, want.challenge_response)
    assert h.h_dlog_verify(P(out)) == 1
    out[32] ^= 1
    assert h.h_dlog_verify(P(out)) == 0
    m, r = rng.scalar(), rng.scalar()
    pp = o.pedersen_prove(m, r, rng.scalar(), rng.scalar())
    ped = np.concatenate([L(pp.e, 8), PT(pp.a1), PT(pp.a2), L(pp.z1, 8), L(pp.z2, 8), np.zeros(8, np.uint32)])
    assert h.h_pedersen_verify(P(ped), P(PT(pp.com))) == 1
    assert h.h_pedersen_verify(P(ped), P(PT(o.pt_add(pp.com, o.G)))) == 0
    R = o.pt_mul(o.G, rng.scalar()); l, sigma = rng.scalar(), rng.scalar()
    T = o.pt_add(o.pt_mul(o.G, sigma), o.pt_mul(o.H2, l)); S = o.pt_mul(R, sigma)
    hp = o.heg_prove(l, sigma, R, o.H2, o.G, T, S, rng.scalar(), rng.scalar())
    heg = np.concatenate([PT(hp.T), PT(hp.A3), L(hp.z1, 8), L(hp.z2, 8)])
    assert h.h_heg_verify(P(heg), P(PT(R)), P(PT(T)), P(PT(S))) == 1
    assert h.h_heg_verify(P(heg), P(PT(R)), P(PT(T)), P(PT(o.pt_add(S, o.G)))) == 0


def test_key_setup_and_decrypt_tail(h, keyset):
    rows = 3
    tabs = [np.zeros((rows, s), np.uint32) for s in KEY_SIZE]
    for r, lk in enumerate(keyset):
        tabs[9][r] = L(lk.dk.p, 32); tabs[10][r] = L(lk.dk.q, 32)
    ptrs = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    h.h_key_setup(ptrs, rows)
    R = 1 << 1024
    rng = random.Random(5)
    for r, lk in enumerate(keyset):
        p, q = lk.dk.p, lk.dk.q
        assert I(tabs[0][r]) == p * q and I(tabs[1][r]) == (p * q) ** 2 and I(tabs[5][r]) == p * p and I(tabs[6][r]) == q * q
        assert I(tabs[7][r]) == p - 1 and I(tabs[8][r]) == q - 1
        assert I(tabs[11][r]) == pow(p, -1, R) and I(tabs[12][r]) == pow(q, -1, R)
        assert I(tabs[13][r]) == (-pow(q, -1, p)) % p * R % p and I(tabs[14][r]) == (-pow(p, -1, q)) % q * R % q
        assert I(tabs[15][r]) == pow(p, -1, q) * R % q
        R64 = 1 << 2048
        assert I(tabs[16][r]) == pow(p * p, -1, q * q) * R64 % (q * q)
        assert I(tabs[17][r]) == q % (p - 1) and I(tabs[18][r]) == p % (q - 1)
        b = rng.randrange(1, p * q)          # the identity behind the 1024-bit stage
        assert pow(pow(b % p, q % (p - 1), p), p, p * p) == pow(b, p * q, p * p)
        for _ in range(2):          # CRT recombination of an own-key power b^N mod N^2
            b = rng.randrange(1, p * q)
            yp, yq = pow(b, p * q, p * p), pow(b, p * q, q * q)
            out = np.zeros(128, np.uint32)
            h.h_crt_combine(P(out), ptrs, r, P(L(yp, 64)), P(L(yq, 64)))
            assert I(out) == pow(b, p * q, (p * q) ** 2)
        ek = lk.paillier_key_vec[r]
        for _ in range(2):
            m = rng.randrange(ek.n)
            c = o.paillier_encrypt(ek, m, rng.randrange(1, ek.n))
            dp, dq = pow(c % (p * p), p - 1, p * p), pow(c % (q * q), q - 1, q * q)
            out = np.zeros(64, np.uint32)
            h.h_decrypt_finish(P(out), ptrs, r, P(L(dp, 64)), P(L(dq, 64)))
            assert I(out) == m


def test_plain_integer_responses(h):
    """s1 = e*a + alpha, s2 = e*ro + gamma (range_proofs.rs:87-88) at the ABI's field widths"""
    rng = random.Random(6)
    for _ in range(5):
        e, a, alpha = rng.getrandbits(256), rng.randrange(o.Q), rng.getrandbits(767)
        d = np.zeros(28, np.uint32)
        h.h_mul_add(P(d), 28, P(L(e, 8)), 8, P(L(a, 8)), 8, P(L(alpha, 24)), 24)
        assert I(d) == e * a + alpha
        ro, gamma = rng.getrandbits(2303), rng.getrandbits(2815)
        d = np.zeros(92, np.uint32)
        h.h_mul_add(P(d), 92, P(L(e, 8)), 8, P(L(ro, 72)), 72, P(L(gamma, 88)), 88)
        assert I(d) == e * ro + gamma


LAMBDA = 0x5363ad4cc05c30e0a5261c028812645a122e22ea20816678df02967c1b23bd72


def test_glv_split_wnaf_and_scalar_mul(h):
    """round-2 EC layer: GLV split k = k1 + k2*lambda with |k_i| < 2^128 for edge and random scalars, signed fixed 5-bit windows, and
    the resulting `Point * Scalar` against the oracle's double-and-add (itself pinned to OpenSSL)"""
    rng = random.Random(0xEC)
    m1, m2, neg = np.zeros(8, np.uint32), np.zeros(8, np.uint32), (ctypes.c_int * 2)()
    edge = [1, 2, 3, o.Q - 1, o.Q - 2, LAMBDA, o.Q - LAMBDA, (o.Q + 1) // 2, (o.Q - 1) // 2, 1 << 255, 1 << 128, (1 << 128) - 1, (1 << 127) + 1, 0xFFFFFFFF]
    for k in edge + [rng.randrange(1, o.Q) for _ in range(3000)]:
        h.h_glv_split(P(m1), P(m2), neg, P(L(k, 8)))
        a, b = I(m1), I(m2)
        assert a < 1 << 128 and b < 1 << 128
        assert ((-a if neg[0] else a) + (-b if neg[1] else b) * LAMBDA) % o.Q == k
    digits = (ctypes.c_byte * 27)()
    for m in [0, 1, 15, 16, 17, 31, 32, (1 << 128) - 1, 1 << 127, 1 << 128, (1 << 129) - 1] + [rng.getrandbits(128) for _ in range(300)]:
        h.h_signed_windows5(digits, P(L(m, 8)))
        d = [digits[i] for i in range(27)]
        assert sum(v << (5 * i) for i, v in enumerate(d)) == m and all(-15 <= v <= 16 for v in d)
    out = np.zeros(16, np.uint32)
    pts = [o.G, o.H2, o.pt_mul(o.G, 0xDEADBEEF)]
    for k in edge + [rng.randrange(1, o.Q) for _ in range(60)]:
        pt = pts[k % 3]
        h.h_pt_mul(P(out), P(PT(pt)), P(L(k, 8)))
        assert UNPT(out) == o.pt_mul(pt, k), hex(k)
    h.h_pt_mul(P(out), P(PT(o.G)), P(L(0, 8))); assert UNPT(out) is None
    h.h_pt_mul(P(out), P(PT(o.G)), P(L(o.Q, 8))); assert UNPT(out) is None           # un-reduced scalar = 0
    h.h_pt_mul(P(out), P(PT(None)), P(L(5, 8))); assert UNPT(out) is None


def test_fixed_base_tables_inversion_chain_and_projective_compare(h):
    rng = random.Random(0xEC2)
    out = np.zeros(16, np.uint32)
    for which, base in ((0, o.G), (1, o.H2)):
        for k in [1, 255, 256, 257, o.Q - 1, 1 << 248, (1 << 256) - 1 - (1 << 255)] + [rng.randrange(1, o.Q) for _ in range(40)]:
            h.h_mul_fixed(P(out), which, P(L(k % o.Q, 8)))
            assert UNPT(out) == o.pt_mul(base, k % o.Q)
    o8 = np.zeros(8, np.uint32)
    for a in [1, 2, o.P - 1, 977] + [rng.randrange(1, o.P) for _ in range(100)]:
        h.h_fe_inv(P(o8), P(L(a, 8))); assert I(o8) == pow(a, -1, o.P)
    A, B = o.pt_mul(o.G, 77), o.pt_mul(o.G, 78)
    for _ in range(20):
        z = rng.randrange(2, o.P)
        assert h.h_jac_eq_affine(P(PT(A)), P(L(z, 8)), P(PT(A))) == 3
        assert h.h_jac_eq_affine(P(PT(A)), P(L(z, 8)), P(PT(B))) == 0
        assert h.h_jac_eq_affine(P(PT(A)), P(L(z, 8)), P(PT(o.pt_neg(A)))) == 0
    assert h.h_jac_eq_affine(P(PT(None)), P(L(1, 8)), P(PT(None))) == 3 and h.h_jac_eq_affine(P(PT(A)), P(L(5, 8)), P(PT(None))) == 0
    out48 = np.zeros(48, np.uint32)
    k = rng.randrange(1, o.Q)
    h.h_to_affine3(P(out48), P(PT(A)), P(PT(None)), P(PT(B)), P(L(k, 8)))
    assert [UNPT(out48[0:16]), UNPT(out48[16:32]), UNPT(out48[32:48])] == [o.pt_mul(A, k), None, o.pt_mul(B, k)]
    # lagrange constants for every signer pair of n = 3
    for own in range(3):
        for peer in range(3):
            if own != peer:
                h.h_lagrange2(P(o8), own, peer); assert I(o8) == o.lagrange_at_zero(own, [own, peer])
