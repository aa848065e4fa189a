"""GPU parity tests of tecdsa_modexp_batch (== curv `BigInt::mod_pow`, GMP mpz_powm) through the
C ABI: bit-exact against the oracle on seeded inputs, the edge cases, and size-independent
properties at BASELINE.json's full batch size."""
import random

import numpy as np
import pytest

from oracle import gg20_oracle as o

pytestmark = pytest.mark.gpu


def _rand_case(rng, bits, count, ebits):
    mods = [rng.getrandbits(bits) | (1 << (bits - 1)) | 1 for _ in range(count)]
    bases = [rng.getrandbits(bits) % m for m in mods]
    exps = [rng.getrandbits(ebits) for _ in range(count)]
    return bases, exps, mods


@pytest.mark.parametrize("bits,tpis", [(2048, (4, 8, 16, 32)), (4096, (8, 16, 32)), (1024, (4, 8, 16))])
def test_modexp_matches_oracle_all_shapes(engine, bits, tpis):
    rng = random.Random(bits)
    for tpi in tpis:
        engine.set_tpi(bits, tpi)
        for ebits in (256, 768, bits):
            bases, exps, mods = _rand_case(rng, bits, 67, ebits)
            got, st = engine.mod_pow(bases, exps, mods, mod_bits=bits, exp_bits=ebits)
            assert list(st) == [0] * len(mods)
            assert got == [pow(b, e, m) for b, e, m in zip(bases, exps, mods)], (bits, tpi, ebits)
    engine.set_tpi(bits, 0)


def test_modexp_edge_cases(engine):
    bits = 2048
    R = 1 << bits
    rng = random.Random(7)
    n = rng.getrandbits(bits) | (1 << (bits - 1)) | 1
    small = rng.getrandbits(2047) | (1 << 2046) | 1           # 2047-bit modulus (PAILLIER_MIN_BIT_LENGTH, party_i.rs:49)
    cases = [
        (0, 0, n), (0, 5, n), (1, 0, n), (5, 0, 1), (5, 7, 1), (n - 1, 2, n), (n - 1, 3, n), (2, R - 1, n),
        (R - 1, 3, n),                       # base >= modulus is reduced
        (n, 3, n), (n + 1, 5, n),
        (3, 12345, small), (small - 1, small - 2, small),
        (2, 10, 3), (7, 2**2047, 5), (R - 1, R - 1, R - 1), (3, 65537, (1 << 2047) + 1),
    ]
    b, e, m = zip(*cases)
    got, st = engine.mod_pow(b, e, m, mod_bits=bits, exp_bits=2048)
    assert list(st) == [0] * len(cases)
    assert got == [pow(x, y, z) for x, y, z in cases]
    # even modulus: flagged per element, never a crash, other elements unaffected
    got, st = engine.mod_pow([3, 3, 3], [5, 5, 5], [n, n + 1, n], mod_bits=bits)
    assert list(st) == [0, 1, 0] and got[0] == got[2] == pow(3, 5, n) and got[1] == 0


def test_modexp_empty_and_ragged(engine, pkg):
    got, st = engine.mod_pow([], [], [])
    assert got == [] and len(st) == 0
    rng = random.Random(9)
    for count in (1, 3, 31, 33, 129):           # not multiples of the lanes-per-block packing
        bases, exps, mods = _rand_case(rng, 2048, count, 300)
        got, st = engine.mod_pow(bases, exps, mods)
        assert got == [pow(b, e, m) for b, e, m in zip(bases, exps, mods)]


def test_modexp_shared_moduli_index(engine, pkg, keyset):
    """mod_idx gather: many operands over a few Paillier moduli (how the proofs use it)."""
    rng = random.Random(10)
    mods = [k.paillier_key_vec[i].n for i, k in enumerate(keyset)]
    nn = [m * m for m in mods]
    count = 90
    idx = np.array([rng.randrange(3) for _ in range(count)], dtype=np.uint32)
    bases = [rng.getrandbits(4096) % nn[i] for i in idx]
    exps = [rng.getrandbits(2048) for _ in range(count)]
    B, E, M = pkg.ints_to_limbs(bases, 128), pkg.ints_to_limbs(exps, 64), pkg.ints_to_limbs(nn, 128)
    out = np.zeros_like(B)
    st = np.full(count, 255, np.uint8)
    engine.modexp_raw(4096, 64, B, E, M, out, st, mod_idx=idx, n_mod=3)
    assert list(st) == [0] * count
    assert pkg.limbs_to_ints(out) == [pow(b, e, nn[i]) for b, e, i in zip(bases, exps, idx)]


def test_modexp_device_buffers(engine, pkg):
    import torch
    rng = random.Random(11)
    bases, exps, mods = _rand_case(rng, 2048, 200, 2048)
    t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
    B, E, M = t(pkg.ints_to_limbs(bases, 64)), t(pkg.ints_to_limbs(exps, 64)), t(pkg.ints_to_limbs(mods, 64))
    out = torch.zeros_like(B)
    st = torch.full((200,), 255, dtype=torch.uint8, device="cuda")
    engine.modexp_raw(2048, 64, B, E, M, out, st, mem=pkg.DEVICE)
    engine.sync()
    assert st.sum().item() == 0
    assert pkg.limbs_to_ints(out.cpu().numpy().view(np.uint32)) == [pow(b, e, m) for b, e, m in zip(bases, exps, mods)]


def test_modexp_full_batch_rsa_roundtrip(engine, pkg, keyset):
    """BASELINE.json configs[1] size (65 536 operands): size-independent property instead of an
    element-wise oracle — RSA round trip x^(e*d) == x over Paillier moduli n = p*q, plus a 2^10
    element sample compared with the oracle."""
    import math
    count = 65536
    rng = np.random.default_rng(0xB2000002)
    ns, ds = [], []
    e_pub = 65537
    for k in keyset:
        lam = math.lcm(k.dk.p - 1, k.dk.q - 1)
        ns.append(k.dk.p * k.dk.q)
        ds.append(pow(e_pub, -1, lam))
    idx = rng.integers(0, 3, size=count).astype(np.uint32)
    X = rng.integers(0, 2**32, size=(count, 64), dtype=np.uint32)
    X[:, 63] &= 0x3FFFFFFF                                      # x < n
    M = pkg.ints_to_limbs(ns, 64)
    E1 = np.zeros((count, 64), np.uint32); E1[:, 0] = e_pub
    D = pkg.ints_to_limbs(ds, 64)
    E2 = D[idx]
    C = np.zeros_like(X); Y = np.zeros_like(X)
    st = np.full(count, 255, np.uint8)
    engine.modexp_raw(2048, 64, X, E1, M, C, st, mod_idx=idx, n_mod=3)
    assert not st.any()
    engine.modexp_raw(2048, 64, C, np.ascontiguousarray(E2), M, Y, st, mod_idx=idx, n_mod=3)
    assert not st.any()
    assert np.array_equal(X, Y)
    sample = np.linspace(0, count - 1, 1024).astype(np.int64)
    xs = pkg.limbs_to_ints(X[sample]); cs = pkg.limbs_to_ints(C[sample])
    assert cs == [pow(x, e_pub, ns[i]) for x, i in zip(xs, idx[sample])]


def test_modexp_ten_thousand_random_cases_vs_gmp(engine, pkg, oracle_lib):
    """>= 10^4 random cases per SURVEY.md section 8(c)'s bit-exactness definition: every output of the GPU path compared
    with GMP mpz_powm (the reference's BigInt backend) — 8192 cases at 2048 bits, 2048 at 4096 bits, 4096 at 1024 bits."""
    import ctypes
    import os
    threads = max(1, len(os.sched_getaffinity(0)))
    for bits, count, seed in ((2048, 8192, 1), (4096, 2048, 2), (1024, 4096, 3)):
        k = bits // 32
        rng = np.random.default_rng(seed)
        base = rng.integers(0, 2**32, size=(count, k), dtype=np.uint32)
        exp = rng.integers(0, 2**32, size=(count, k), dtype=np.uint32)
        mod = rng.integers(0, 2**32, size=(count, k), dtype=np.uint32)
        mod[:, 0] |= 1
        mod[: count // 2, k - 1] |= 0x80000000                 # half with the top bit set, half arbitrary odd moduli
        exp[::7, k // 2:] = 0                                  # short exponents
        out = np.zeros_like(base)
        st = np.full(count, 255, np.uint8)
        engine.modexp_raw(bits, k, base, exp, mod, out, st)
        assert not st.any()
        want = np.zeros_like(base)
        oracle_lib.oracle_modexp_batch(base.ctypes.data_as(ctypes.c_void_p), exp.ctypes.data_as(ctypes.c_void_p), mod.ctypes.data_as(ctypes.c_void_p),
                                       None, want.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(count), k, k, threads)
        assert np.array_equal(out, want), bits


def test_block_partitioned_squaring_is_bit_identical(engine, pkg):
    """csrc/sqr.cuh (opt-in: fewer multiply-accumulates, measured slower on B200): same residues as the mont_mul(a, a) path and
    as Python integers, for every lane-group width the batch entry point offers, including moduli with leading zero limbs"""
    rng = random.Random(0x5A5A)
    try:
        for bits, tpis in ((2048, (4, 8, 16, 32)), (1024, (4, 8, 16)), (4096, (8, 16, 32))):
            mods = [rng.getrandbits(bits) | (1 << (bits - 1)) | 1 for _ in range(40)] + [(1 << 700) + 1, 3, (1 << bits) - 1]
            bases = [rng.getrandbits(bits) % m for m in mods]
            exps = [rng.getrandbits(600) for _ in mods]
            want = [pow(b, e, m) for b, e, m in zip(bases, exps, mods)]
            for tpi in tpis:
                engine.set_tpi(bits, tpi)
                for sqr in (1, 0):
                    engine.set_option("sqr", sqr)
                    got, st = engine.mod_pow(bases, exps, mods, mod_bits=bits)
                    assert not st.any() and got == want, (bits, tpi, sqr)
    finally:
        engine.set_option("sqr", 0)
        for bits in (1024, 2048, 4096):
            engine.set_tpi(bits, 0)
