"""GPU parity tests of the L0 / L1 / L2 batch entry points against the oracle."""
import random

import numpy as np
import pytest

from oracle import gg20_oracle as o
from oracle.sampling import Drbg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [2048, 4096])
def test_modmul_modinv(engine, bits):
    rng = random.Random(bits + 1)
    n = 70
    mods = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(n)]
    mods[0] = 3; mods[1] = (1 << bits) - 1; mods[2] = rng.getrandbits(bits - 2) | 1         # small / all-ones / short modulus
    a = [rng.getrandbits(bits) for _ in range(n)]
    b = [rng.getrandbits(bits) for _ in range(n)]
    assert engine.mod_mul(a, b, mods, bits) == [x * y % m for x, y, m in zip(a, b, mods)]
    # inverses: mix invertible and non-invertible operands (p*q moduli, multiples of p)
    p1, q1 = 1000003, 998244353
    mods[5] = p1 * q1; a[5] = p1 * 12345                      # gcd != 1
    mods[6] = p1 * q1; a[6] = 0
    a[7] = 1
    got = engine.mod_inv(a, mods, bits)
    for x, m, g in zip(a, mods, got):
        try:
            want = pow(x, -1, m)
        except ValueError:
            want = None
        assert g == want, (x % 1000, m % 1000)


def test_secp_mul_1m_style(engine):
    """BASELINE.json configs[2] shape (scaled to what the oracle checks in seconds): generator and
    variable-base scalar multiplications vs the oracle (itself pinned to OpenSSL)."""
    rng = random.Random(3)
    ks = [1, 2, o.Q - 1, 0] + [rng.randrange(1, o.Q) for _ in range(60)]
    got = engine.secp_mul(None, ks)
    assert got == [o.pt_mul(o.G, k) for k in ks]
    pts = [o.pt_mul(o.G, rng.randrange(1, o.Q)) for _ in range(len(ks))]
    pts[3] = None
    got = engine.secp_mul(pts, ks)
    assert got == [o.pt_mul(p, k) for p, k in zip(pts, ks)]
    # size-independent property at a larger batch: (k1 + k2) G == k1 G + k2 G via three batched calls
    n = 4096
    k1 = [rng.randrange(1, o.Q) for _ in range(n)]
    k2 = [rng.randrange(1, o.Q) for _ in range(n)]
    s = engine.secp_mul(None, [(x + y) % o.Q for x, y in zip(k1, k2)])
    a = engine.secp_mul(None, k1)
    twice = engine.secp_mul(a, k2)                       # k2 * (k1 G)
    both = engine.secp_mul(None, [x * y % o.Q for x, y in zip(k1, k2)])
    assert twice == both
    for i in range(0, n, 97):
        assert s[i] == o.pt_add(a[i], o.pt_mul(o.G, k2[i]))


def test_paillier_ops(engine, pkg, keyset):
    from mpecdsa_b200 import gg20
    rng = random.Random(4)
    ns = [k.dk.p * k.dk.q for k in keyset]
    n = 48
    idx = [rng.randrange(3) for _ in range(n)]
    m = [rng.randrange(ns[i]) for i in idx]
    r = [rng.randrange(1, ns[i]) for i in idx]
    c = engine.paillier_encrypt(ns, idx, m, r)
    eks = [o.EncryptionKey(x, x * x) for x in ns]
    assert c == [o.paillier_encrypt(eks[i], mm, rr) for i, mm, rr in zip(idx, m, r)]
    k = [rng.randrange(o.Q) for _ in range(n)]
    ck = engine.paillier_mul(ns, idx, c, k)
    assert ck == [pow(cc, kk, eks[i].nn) for i, cc, kk in zip(idx, c, k)]
    cs = engine.paillier_add(ns, idx, c, ck)
    assert cs == [x * y % eks[i].nn for i, x, y in zip(idx, c, ck)]
    ks = gg20.KeySets(engine, [keyset])
    dec = engine.paillier_decrypt(ks.handle, idx, cs)
    assert dec == [(mm + mm * kk) % ns[i] for i, mm, kk in zip(idx, m, k)]
    assert dec == [o.paillier_decrypt(keyset[i].dk, x) for i, x in zip(idx, cs)]
    ks.free()


def test_paillier_nadic_edge_cases(engine):
    """The N^2 jobs run in N-adic form (csrc/nadic.cuh); the value must equal plain 4096-bit arithmetic for ANY odd N
    and ANY 4096-bit operand: short moduli, N = 3, all-ones N, ciphertexts >= N^2, zero / one operands and exponents."""
    rng = random.Random(77)
    B = 2048
    ns = [3, (1 << B) - 1, rng.getrandbits(1500) | 1, rng.getrandbits(B - 1) | 1 | (1 << (B - 2)), (1 << (B - 1)) + 1, 5 ** 800]
    ns += [rng.getrandbits(B) | 1 | (1 << (B - 1)) for _ in range(10)]
    n = 96
    idx = [i % len(ns) for i in range(n)]
    c = [rng.getrandbits(4096) for _ in range(n)]                 # mostly >= N^2: reduced like BigInt::mod_pow does
    k = [rng.getrandbits(2048) for _ in range(n)]
    c[0] = 0; c[1] = 1; k[2] = 0; k[3] = 1; c[4] = (1 << 4096) - 1; k[5] = (1 << 2048) - 1
    for j in range(6, 22):
        c[j] = ns[idx[j]] ** 2 - 1 - (j & 1) * rng.getrandbits(40)   # just below N^2: -1 and neighbours
    got = engine.paillier_mul(ns, idx, c, k, k_limbs=64)
    assert got == [pow(cc, kk, ns[i] ** 2) for i, cc, kk in zip(idx, c, k)]
    c2 = [rng.getrandbits(4096) for _ in range(n)]
    got = engine.paillier_add(ns, idx, c, c2)
    assert got == [x * y % ns[i] ** 2 for i, x, y in zip(idx, c, c2)]
    m = [rng.getrandbits(2048) for _ in range(n)]
    r = [rng.getrandbits(2048) for _ in range(n)]
    got = engine.paillier_encrypt(ns, idx, m, r)
    assert got == [(1 + mm * ns[i]) * pow(rr, ns[i], ns[i] ** 2) % ns[i] ** 2 for i, mm, rr in zip(idx, m, r)]


def test_alice_proof_generate_verify_batch(engine, pkg, keyset):
    """BASELINE.json configs[3] shape at a size the oracle checks in seconds: proofs generated on the
    GPU are byte-identical to the oracle's, verify on the GPU and under the oracle; tampered ones are
    rejected with the reference's failure reason."""
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    rng = Drbg(0xB2000004, "alice-batch")
    n = 24
    q3 = o.Q ** 3
    ek_row = [i % 3 for i in range(n)]
    st_row = [(i // 3) % 3 for i in range(n)]
    a, r, cph, al, be, ga, ro = [], [], [], [], [], [], []
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        a.append(rng.scalar()); r.append(rng.unit_mod(ek.n))
        cph.append(o.paillier_encrypt(ek, a[-1], r[-1]))
        al.append(rng.below(q3)); be.append(rng.unit_mod(ek.n)); ga.append(rng.below(q3 * st.N)); ro.append(rng.below(o.Q * st.N))
    pf = gg20.alice_proof_generate(engine, ks, ek_row, st_row, a, cph, r, al, be, ga, ro)
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        want = o.alice_proof_generate(a[i], cph[i], ek, st, r[i], al[i], be[i], ga[i], ro[i])
        assert (pf["z"][i], pf["e"][i], pf["s"][i], pf["s1"][i], pf["s2"][i]) == (want.z, want.e, want.s, want.s1, want.s2)
        assert o.alice_proof_verify(want, cph[i], ek, st)
    st_ok = gg20.alice_proof_verify(engine, ks, ek_row, st_row, cph, pf["z"], pf["e"], pf["s"], pf["s1"], pf["s2"])
    assert not st_ok.any()
    # tampering: wrong ciphertext, out-of-range s1, non-invertible z
    bad_c = list(cph); bad_c[0] += 1
    bad_s1 = list(pf["s1"]); bad_s1[1] = q3 + 1
    bad_z = list(pf["z"]); bad_z[2] = 0
    st1 = gg20.alice_proof_verify(engine, ks, ek_row, st_row, bad_c, bad_z, pf["e"], pf["s"], bad_s1, pf["s2"])
    assert st1[0] == pkg.ST_HASH_MISMATCH and st1[1] == pkg.ST_RANGE and st1[2] == pkg.ST_NOT_INVERTIBLE and not st1[3:].any()
    # fields wider than their ABI slot (or negative) reject THAT proof and leave the rest of the batch alone
    wide_s1 = list(pf["s1"]); wide_s1[3] = 1 << 900
    wide_s2 = list(pf["s2"]); wide_s2[4] = 1 << 2944
    wide_z = list(pf["z"]); wide_z[5] = pf["z"][5] + (1 << 2048); wide_z[6] = -1
    st2 = gg20.alice_proof_verify(engine, ks, ek_row, st_row, cph, wide_z, pf["e"], pf["s"], wide_s1, wide_s2)
    assert st2[3] == pkg.ST_RANGE and list(st2[4:7]) == [pkg.ST_HASH_MISMATCH] * 3 and not st2[:3].any() and not st2[7:].any()
    ks.free()


def test_pdl_prove_verify_batch(engine, pkg, keyset):
    """mirrors zk_pdl_with_slack/test.rs: accept (:11-68) and the x+1 soundness negative (:70-129), batched; proof
    bytes identical to the oracle's."""
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    rng = Drbg(14, "pdl-batch")
    n = 12
    q3 = o.Q ** 3
    ek_row = [i % 3 for i in range(n)]
    st_row = [(i + 1) % 3 for i in range(n)]
    x, r, c, Qs, Gs, al, be, rh, ga = [], [], [], [], [], [], [], [], []
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]; st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        x.append(rng.scalar()); r.append(rng.unit_mod(ek.n)); c.append(o.paillier_encrypt(ek, x[-1], r[-1]))
        Gs.append(o.pt_mul(o.G, rng.scalar())); Qs.append(o.pt_mul(Gs[-1], x[-1]))
        al.append(rng.below(q3)); be.append(1 + rng.below(ek.n - 2)); rh.append(rng.below(o.Q * st.N)); ga.append(rng.below(q3 * st.N))
    pf = gg20.pdl_prove(engine, ks, ek_row, st_row, x, r, c, Qs, Gs, al, be, rh, ga)
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]; st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        w = o.pdl_prove(x[i], r[i], c[i], ek, Qs[i], Gs[i], st.g, st.ni, st.N, al[i], be[i], rh[i], ga[i])
        assert (pf["z"][i], pf["u1"][i], pf["u2"][i], pf["u3"][i], pf["s1"][i], pf["s2"][i], pf["s3"][i]) == (w.z, w.u1, w.u2, w.u3, w.s1, w.s2, w.s3)
    st_ok = gg20.pdl_verify(engine, ks, ek_row, st_row, c, Qs, Gs, pf["z"], pf["u1"], pf["u2"], pf["u3"], pf["s1"], pf["s2"], pf["s3"])
    assert not st_ok.any()
    # soundness negative: ciphertext of x+1 with a proof for x; and a non-invertible z
    c_bad = list(c)
    ek0 = keyset[0].paillier_key_vec[ek_row[0]]
    c_bad[0] = o.paillier_encrypt(ek0, x[0] + 1, r[0])
    z_bad = list(pf["z"]); z_bad[1] = 0
    st1 = gg20.pdl_verify(engine, ks, ek_row, st_row, c_bad, Qs, Gs, z_bad, pf["u1"], pf["u2"], pf["u3"], pf["s1"], pf["s2"], pf["s3"])
    assert st1[0] == pkg.ST_PDL_VERIFY and st1[1] == pkg.ST_PDL_VERIFY and not st1[2:].any()
    ks.free()


@pytest.mark.parametrize("check", [False, True])
def test_bob_proofs_batch(engine, pkg, keyset, check):
    """mirrors range_proofs.rs `bob_zkp` (:636-709): BobProof (MtA) and BobProofExt (MtAwc), batched, byte-identical
    to the oracle and accepted by both verifiers; tampered proofs rejected."""
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    rng = Drbg(15 + check, "bob-batch")
    n = 9
    q3 = o.Q ** 3
    ek_row = [i % 3 for i in range(n)]
    st_row = [(i + 2) % 3 for i in range(n)]
    cols = {k: [] for k in ("a_enc", "mta", "b", "bp", "r", "al", "be", "ga", "ro", "rp", "si", "ta")}
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]; st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        a, b = rng.scalar(), rng.scalar()
        enc_a = o.paillier_encrypt(ek, a, rng.unit_mod(ek.n))
        bp, r = rng.below(ek.n), rng.unit_mod(ek.n)
        mta = o.paillier_add(ek, o.paillier_mul(ek, enc_a, b), o.paillier_encrypt(ek, bp, r))
        vals = (enc_a, mta, b, bp, r, rng.below(q3), rng.unit_mod(ek.n), rng.below(o.Q ** 2 * ek.n), rng.below(o.Q * st.N), rng.below(q3 * st.N),
                rng.below(o.Q * st.N), rng.below(q3 * st.N))
        for k, v in zip(cols, vals):
            cols[k].append(v)
    pf = gg20.bob_proof_generate(engine, ks, ek_row, st_row, check, cols["a_enc"], cols["mta"], cols["b"], cols["bp"], cols["r"], cols["al"],
                                 cols["be"], cols["ga"], cols["ro"], cols["rp"], cols["si"], cols["ta"])
    Xs = [o.pt_mul(o.G, b) for b in cols["b"]]
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]; st = keyset[0].h1_h2_n_tilde_vec[st_row[i]]
        w, u = o.bob_proof_generate(cols["a_enc"][i], cols["mta"][i], cols["b"][i], cols["bp"][i], ek, st, cols["r"][i], check, cols["al"][i],
                                    cols["be"][i], cols["ga"][i], cols["ro"][i], cols["rp"][i], cols["si"][i], cols["ta"][i])
        got = tuple(pf[k][i] for k in ("t", "z", "e", "s", "s1", "s2", "t1", "t2"))
        assert got == (w.t, w.z, w.e, w.s, w.s1, w.s2, w.t1, w.t2), i
        if check:
            assert pf["u"][i] == u
            assert o.bob_proof_ext_verify(w, u, cols["a_enc"][i], cols["mta"][i], ek, st, Xs[i])
        else:
            assert o.bob_proof_verify(w, cols["a_enc"][i], cols["mta"][i], ek, st)
    st_ok = gg20.bob_proof_verify(engine, ks, ek_row, st_row, cols["a_enc"], cols["mta"], pf, Xs if check else None, pf["u"] if check else None)
    assert not st_ok.any()
    bad = dict(pf); bad["t1"] = list(pf["t1"]); bad["t1"][0] += 1
    mta_bad = list(cols["mta"]); mta_bad[1] += 1
    st1 = gg20.bob_proof_verify(engine, ks, ek_row, st_row, cols["a_enc"], mta_bad, bad, Xs if check else None, pf["u"] if check else None)
    assert st1[0] == pkg.ST_HASH_MISMATCH and st1[1] == pkg.ST_HASH_MISMATCH and not st1[2:].any()
    if check:
        Xbad = list(Xs); Xbad[2] = o.pt_add(Xs[2], o.G)
        st2 = gg20.bob_proof_verify(engine, ks, ek_row, st_row, cols["a_enc"], cols["mta"], pf, Xbad, pf["u"])
        assert st2[2] in (pkg.ST_HASH_MISMATCH, pkg.ST_PROOF) and not st2[3:].any()
    ks.free()


def test_sigma_proofs_and_hashes_batch(engine, pkg):
    """curv's DLogProof / PedersenProof / HomoELGamalProof / HashCommitment / chain_bigint as stand-alone batches,
    byte-identical to the oracle; tampered proofs rejected."""
    from mpecdsa_b200 import gg20
    rng = Drbg(21, "sigma-batch")
    n = 20
    sk, nonce = [rng.scalar() for _ in range(n)], [rng.scalar() for _ in range(n)]
    pf = gg20.dlog_prove(engine, sk, nonce)
    for i in range(n):
        w = o.dlog_prove(sk[i], nonce[i])
        got = pkg.limbs_to_ints(pf[i:i + 1, :16])[0], pkg.limbs_to_ints(pf[i:i + 1, 16:32])[0], pkg.limbs_to_ints(pf[i:i + 1, 32:])[0]
        assert got == (gg20.pack_point(w.pk), gg20.pack_point(w.pk_t_rand_commitment), w.challenge_response)
    assert not gg20.dlog_verify(engine, pf).any()
    bad = pf.copy(); bad[0, 32] ^= 1; bad[1, 0] ^= 1
    st = gg20.dlog_verify(engine, bad)
    assert st[0] == pkg.ST_PROOF and st[1] == pkg.ST_PROOF and not st[2:].any()
    # Pedersen
    m, r, s1, s2 = ([rng.scalar() for _ in range(n)] for _ in range(4))
    com, ped = gg20.pedersen_prove(engine, m, r, s1, s2)
    for i in range(n):
        w = o.pedersen_prove(m[i], r[i], s1[i], s2[i])
        assert gg20.unpack_point(pkg.limbs_to_ints(com[i:i + 1])[0]) == w.com
        assert pkg.limbs_to_ints(ped[i:i + 1, :8])[0] == w.e and pkg.limbs_to_ints(ped[i:i + 1, 40:48])[0] == w.z1 and pkg.limbs_to_ints(ped[i:i + 1, 48:56])[0] == w.z2
    assert not gg20.pedersen_verify(engine, com, ped).any()
    ped_bad = ped.copy(); ped_bad[2, 40] ^= 1
    assert gg20.pedersen_verify(engine, com, ped_bad)[2] == pkg.ST_PROOF
    # HomoElGamal on the statement shape of phase 6 (party_i.rs:778-799)
    R = [o.pt_mul(o.G, rng.scalar()) for _ in range(n)]
    l, sigma = [rng.scalar() for _ in range(n)], [rng.scalar() for _ in range(n)]
    T = [o.pt_add(o.pt_mul(o.G, s), o.pt_mul(o.H2, x)) for s, x in zip(sigma, l)]
    S = [o.pt_mul(Rp, s) for Rp, s in zip(R, sigma)]
    a, b = [rng.scalar() for _ in range(n)], [rng.scalar() for _ in range(n)]
    heg = gg20.heg_prove(engine, R, T, S, l, sigma, a, b)
    for i in range(n):
        w = o.heg_prove(l[i], sigma[i], R[i], o.H2, o.G, T[i], S[i], a[i], b[i])
        assert (gg20.unpack_point(pkg.limbs_to_ints(heg[i:i + 1, :16])[0]), gg20.unpack_point(pkg.limbs_to_ints(heg[i:i + 1, 16:32])[0]),
                pkg.limbs_to_ints(heg[i:i + 1, 32:40])[0], pkg.limbs_to_ints(heg[i:i + 1, 40:48])[0]) == (w.T, w.A3, w.z1, w.z2)
    assert not gg20.heg_verify(engine, R, T, S, heg).any()
    S_bad = list(S); S_bad[3] = o.pt_add(S[3], o.G)
    assert gg20.heg_verify(engine, R, T, S_bad, heg)[3] == pkg.ST_PROOF
    # hashes
    rows = [(rng.bits(2048), rng.bits(4096) >> (8 * (i % 5)), rng.bits(250), 0 if i == 0 else rng.bits(64)) for i in range(n)]
    assert gg20.sha256_bigints(engine, rows, [64, 128, 8, 4]) == [o.sha256_bigints(list(row)) for row in rows]
    blinds = [rng.bits(256) >> (i * 7) for i in range(n)]
    assert gg20.hash_commitment(engine, R, blinds) == [o.hash_commitment(o.bn_from_bytes(o.pt_compress(p)), bl) for p, bl in zip(R, blinds)]


def test_mta_messages_batch(engine, pkg, keyset):
    """mirrors mta/test.rs:5-18 batched and at message level: MessageA / MessageB bytes identical to the oracle,
    alpha + beta == a*b (mod q); a tampered range proof makes MessageB::b fail with InvalidKey (mta/mod.rs:123-131)."""
    from mpecdsa_b200 import gg20
    ks = gg20.KeySets(engine, [keyset])
    rng = Drbg(31, "mta-batch")
    n = 6
    q3 = o.Q ** 3
    ek_row = [i % 3 for i in range(n)]                      # Alice's key row
    st_rows = [[0, 1, 2]] * n                               # the full h1_h2_n_tilde_vec, as Round0 passes it (sign/rounds.rs:85)
    stmts = keyset[0].h1_h2_n_tilde_vec
    a, r, pr = [], [], []
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        a.append(rng.scalar()); r.append(rng.below(ek.n))
        pr.append([(rng.below(q3), rng.unit_mod(ek.n), rng.below(q3 * st.N), rng.below(o.Q * st.N)) for st in stmts])
    c, proofs = gg20.mta_message_a(engine, ks, ek_row, st_rows, a, r, pr)
    m_as = []
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        m_a = o.message_a(a[i], ek, r[i], stmts, pr[i])
        m_as.append(m_a)
        assert c[i] == m_a.c
        for x, pf in enumerate(m_a.range_proofs):
            assert tuple(proofs[k][i][x] for k in ("z", "e", "s", "s1", "s2")) == (pf.z, pf.e, pf.s, pf.s1, pf.s2)
    b = [rng.scalar() for _ in range(n)]
    rand_b, beta_tag = [], []
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        rand_b.append(rng.below(ek.n)); beta_tag.append(rng.below(ek.n))
    nb, nbt = [rng.scalar() for _ in range(n)], [rng.scalar() for _ in range(n)]
    c_b, bp, btp, beta, st = gg20.mta_message_b(engine, ks, ek_row, st_rows, b, c, proofs, rand_b, beta_tag, nb, nbt)
    assert not st.any()
    for i in range(n):
        ek = keyset[0].paillier_key_vec[ek_row[i]]
        m_b, beta_w = o.message_b(b[i], ek, m_as[i], rand_b[i], beta_tag[i], stmts, nb[i], nbt[i])
        assert c_b[i] == m_b.c and beta[i] == beta_w
        assert pkg.limbs_to_ints(bp[i:i + 1, 32:])[0] == m_b.b_proof.challenge_response
        assert gg20.unpack_point(pkg.limbs_to_ints(btp[i:i + 1, :16])[0]) == m_b.beta_tag_proof.pk
    alpha, plain, st2 = gg20.mta_get_alpha(engine, ks, ek_row, a, c_b, bp, btp)
    assert not st2.any()
    for i in range(n):
        assert (alpha[i] + beta[i]) % o.Q == a[i] * b[i] % o.Q
        assert plain[i] == o.paillier_decrypt(keyset[ek_row[i]].dk, c_b[i])
    # tampered proof of instance 1 (statement 2): Bob refuses; wrong `a` at Alice's check: InvalidKey
    bad = {k: [list(v) for v in proofs[k]] for k in proofs}
    bad["s"][1][2] += 1
    _, _, _, _, st3 = gg20.mta_message_b(engine, ks, ek_row, st_rows, b, c, bad, rand_b, beta_tag, nb, nbt)
    assert st3[1] == pkg.ST_INVALID_KEY and not np.delete(st3, 1).any()
    a_bad = list(a); a_bad[0] = (a[0] + 1) % o.Q
    _, _, st4 = gg20.mta_get_alpha(engine, ks, ek_row, a_bad, c_b, bp, btp)
    assert st4[0] == pkg.ST_INVALID_KEY and not st4[1:].any()
    ks.free()


def test_scalar_point_bigint_surface(engine, pkg):
    """The rest of the Scalar / Point / BigInt surface the protocol code calls (include/tecdsa_b200.h L0), against Python
    integers, the oracle's secp256k1 (itself pinned to OpenSSL) and hashlib."""
    import hashlib
    from math import gcd
    rng = random.Random(0xB2E0)
    Qn, Pn = o.Q, o.P
    a = [rng.randrange(Qn) for _ in range(40)] + [0, 1, Qn - 1]
    b = [rng.randrange(Qn) for _ in range(40)] + [Qn - 1, 0, Qn - 1]
    assert engine.scalar_op("mul", a, b) == [x * y % Qn for x, y in zip(a, b)]
    assert engine.scalar_op("add", a, b) == [(x + y) % Qn for x, y in zip(a, b)]
    assert engine.scalar_op("sub", a, b) == [(x - y) % Qn for x, y in zip(a, b)]
    assert engine.scalar_op("inv", a) == [pow(x, -1, Qn) if x else None for x in a]
    big = [rng.getrandbits(2048) for _ in range(8)] + [Qn, Qn + 5, 0, (1 << 2048) - 1]
    assert engine.scalar_from_bigint(big, 64) == [x % Qn for x in big]
    pts = [o.pt_mul(o.G, rng.randrange(1, Qn)) for _ in range(12)]
    A = pts + [None, pts[0], pts[1], None]
    B = pts[1:] + pts[:1] + [pts[3], pts[0], o.pt_neg(pts[1]), None]
    assert engine.point_add(A, B) == [o.pt_add(x, y) for x, y in zip(A, B)]
    assert engine.point_add(A, B, subtract=True) == [o.pt_sub(x, y) for x, y in zip(A, B)]
    enc = engine.point_compress(pts)
    assert enc == [o.pt_compress(p) for p in pts]
    assert engine.point_compress([None]) == [bytes(33)]
    bad_x = next(x for x in range(2, 100) if pow((x ** 3 + 7) % Pn, (Pn - 1) // 2, Pn) != 1)
    bogus = [b"\x04" + enc[0][1:], b"\x02" + Pn.to_bytes(32, "big"), b"\x02" + bad_x.to_bytes(32, "big")]
    assert engine.point_decompress(enc + bogus) == pts + [None, None, None]
    # e * a + alpha over the integers (range_proofs.rs:87-88)
    e, x, al = [rng.getrandbits(256) for _ in range(6)], [rng.randrange(Qn) for _ in range(6)], [rng.randrange(Qn ** 3) for _ in range(6)]
    assert engine.wide_muladd(e, x, al, 8, 8, 24, 28) == [p * q + r for p, q, r in zip(e, x, al)]
    # SampleFromMultiplicativeGroup acceptance: r < N and gcd(r, N) == 1
    p1, p2 = 0xFFFFFFFFFFFFFFC5, 2 ** 89 - 1
    N = [((rng.getrandbits(2048) | (1 << 2047) | 1) // (p1 * p2)) * p1 * p2 | 0 for _ in range(6)]
    N = [n if n & 1 else n + p1 * p2 for n in N]
    r = [rng.randrange(N[0]), p1 * 12345, N[2] + 5, p2, 1, 0]
    want = [x < n and gcd(x, n) == 1 for x, n in zip(r, N)]
    assert want[1:4] == [False, False, False] and want[4] and not want[5]
    assert engine.unit_mod_check(r, N) == want
    msgs = [b"", b"abc", b"ZenGo", bytes(range(256)) * 5, b"x" * 55, b"y" * 56, b"z" * 64]
    assert engine.sha256(msgs) == [hashlib.sha256(m).digest() for m in msgs]
