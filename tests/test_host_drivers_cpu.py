"""Host logic of the composed drivers WITHOUT a GPU: multi-party-ecdsa_b200/gg20_general.py (GG20 offline stage for any signing set) is run
against a stand-in for the batch calls whose every entry point answers with the ORACLE's value for each element.  What is under test is
the part that lives on the host — which element talks to which, pair indexing (`ind = if j < i {j} else {j+1}`), the order in which a
party's sums are formed, when a session stops and which status a party gets — compared with oracle/gg20_general_oracle.py, which runs
the same protocol party by party.  The arithmetic itself is covered by the `-m gpu` parity tests (tests/test_gg20_general.py)."""
import random

import numpy as np
import pytest

from oracle import gg18_oracle as e18
from oracle import gg20_general_oracle as gen
from oracle import gg20_oracle as o

from tests import test_gg20_general as tg

Q, G = o.Q, o.G


def _pt_row(p):
    return np.frombuffer((0 if p is None else p[0] | (p[1] << 256)).to_bytes(64, "little"), dtype="<u4")


def _sc_row(x, k=8):
    return np.frombuffer(int(x).to_bytes(4 * k, "little"), dtype="<u4")


def _row_pt(r):
    v = int.from_bytes(np.ascontiguousarray(r).tobytes(), "little")
    return None if v == 0 else (v & ((1 << 256) - 1), v >> 256)


def _row_int(r):
    return int.from_bytes(np.ascontiguousarray(r).tobytes(), "little")


def _dlog_row(pf):
    return np.concatenate([_pt_row(pf.pk), _pt_row(pf.pk_t_rand_commitment), _sc_row(pf.challenge_response)])


def _row_dlog(r):
    return o.DLogProof(_row_pt(r[:16]), _row_pt(r[16:32]), _row_int(r[32:40]))


class FakeKeys:
    """key row -> (DecryptionKey, EncryptionKey, DLogStatement), like gg20.KeySets over fixture key sets (row = 3 * keyset + party)"""

    def __init__(self, keysets):
        self.rows = []
        for ks in keysets:
            for lk in ks:
                n = lk.dk.p * lk.dk.q
                self.rows.append((lk.dk, o.EncryptionKey(n, n * n), lk.h1_h2_n_tilde_vec[lk.i - 1]))

    def dk(self, r): return self.rows[r][0]
    def ek(self, r): return self.rows[r][1]
    def st(self, r): return self.rows[r][2]


class FakeEngine:
    lib = None                                   # no library: every batch call is stood in

    def secp_mul(self, points, scalars):
        return [o.pt_mul(G if points is None else points[i], s % Q) for i, s in enumerate(scalars)]

    def point_add(self, a, b, subtract=False):
        return [o.pt_sub(x, y) if subtract else o.pt_add(x, y) for x, y in zip(a, b)]

    def scalar_op(self, op, a, b=None):
        if op == "inv":
            return [pow(x % Q, -1, Q) if x % Q else None for x in a]
        f = {"mul": lambda x, y: x * y % Q, "add": lambda x, y: (x + y) % Q, "sub": lambda x, y: (x - y) % Q}[op]
        return [f(x, y) for x, y in zip(a, b)]


def _install(monkeypatch):
    """replace the batch-call wrappers the driver uses by per-element oracle evaluations with the same signatures and return shapes"""
    from mpecdsa_b200 import gg18, gg20

    def hash_commitment(eng, points, blinds):
        return [o.hash_commitment(o.bn_from_bytes(o.pt_compress(p)), b) for p, b in zip(points, blinds)]

    def mta_message_a(eng, keys, ek_row, st_rows, a, r, proof_rand):
        cs, pf = [], {f: [] for f in ("z", "e", "s", "s1", "s2")}
        for i in range(len(a)):
            m = o.message_a(a[i] % Q, keys.ek(ek_row[i]), r[i], [keys.st(x) for x in st_rows[i]], proof_rand[i])
            cs.append(m.c)
            for f in pf:
                pf[f].append([getattr(p, f) for p in m.range_proofs])
        return cs, pf

    def mta_message_b(eng, keys, ek_row, st_rows, b, c_a, proofs, randomness, beta_tag, nonce_b, nonce_beta):
        n = len(b)
        c_b, bp, btp, beta, st = [], np.zeros((n, 40), np.uint32), np.zeros((n, 40), np.uint32), [], np.zeros(n, np.uint8)
        for i in range(n):
            pfs = [o.AliceProof(*(proofs[f][i][x] for f in ("z", "e", "s", "s1", "s2"))) for x in range(len(st_rows[i]))]
            res = o.message_b(b[i] % Q, keys.ek(ek_row[i]), o.MessageA(c_a[i], pfs), randomness[i], beta_tag[i], [keys.st(x) for x in st_rows[i]], nonce_b[i], nonce_beta[i])
            if res is None:
                st[i] = 2; c_b.append(0); beta.append(0)
                continue
            mb, be = res
            c_b.append(mb.c); beta.append(be); bp[i] = _dlog_row(mb.b_proof); btp[i] = _dlog_row(mb.beta_tag_proof)
        return c_b, bp, btp, beta, st

    def mta_get_alpha(eng, keys, dk_row, a, c_b, b_proof, beta_tag_proof):
        n = len(a)
        alpha, plain, st = [], [], np.zeros(n, np.uint8)
        for i in range(n):
            res = o.verify_proofs_get_alpha(o.MessageB(c_b[i], _row_dlog(b_proof[i]), _row_dlog(beta_tag_proof[i])), keys.dk(dk_row[i]), a[i] % Q)
            if res is None:
                st[i] = 2; alpha.append(0); plain.append(0)
            else:
                alpha.append(res[0]); plain.append(res[1])
        return alpha, plain, st

    def pedersen_prove(eng, m, r, s1, s2):
        n = len(m)
        com, pf = np.zeros((n, 16), np.uint32), np.zeros((n, 64), np.uint32)
        for i in range(n):
            p = o.pedersen_prove(m[i] % Q, r[i] % Q, s1[i] % Q, s2[i] % Q)
            com[i] = _pt_row(p.com)
            pf[i, :56] = np.concatenate([_sc_row(p.e), _pt_row(p.a1), _pt_row(p.a2), _sc_row(p.z1), _sc_row(p.z2)])
        return com, pf

    def pedersen_verify(eng, com, pf):
        return np.array([0 if o.pedersen_verify(o.PedersenProof(_row_int(pf[i, :8]), _row_pt(pf[i, 8:24]), _row_pt(pf[i, 24:40]), _row_pt(com[i]),
                                                                 _row_int(pf[i, 40:48]), _row_int(pf[i, 48:56]))) else 10 for i in range(com.shape[0])], np.uint8)

    def pdl_prove(eng, keys, ek_row, st_row, x, r, cipher, Qp, Gp, alpha, beta, rho, gamma):
        out = {f: [] for f in ("z", "u1", "u2", "u3", "s1", "s2", "s3")}
        for i in range(len(x)):
            st = keys.st(st_row[i])
            p = o.pdl_prove(x[i] % Q, r[i], cipher[i], keys.ek(ek_row[i]), Qp[i], Gp[i], st.g, st.ni, st.N, alpha[i], beta[i], rho[i], gamma[i])
            for f in out:
                out[f].append(getattr(p, f))
        return out

    def pdl_verify(eng, keys, ek_row, st_row, cipher, Qp, Gp, z, u1, u2, u3, s1, s2, s3):
        res = []
        for i in range(len(z)):
            st = keys.st(st_row[i])
            pf = o.PDLwSlackProof(z[i], u1[i], u2[i], u3[i], s1[i], s2[i], s3[i])
            res.append(0 if o.pdl_verify(pf, cipher[i], keys.ek(ek_row[i]), Qp[i], Gp[i], st.g, st.ni, st.N) else 6)
        return np.array(res, np.uint8)

    def heg_prove(eng, Gp, D, E, x, r, s1, s2):
        n = len(x)
        pf = np.zeros((n, 48), np.uint32)
        for i in range(n):
            p = o.heg_prove(x[i] % Q, r[i] % Q, Gp[i], o.H2, G, D[i], E[i], s1[i] % Q, s2[i] % Q)
            pf[i] = np.concatenate([_pt_row(p.T), _pt_row(p.A3), _sc_row(p.z1), _sc_row(p.z2)])
        return pf

    def heg_verify(eng, Gp, D, E, pf):
        return np.array([0 if o.heg_verify(o.HomoElGamalProof(_row_pt(pf[i, :16]), _row_pt(pf[i, 16:32]), _row_int(pf[i, 32:40]), _row_int(pf[i, 40:48])),
                                           Gp[i], o.H2, G, D[i], E[i]) else 10 for i in range(pf.shape[0])], np.uint8)

    def phase4(eng, parties, delta_inv, b_proof_pk, g_gamma, blind, com):
        R, st = [], np.zeros(len(delta_inv), np.uint8)
        for u in range(len(delta_inv)):
            s0 = u // parties * parties
            res = e18.phase4(delta_inv[u], b_proof_pk[u], [(blind[v], g_gamma[v]) for v in range(s0, s0 + parties)], com[s0:s0 + parties])
            R.append(res)
            if res is None:
                st[u] = 2
        return R, st

    for name, fn in (("hash_commitment", hash_commitment), ("mta_message_a", mta_message_a), ("mta_message_b", mta_message_b), ("mta_get_alpha", mta_get_alpha),
                     ("pedersen_prove", pedersen_prove), ("pedersen_verify", pedersen_verify), ("pdl_prove", pdl_prove), ("pdl_verify", pdl_verify),
                     ("heg_prove", heg_prove), ("heg_verify", heg_verify)):
        monkeypatch.setattr(gg20, name, fn)
    monkeypatch.setattr(gg18, "phase4", phase4)


def test_general_driver_bookkeeping_matches_party_by_party_oracle(pkg, monkeypatch):
    from mpecdsa_b200 import gg20_general
    from tests.golden import fixtures
    _install(monkeypatch)
    keysets = fixtures.load_all_keysets()[:2]
    keys = FakeKeys(keysets)
    rng = random.Random(0x60D)
    sessions, owner = [], {}
    for kidx, s_l in ((0, [2, 3, 1]), (1, [1, 3, 2])):                   # three signers: every party has two peers, positions permuted
        ks_, rnd = tg._session(rng, keysets[kidx], s_l)
        sessions.append((ks_, s_l, rnd))
        owner.update({id(lk): kidx for lk in ks_})
    args = list(tg._flatten(sessions, lambda lk, j: 3 * owner[id(lk)] + j))
    out = gg20_general.offline_batch(FakeEngine(), keys, *args, messages=True)
    u = 0
    for ks_, s_l, rnd in sessions:
        want = gen.offline_session(ks_, s_l, rnd)
        for p, wv in enumerate(want):
            assert int(out["status"][u]) == wv.status == 0
            assert (out["R"][u], out["sigma"][u], out["k"][u], out["T"][u]) == (wv.R, wv.sigma_i, wv.k_i, wv.t_vec[p])
            u += 1
    tg._check_messages(out, sessions)
    # failures land on the party the reference blames, and the session stops where the state machines would (one-session batches)
    one = list(tg._flatten(sessions[1:], lambda lk, j: 3 * owner[id(lk)] + j))
    # (a) position 0 runs with a wrong share: its two peers stop in round 2 with InvalidKey (rounds.rs:281), nothing is output
    bad = list(one)
    bad[3] = list(one[3]); bad[3][0] = (bad[3][0] + 1) % Q
    res = gg20_general.offline_batch(FakeEngine(), keys, *bad)
    assert list(res["status"]) == [0, 2, 2] and res["R"] == [None, None, None]
    # (b) the decommitment of position 1 does not match its phase-1 commitment -> "bad gamma_i decommit" (party_i.rs:650-687) for the session.
    # (The batched check also covers a party's own entry, which the reference skips; an honest party's own entry always passes, so the
    # difference is unobservable outside this injected fault.)
    from mpecdsa_b200 import gg18
    orig = gg18.phase4

    def phase4_bad(eng, parties, delta_inv, b_proof_pk, g_gamma, blind, com):
        blind = list(blind); blind[1] ^= 1
        return orig(eng, parties, delta_inv, b_proof_pk, g_gamma, blind, com)
    monkeypatch.setattr(gg18, "phase4", phase4_bad)
    res = gg20_general.offline_batch(FakeEngine(), keys, *one)
    assert list(res["status"]) == [11, 11, 11] and res["R"] == [None, None, None]


def _install_gg18(monkeypatch):
    """oracle-backed stand-ins for the GG18 phase wrappers (same signatures and array shapes as multi-party-ecdsa_b200/gg18.py)"""
    from mpecdsa_b200 import gg18

    def local_sig(eng, message, R, k_i, sigma_i):
        return [e18.phase5_local_sig(k_i[u] % Q, message[u], R[u], sigma_i[u] % Q) for u in range(len(k_i))]

    def phase5a(eng, R, s_i, l_i, rho_i, blind, heg_s1, heg_s2, dlog_nonce):
        n = len(s_i)
        out = {"com": np.zeros((n, 8), np.uint32), "decom": np.zeros((n, 48), np.uint32), "heg": np.zeros((n, 48), np.uint32), "dlog": np.zeros((n, 40), np.uint32),
               "status": np.zeros(n, np.uint8)}
        for u in range(n):
            a = e18.phase5a(s_i[u], l_i[u], rho_i[u], R[u], blind[u], heg_s1[u], heg_s2[u], dlog_nonce[u])
            out["com"][u] = _sc_row(a.com)
            out["decom"][u] = np.concatenate([_pt_row(a.V), _pt_row(a.A), _pt_row(a.B)])
            out["heg"][u] = np.concatenate([_pt_row(a.heg.T), _pt_row(a.heg.A3), _sc_row(a.heg.z1), _sc_row(a.heg.z2)])
            out["dlog"][u] = _dlog_row(a.dlog)
        return out

    def _a5(com, decom, blind, heg, dlog, v):
        return e18.Phase5A(_row_int(com[v]), _row_pt(decom[v, :16]), _row_pt(decom[v, 16:32]), _row_pt(decom[v, 32:48]), blind[v],
                           o.HomoElGamalProof(_row_pt(heg[v, :16]), _row_pt(heg[v, 16:32]), _row_int(heg[v, 32:40]), _row_int(heg[v, 40:48])), _row_dlog(dlog[v]))

    def phase5c(eng, parties, R, y, message, rho_i, l_i, blind2, com, decom, blind, heg, dlog):
        n = len(rho_i)
        out = {"com2": np.zeros((n, 8), np.uint32), "decom2": np.zeros((n, 32), np.uint32), "status": np.zeros(n, np.uint8)}
        for u in range(n):
            s0 = u // parties * parties
            others = [_a5(com, decom, blind, heg, dlog, v) for v in range(s0, s0 + parties) if v != u]
            code, res = e18.phase5c(message[u], R[u], y[u], rho_i[u], l_i[u], others, _row_pt(decom[u, :16]), blind2[u])
            out["status"][u] = code
            if res is not None:
                out["com2"][u] = _sc_row(res[0]); out["decom2"][u] = np.concatenate([_pt_row(res[1]), _pt_row(res[2])])
        return out

    def phase5d(eng, parties, decom2, blind2, com2, decom):
        n = decom2.shape[0]
        st = np.zeros(n, np.uint8)
        for u in range(n):
            s0 = u // parties * parties
            rng_ = range(s0, s0 + parties)
            st[u] = e18.phase5d([(_row_pt(decom2[v, :16]), _row_pt(decom2[v, 16:32]), blind2[v]) for v in rng_], [_row_int(com2[v]) for v in rng_],
                                [_row_pt(decom[v, 32:48]) for v in rng_])
        return st

    def output_signature(eng, parties, R, y, message, s_i):
        n = len(s_i)
        r, s, rec, st = [0] * n, [0] * n, np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        for u in range(n):
            s0 = u // parties * parties
            code, sig = e18.output_signature(R[u], y[u], message[u], s_i[s0:s0 + parties])
            st[u] = code
            if sig is not None:
                r[u], s[u], rec[u] = sig
        return r, s, rec, st

    for name, fn in (("local_sig", local_sig), ("phase5a", phase5a), ("phase5c", phase5c), ("phase5d", phase5d), ("output_signature", output_signature)):
        monkeypatch.setattr(gg18, name, fn)
    monkeypatch.setattr(gg18, "_bind", lambda lib: None)


def test_gg18_signing_driver_bookkeeping(pkg, monkeypatch):
    """gg18.sign_batch (phases 1-5 as ~25 batch calls) with the batch calls stood in by the oracle: two three-signer sessions over the
    fixture key set sign, the signature is k^-1 (m + r x) of the opened secrets and verifies; a wrong share is caught in phase 5d"""
    from mpecdsa_b200 import gg18
    from tests.golden import fixtures
    _install(monkeypatch)
    _install_gg18(monkeypatch)
    keyset = fixtures.load_keyset()
    keys = FakeKeys([keyset])
    rng = random.Random(0x1855)
    signers = [[0, 1, 2], [2, 0, 1]]
    parties, U, P1 = 3, 6, 2
    rows = [p for s in signers for p in s]
    w = [o.lagrange_at_zero(p, s) * keyset[p].x_i % Q for s in signers for p in s]
    y = keyset[0].y_sum_s
    msg = [m for _ in signers for m in [rng.getrandbits(256)] * parties]
    sc = lambda n_: [rng.randrange(1, Q) for _ in range(n_)]
    nm = lambda elems: [rng.randrange(1, (keyset[rows[e]].dk.p * keyset[rows[e]].dk.q) >> 1) for e in elems]
    alice = [u for u in range(U) for _ in range(P1)]
    rnd = dict(k=sc(U), gamma=sc(U), blind=sc(U), r_a=nm(range(U)), l=sc(U), rho=sc(U), blind5=sc(U), blind5c=sc(U), heg_s1=sc(U), heg_s2=sc(U), dlog_nonce=sc(U),
               r_b_gamma=nm(alice), r_b_w=nm(alice), nb_gamma=sc(U * P1), nbt_gamma=sc(U * P1), nb_w=sc(U * P1), nbt_w=sc(U * P1), beta_tag_gamma=nm(alice), beta_tag_w=nm(alice))
    out = gg18.sign_batch(FakeEngine(), keys, parties, rows, w, [y] * U, msg, rnd)
    assert list(out["status"]) == [0] * U
    x = sum(o.lagrange_at_zero(p, [0, 1]) * keyset[p].x_i for p in (0, 1)) % Q
    for si in range(2):
        u0 = si * parties
        kk = sum(rnd["k"][u0:u0 + parties]) % Q
        R = o.pt_mul(G, pow(kk, -1, Q))
        s = kk * (msg[u0] + (R[0] % Q) * x) % Q
        s = min(s, Q - s)
        assert all(out["R"][u] == R and (out["r"][u], out["s"][u]) == (R[0] % Q, s) for u in range(u0, u0 + parties))
        assert o.ecdsa_verify(out["r"][u0], out["s"][u0], y, msg[u0])
    bad_w = list(w); bad_w[4] = (bad_w[4] + 1) % Q
    st = gg18.sign_batch(FakeEngine(), keys, parties, rows, bad_w, [y] * U, msg, rnd)["status"]
    assert list(st) == [0, 0, 0, 2, 2, 2]                                # phase 5d: Err(InvalidKey) for every signer of the second session


def test_gg18_keygen_driver_bookkeeping(pkg, monkeypatch):
    """gg18.keygen_batch with the batch calls stood in by oracle/keygen_oracle.py: two groups of three, group key = sum of the y_i, every
    party's x_i = sum of the shares addressed to it (Shamir-consistent), InvalidKey for a group with a bad NiCorrectKeyProof, InvalidSS
    for the receiver of a tampered share only"""
    from mpecdsa_b200 import gg18, gg20, keygen
    from oracle import keygen_oracle as kg
    from tests.golden import fixtures
    _install(monkeypatch)
    keyset = fixtures.load_keyset()
    tamper = {}

    def correct_key_prove(eng, pq, salt=None):
        return [kg.correct_key_proof(o.DecryptionKey(p, q)) for p, q in pq], np.zeros(len(pq), np.uint8)

    def correct_key_verify(eng, n_list, sigma_vecs, salt=None):
        return np.array([0 if kg.correct_key_verify(sv, o.EncryptionKey(n, n * n)) else 10 for n, sv in zip(n_list, sigma_vecs)], np.uint8)

    def vss_share(eng, t, n, polynomials):
        sh, cm = [], []
        for poly in polynomials:
            v, s_ = kg.vss_share(t, n, poly[0], poly[1:])
            sh.append(list(s_)); cm.append(v.commitments)
        for (e, j), d in tamper.items():
            sh[e][j] = (sh[e][j] + d) % Q
        return sh, cm

    def vss_validate_share(eng, commitments, shares, indices):
        return np.array([0 if kg.vss_validate_share(kg.VerifiableSS(len(c) - 1, 0, list(c)), s_, i) else 10 for c, s_, i in zip(commitments, shares, indices)], np.uint8)

    def dlog_prove(eng, sk, nonce):
        return np.stack([_dlog_row(o.dlog_prove(x % Q, r % Q)) for x, r in zip(sk, nonce)])

    def dlog_verify(eng, proofs):
        return np.array([0 if o.dlog_verify(_row_dlog(proofs[i])) else 10 for i in range(proofs.shape[0])], np.uint8)

    for name, fn in (("correct_key_prove", correct_key_prove), ("correct_key_verify", correct_key_verify), ("vss_share", vss_share), ("vss_validate_share", vss_validate_share)):
        monkeypatch.setattr(keygen, name, fn)
    monkeypatch.setattr(gg20, "dlog_prove", dlog_prove)
    monkeypatch.setattr(gg20, "dlog_verify", dlog_verify)
    monkeypatch.setattr(gg18, "_bind", lambda lib: None)
    rng = random.Random(0x18C9)
    t, n, groups = 1, 3, 2
    E = n * groups
    u = [rng.randrange(1, Q) for _ in range(E)]
    polys = [[u[e], rng.randrange(1, Q)] for e in range(E)]
    p_q = [(keyset[e % n].dk.p, keyset[e % n].dk.q) for e in range(E)]
    blind, nonce = [rng.getrandbits(256) for _ in range(E)], [rng.randrange(1, Q) for _ in range(E)]
    out = gg18.keygen_batch(FakeEngine(), t, n, u, p_q, blind, polys, nonce)
    assert list(out["status"]) == [0] * E
    for g in range(groups):
        ys = o.pt_mul(G, sum(u[g * n:(g + 1) * n]) % Q)
        assert all(out["y"][e] == ys for e in range(g * n, (g + 1) * n))
        for pair in ([0, 1], [0, 2], [1, 2]):
            lam = [o.lagrange_at_zero(i, pair) for i in pair]
            assert o.pt_mul(G, sum(l * out["x_i"][g * n + i] for l, i in zip(lam, pair)) % Q) == ys
    tamper[(4, 2)] = 1                                                   # what party 4 sends to party 5 (both of group 1)
    assert list(gg18.keygen_batch(FakeEngine(), t, n, u, p_q, blind, polys, nonce)["status"]) == [0, 0, 0, 0, 0, 12]
    tamper.clear()
    bad_pq = list(p_q); bad_pq[1] = (p_q[1][0], p_q[1][0])                # N = p^2: gcd(N, phi(N)) != 1, no valid NiCorrectKeyProof exists
    st = gg18.keygen_batch(FakeEngine(), t, n, u, bad_pq, blind, polys, nonce)["status"]
    assert list(st[:3]) == [2, 2, 2] and list(st[3:]) == [0, 0, 0]
