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
