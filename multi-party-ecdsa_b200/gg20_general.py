"""GG20 offline stage for ANY signing set over the batched L0-L2 calls: the size-generic driver next to the fused L3 call
(`tecdsa_gg20_offline_batch`, which is specialised to the t = 1 / two-signer work unit of SURVEY.md section 8).  Same functions, in the
round order of /root/reference/src/protocols/multi_party_ecdsa/gg_2020/state_machine/sign/rounds.rs:68-636, with the reference's
index convention (`ind = if j < i {j} else {j+1}`); every round is a handful of batch calls over all (session, signer[, peer])
instances.  Element u = session * ttag + position; pair t = u * (ttag - 1) + j is "element u towards its j-th peer".
No arithmetic happens on the host — only index bookkeeping and equality tests.  Oracle: oracle/gg20_general_oracle.py."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import Engine, limbs_to_ints
from . import gg18, gg20
from .gg20 import KeySets, unpack_point

ST_INVALID_KEY, ST_PDL, ST_PHASE5, ST_PHASE6, ST_PROOF, ST_COMMIT = 2, 6, 7, 8, 10, 11


def _pts_of(a: np.ndarray):
    return [unpack_point(v) for v in limbs_to_ints(a)]


def offline_batch(eng: Engine, keys: KeySets, ttag: int, key_rows: Sequence[int], all_rows: Sequence[Sequence[int]], w: Sequence[int],
                  g_w: Sequence, y: Sequence, rnd: Dict[str, list], messages: bool = False) -> Dict[str, list]:
    """`OfflineStage` Round0..Round6 for sessions of `ttag` signers.
    key_rows[u]  key row (Paillier key + N~/h1/h2 setup) of element u in `keys`;
    all_rows[u]  the key rows of ALL keygen parties of its key, in keygen order (the statements `MessageA::a` proves against);
    w[u], g_w[u] the Lagrange-weighted share and its public image (`SignKeys::g_w_vec`), y[u] the public key;
    rnd          per element: gamma, k, blind, r_k, l, ped_s1, ped_s2, heg_s1, heg_s2 and alice[u] = n tuples (alpha, beta, gamma, ro);
                 per pair: beta_tag_gamma, r_gamma, nonce_gamma_b, nonce_gamma_beta, beta_tag_w, r_w, nonce_w_b, nonce_w_beta and
                 pdl[t] = (alpha, beta, rho, gamma).
    -> dict(status, R, sigma, k, T): `CompletedOfflineStage` per element (T = its own T_i); status = the reference's first error.
    With messages=True also "messages": per element the list of `Msg<OfflineProtocolMessage>` documents it sends (serde-JSON shape of
    state_machine/sign.rs:478-490 through wire.py; sender / receiver = 1-based signer positions), for sessions that completed."""
    U, P1 = len(key_rows), ttag - 1
    status = np.zeros(U, np.uint8)
    sess = lambda u: u // ttag * ttag
    peer = [sess(u) + (j if j < u % ttag else j + 1) for u in range(U) for j in range(P1)]       # pair t -> the other element
    me = [u for u in range(U) for _ in range(P1)]
    back = {(me[t], peer[t]): t for t in range(U * P1)}                                          # (from, to) -> pair index

    stopped = set()                      # sessions in which some party failed in an EARLIER round: nothing more is recorded for them

    def fail(elems, code):
        for u in elems:
            if not status[u] and sess(u) not in stopped:
                status[u] = code

    def end_round():
        stopped.update(sess(u) for u in range(U) if status[u])

    def session_ok(u):
        return not status[sess(u):sess(u) + ttag].any()

    k, gamma = list(rnd["k"]), list(rnd["gamma"])
    # ---- Round 0: commitment to g^gamma, MessageA with one range proof per keygen party
    g_gamma = eng.secp_mul(None, gamma)
    com = gg20.hash_commitment(eng, g_gamma, rnd["blind"])
    c_a, proofs = gg20.mta_message_a(eng, keys, list(key_rows), [list(r) for r in all_rows], k, rnd["r_k"], rnd["alice"])
    # ---- Round 1: element `me` answers peer's MessageA twice (b = gamma, b = w) under the PEER's Paillier key
    a_rows = [key_rows[v] for v in peer]
    a_st = [list(all_rows[u]) for u in me]
    a_ca = [c_a[v] for v in peer]
    a_pf = {f: [proofs[f][v] for v in peer] for f in proofs}
    cb_g, bp_g, btp_g, beta_g, st = gg20.mta_message_b(eng, keys, a_rows, a_st, [gamma[u] for u in me], a_ca, a_pf, rnd["r_gamma"], rnd["beta_tag_gamma"],
                                                        rnd["nonce_gamma_b"], rnd["nonce_gamma_beta"])
    fail([me[t] for t in range(U * P1) if st[t]], ST_INVALID_KEY)
    cb_w, bp_w, btp_w, beta_w, st = gg20.mta_message_b(eng, keys, a_rows, a_st, [w[u] for u in me], a_ca, a_pf, rnd["r_w"], rnd["beta_tag_w"],
                                                        rnd["nonce_w_b"], rnd["nonce_w_beta"])
    fail([me[t] for t in range(U * P1) if st[t]], ST_INVALID_KEY)
    end_round()
    # ---- Round 2: alice = peer[t] opens what bob = me[t] sent her
    al_rows = [key_rows[v] for v in peer]
    al_k = [k[v] for v in peer]
    alpha_g, _, st = gg20.mta_get_alpha(eng, keys, al_rows, al_k, cb_g, bp_g, btp_g)
    fail([peer[t] for t in range(U * P1) if st[t]], ST_INVALID_KEY)
    alpha_w, _, st = gg20.mta_get_alpha(eng, keys, al_rows, al_k, cb_w, bp_w, btp_w)
    fail([peer[t] for t in range(U * P1) if st[t]], ST_INVALID_KEY)
    pk_w = _pts_of(bp_w[:, :16])
    fail([peer[t] for t in range(U * P1) if pk_w[t] != g_w[me[t]]], ST_INVALID_KEY)               # rounds.rs:281
    delta = eng.scalar_op("mul", k, gamma)
    sigma = eng.scalar_op("mul", k, list(w))
    for j in range(P1):
        incoming = [back[(sess(u) + (j if j < u % ttag else j + 1), u)] for u in range(U)]         # the pair in which u is alice of its j-th peer
        delta = eng.scalar_op("add", delta, [alpha_g[t] for t in incoming])
        sigma = eng.scalar_op("add", sigma, [alpha_w[t] for t in incoming])
        delta = eng.scalar_op("add", delta, [beta_g[u * P1 + j] for u in range(U)])
        sigma = eng.scalar_op("add", sigma, [beta_w[u * P1 + j] for u in range(U)])
    T, t_proof = gg20.pedersen_prove(eng, sigma, rnd["l"], rnd["ped_s1"], rnd["ped_s2"])
    end_round()
    # ---- Round 3: every T_i's Pedersen proof, delta^-1
    st = gg20.pedersen_verify(eng, T, t_proof)
    tot = [0] * U
    for j in range(ttag):
        tot = eng.scalar_op("add", tot, [delta[sess(u) + j] for u in range(U)])
    dinv = eng.scalar_op("inv", tot)
    for u in range(U):
        if dinv[u] is None or st[sess(u):sess(u) + ttag].any():
            fail([u], ST_PROOF)
    dinv = [d or 1 for d in dinv]
    end_round()
    # ---- Round 4: decommitments, R, R_dash, one PDL-with-slack proof per peer
    pk_g = _pts_of(bp_g[:, :16])
    pks = [[g_gamma[u] if sess(u) + j == u else pk_g[back[(sess(u) + j, u)]] for j in range(ttag)] for u in range(U)]
    R, st = gg18.phase4(eng, ttag, dinv, pks, g_gamma, rnd["blind"], com)
    fail([u for u in range(U) if st[u]], ST_COMMIT)
    Rs = [p if p is not None else g_gamma[u] for u, p in enumerate(R)]
    R_dash = eng.secp_mul(Rs, k)
    pa, pb, pr, pg = zip(*rnd["pdl"])
    pdl = gg20.pdl_prove(eng, keys, [key_rows[u] for u in me], [key_rows[v] for v in peer], [k[u] for u in me], [rnd["r_k"][u] for u in me],
                         [c_a[u] for u in me], [R_dash[u] for u in me], [Rs[u] for u in me], list(pa), list(pb), list(pr), list(pg))
    end_round()
    # ---- Round 5: every proof of every signer, sum of R_dash, S_i with its consistency proof
    st = gg20.pdl_verify(eng, keys, [key_rows[u] for u in me], [key_rows[v] for v in peer], [c_a[u] for u in me], [R_dash[u] for u in me], [Rs[u] for u in me],
                         pdl["z"], pdl["u1"], pdl["u2"], pdl["u3"], pdl["s1"], pdl["s2"], pdl["s3"])
    bad_sessions = {sess(me[t]) for t in range(U * P1) if st[t]}
    fail([u for u in range(U) if sess(u) in bad_sessions], ST_PDL)
    gen = eng.secp_mul(None, [1])[0]
    acc = [R_dash[sess(u)] for u in range(U)]
    for j in range(1, ttag):
        acc = eng.point_add(acc, [R_dash[sess(u) + j] for u in range(U)])
    fail([u for u in range(U) if acc[u] != gen], ST_PHASE5)
    S = eng.secp_mul(Rs, sigma)
    heg = gg20.heg_prove(eng, Rs, _pts_of(T), S, rnd["l"], sigma, rnd["heg_s1"], rnd["heg_s2"])
    end_round()
    # ---- Round 6
    st = gg20.heg_verify(eng, Rs, _pts_of(T), S, heg)
    acc = [S[sess(u)] for u in range(U)]
    for j in range(1, ttag):
        acc = eng.point_add(acc, [S[sess(u) + j] for u in range(U)])
    for u in range(U):
        if st[sess(u):sess(u) + ttag].any() or acc[u] != y[u]:
            fail([u], ST_PHASE6)
    # a session stops in the round where its first party failed: the other parties of that session produce nothing either
    out_R = [R[u] if session_ok(u) else None for u in range(U)]
    out = {"status": status, "R": out_R, "sigma": sigma, "k": k, "T": _pts_of(T), "session_ok": [session_ok(u) for u in range(U)]}
    if messages:
        from . import wire
        E = wire.DEFAULT
        Tp = _pts_of(T)
        docs = []
        for u in range(U):
            if not session_ok(u):
                docs.append([])
                continue
            me_pos = u % ttag + 1
            n_st = len(all_rows[u])
            rp = [{f: proofs[f][u][x] for f in ("z", "e", "s", "s1", "s2")} for x in range(n_st)]
            ped = wire.pedersen_proof(t_proof[u]); ped["com"] = E.point(Tp[u])
            m = [wire.msg(me_pos, None, wire.offline_message("M1", [wire.message_a(c_a[u], rp), wire.sign_broadcast_phase1(com[u])]))]
            for j in range(P1):
                t = u * P1 + j
                m.append(wire.msg(me_pos, peer[t] % ttag + 1, wire.offline_message("M2", [wire.message_b(cb_g[t], bp_g[t], btp_g[t]), wire.message_b(cb_w[t], bp_w[t], btp_w[t])])))
            m.append(wire.msg(me_pos, None, wire.offline_message("M3", [E.scalar(delta[u]), E.point(Tp[u]), ped])))
            m.append(wire.msg(me_pos, None, wire.offline_message("M4", wire.sign_decommit_phase1(rnd["blind"][u], g_gamma[u]))))
            pl = [wire.pdl_proof({f: pdl[f][u * P1 + j] for f in ("z", "u1", "u2", "u3", "s1", "s2", "s3")}) for j in range(P1)]
            m.append(wire.msg(me_pos, None, wire.offline_message("M5", [E.point(R_dash[u]), pl])))
            m.append(wire.msg(me_pos, None, wire.offline_message("M6", [E.point(S[u]), wire.heg_proof(heg[u])])))
            docs.append(m)
        out["messages"] = docs
    return out
