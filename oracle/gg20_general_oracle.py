"""CPU restatement of the GG20 offline stage for ANY signing set (t + 1 <= |s_l| <= n signers of a (t, n) key) — the size-generic
companion of oracle/gg20_oracle.py::offline_session, which covers the t = 1, two-signer work unit of SURVEY.md section 8 and carries
the transcript digests.  Same functions, same out-of-tree conventions ([R]), same parity status (UNPINNED); index conventions of
/root/reference/src/protocols/multi_party_ecdsa/gg_2020/state_machine/sign/rounds.rs:68-636 (`ind = if j < i {j} else {j+1}`,
`l_s[x] = s_l[x] - 1`) and gg_2020/party_i.rs:526-848.

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from oracle.gg20_oracle import (G, H2, Q, LocalKey, Point, bn_from_bytes, hash_commitment, heg_prove, heg_verify, lagrange_at_zero,
                                message_a, message_b, pdl_prove, pdl_verify, pedersen_prove, pedersen_verify, pt_add, pt_compress,
                                pt_mul, pt_sub, verify_proofs_get_alpha)

ST_OK, ST_INVALID_KEY, ST_PDL, ST_PHASE5, ST_PHASE6, ST_PROOF, ST_COMMIT = 0, 2, 6, 7, 8, 10, 11


@dataclass
class PartyRandomness:
    """Every value one party's OfflineStage samples; the per-peer lists are ordered by j = 0..ttag-2 (peer position `ind(j)`)."""
    gamma_i: int = 0
    k_i: int = 0
    blind: int = 0
    r_k: int = 0
    alice: List[Tuple[int, int, int, int]] = field(default_factory=list)      # one (alpha, beta, gamma, ro) per keygen party (n of them)
    beta_tag_gamma: List[int] = field(default_factory=list)
    r_gamma: List[int] = field(default_factory=list)
    nonce_gamma_b: List[int] = field(default_factory=list)
    nonce_gamma_beta: List[int] = field(default_factory=list)
    beta_tag_w: List[int] = field(default_factory=list)
    r_w: List[int] = field(default_factory=list)
    nonce_w_b: List[int] = field(default_factory=list)
    nonce_w_beta: List[int] = field(default_factory=list)
    l: int = 0
    ped_s1: int = 0
    ped_s2: int = 0
    pdl: List[Tuple[int, int, int, int]] = field(default_factory=list)        # per peer: alpha, beta, rho, gamma
    heg_s1: int = 0
    heg_s2: int = 0


@dataclass
class PartyResult:
    status: int
    R: Optional[Point] = None
    sigma_i: int = 0
    k_i: int = 0
    t_vec: List[Point] = field(default_factory=list)


def _ind(i: int, j: int) -> int:
    return j if j < i else j + 1


def offline_session(keys: Sequence[LocalKey], s_l: Sequence[int], rnd: Sequence[PartyRandomness]) -> List[PartyResult]:
    """All signers' `OfflineStage` Round0..Round6 in lock step.  keys[p], rnd[p] belong to signer POSITION p (0-based);
    s_l[p] is its 1-based keygen index.  A failing check stops that party with the reference's error (as a status code);
    the session stops at the end of the round in which any party failed, like the state machines would."""
    ttag = len(s_l)
    l_s = [x - 1 for x in s_l]
    res = [PartyResult(ST_OK) for _ in range(ttag)]
    bad = lambda: any(r.status for r in res)
    # ---- Round 0 (rounds.rs:68-104)
    w, gamma, k, g_gamma, com, m_a = [], [], [], [], [], []
    for p in range(ttag):
        lk, r = keys[p], rnd[p]
        w.append(lagrange_at_zero(l_s[p], l_s) * lk.x_i % Q)
        gamma.append(r.gamma_i % Q); k.append(r.k_i % Q)
        g_gamma.append(pt_mul(G, gamma[p]))
        com.append(hash_commitment(bn_from_bytes(pt_compress(g_gamma[p])), r.blind))
        m_a.append(message_a(k[p], lk.paillier_key_vec[lk.i - 1], r.r_k, lk.h1_h2_n_tilde_vec, r.alice))
    # ---- Round 1 (rounds.rs:122-206): MessageB for gamma_i and w_i to every other signer
    m_b_gamma = [[None] * (ttag - 1) for _ in range(ttag)]          # [bob p][j] -> for alice ind(p, j)
    m_b_w = [[None] * (ttag - 1) for _ in range(ttag)]
    beta_v = [[0] * (ttag - 1) for _ in range(ttag)]
    ni_v = [[0] * (ttag - 1) for _ in range(ttag)]
    for p in range(ttag):
        lk, r = keys[p], rnd[p]
        for j in range(ttag - 1):
            a = _ind(p, j)
            ek_a = lk.paillier_key_vec[l_s[a]]
            rb = message_b(gamma[p], ek_a, m_a[a], r.r_gamma[j], r.beta_tag_gamma[j], lk.h1_h2_n_tilde_vec, r.nonce_gamma_b[j], r.nonce_gamma_beta[j])
            rw = message_b(w[p], ek_a, m_a[a], r.r_w[j], r.beta_tag_w[j], lk.h1_h2_n_tilde_vec, r.nonce_w_b[j], r.nonce_w_beta[j])
            if rb is None or rw is None:
                res[p].status = ST_INVALID_KEY
                break
            m_b_gamma[p][j], beta_v[p][j] = rb
            m_b_w[p][j], ni_v[p][j] = rw
    if bad():
        return res
    # what alice p received from bob b: bob's list entry whose target is p
    recv = lambda table, p, b: table[b][p if p < b else p - 1]
    # ---- Round 2 (rounds.rs:234-317)
    delta, sigma, T, l_v, t_proof = [0] * ttag, [0] * ttag, [None] * ttag, [0] * ttag, [None] * ttag
    for p in range(ttag):
        lk, r = keys[p], rnd[p]
        g_w_vec = [pt_mul(lk.pk_vec[l_s[x]], lagrange_at_zero(l_s[x], l_s)) for x in range(ttag)]       # party_i.rs:527-544
        alpha_sum, miu_sum, ok = 0, 0, True
        for j in range(ttag - 1):
            b = _ind(p, j)
            ra = verify_proofs_get_alpha(recv(m_b_gamma, p, b), lk.dk, k[p])
            rm = verify_proofs_get_alpha(recv(m_b_w, p, b), lk.dk, k[p])
            if ra is None or rm is None or recv(m_b_w, p, b).b_proof.pk != g_w_vec[b]:                  # rounds.rs:281 (assert_eq!)
                ok = False
                break
            alpha_sum += ra[0]; miu_sum += rm[0]
        if not ok:
            res[p].status = ST_INVALID_KEY
            continue
        delta[p] = (k[p] * gamma[p] + alpha_sum + sum(beta_v[p])) % Q                                     # party_i.rs:591-604
        sigma[p] = (k[p] * w[p] + miu_sum + sum(ni_v[p])) % Q                                             # :606-618
        l_v[p] = r.l % Q
        T[p] = pt_add(pt_mul(G, sigma[p]), pt_mul(H2, l_v[p]))                                            # :620-634
        t_proof[p] = pedersen_prove(sigma[p], l_v[p], r.ped_s1 % Q, r.ped_s2 % Q)
    if bad():
        return res
    # ---- Round 3 (rounds.rs:347-402)
    dsum = sum(delta) % Q
    for p in range(ttag):
        if dsum == 0 or any(T[x] != t_proof[x].com for x in range(ttag)) or not all(pedersen_verify(t_proof[x]) for x in range(ttag)):
            res[p].status = ST_PROOF
    if bad():
        return res
    delta_inv = pow(dsum, -1, Q)                                                                          # party_i.rs:635-640
    # ---- Round 4 (rounds.rs:431-498)
    R, R_dash = [None] * ttag, [None] * ttag
    pdl = [[None] * (ttag - 1) for _ in range(ttag)]
    for p in range(ttag):
        lk, r = keys[p], rnd[p]
        ok = True
        for j in range(ttag - 1):                                                                         # party_i.rs:642-690
            b = _ind(p, j)
            ok = ok and recv(m_b_gamma, p, b).b_proof.pk == g_gamma[b] and \
                hash_commitment(bn_from_bytes(pt_compress(g_gamma[b])), rnd[b].blind) == com[b]
        if not ok:
            res[p].status = ST_COMMIT
            continue
        acc = None
        for x in range(ttag):
            acc = pt_add(acc, g_gamma[x])
        R[p] = pt_mul(acc, delta_inv)
        R_dash[p] = pt_mul(R[p], k[p])                                                                    # rounds.rs:452
        for j in range(ttag - 1):
            st = lk.h1_h2_n_tilde_vec[l_s[_ind(p, j)]]
            pdl[p][j] = pdl_prove(k[p], r.r_k, m_a[p].c, lk.paillier_key_vec[l_s[p]], R_dash[p], R[p], st.g, st.ni, st.N, *r.pdl[j])
    if bad():
        return res
    # ---- Round 5 (rounds.rs:525-592)
    S, heg = [None] * ttag, [None] * ttag
    for p in range(ttag):
        lk, r = keys[p], rnd[p]
        ok = True
        for x in range(ttag):                      # every signer's proof list, own included (party_i.rs:719-766)
            for j in range(ttag - 1):
                st = lk.h1_h2_n_tilde_vec[l_s[_ind(x, j)]]
                ok = ok and pdl_verify(pdl[x][j], m_a[x].c, lk.paillier_key_vec[l_s[x]], R_dash[x], R[p], st.g, st.ni, st.N)
        if not ok:
            res[p].status = ST_PDL
            continue
        ssum = G
        for x in range(ttag):
            ssum = pt_add(ssum, R_dash[x])                                                                # party_i.rs:768-776
        if pt_sub(ssum, G) != G:
            res[p].status = ST_PHASE5
            continue
        S[p] = pt_mul(R[p], sigma[p])                                                                     # :784
        heg[p] = heg_prove(l_v[p], sigma[p], R[p], H2, G, T[p], S[p], r.heg_s1 % Q, r.heg_s2 % Q)
    if bad():
        return res
    # ---- Round 6 (rounds.rs:612-636)
    for p in range(ttag):
        lk = keys[p]
        if not all(heg_verify(heg[x], R[p], H2, G, T[x], S[x]) for x in range(ttag)):                     # party_i.rs:801-833
            res[p].status = ST_PHASE6
            continue
        ssum = G
        for x in range(ttag):
            ssum = pt_add(ssum, S[x])
        if pt_sub(ssum, G) != lk.y_sum_s:                                                                 # :835-848
            res[p].status = ST_PHASE6
            continue
        res[p].R, res[p].sigma_i, res[p].k_i, res[p].t_vec = R[p], sigma[p], k[p], list(T)
    return res
