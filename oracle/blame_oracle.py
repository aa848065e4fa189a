"""CPU oracle for SURVEY.md section 8(f) rank 3: GG20 identifiable abort (/root/reference/src/protocols/multi_party_ecdsa/
gg_2020/blame.rs) — TEST INFRASTRUCTURE ONLY (same rule as gg20_oracle.py).  Each function returns the sorted list of bad
signer positions the reference puts into `ErrorType.bad_actors`."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

from oracle import gg20_oracle as o

Q = o.Q


@dataclass
class GlobalStatePhase5:
    """blame.rs:43-56; beta_randomness_vec[i][j] / beta_tag_vec[i][j] are already re-indexed per Alice i as
    `local_state_to_global_state` does (:60-116)"""
    k_vec: List[int]
    k_randomness_vec: List[int]
    gamma_vec: List[int]
    beta_randomness_vec: List[List[int]]
    beta_tag_vec: List[List[int]]
    encryption_key_vec: List[o.EncryptionKey]
    delta_vec: List[int]
    g_gamma_vec: List[o.Point]
    m_a_c: List[int]                     # MessageA.c of every signer
    m_b_c: List[List[int]]               # m_b_mat[i][j].c


def phase5_blame(st: GlobalStatePhase5) -> List[int]:
    """`GlobalStatePhase5::phase5_blame` (blame.rs:116-224)"""
    n = len(st.delta_vec)
    bad: List[int] = []
    for i in range(n):
        if st.g_gamma_vec[i] != o.pt_mul(o.G, st.gamma_vec[i]):                     # :121-125
            bad.append(i)
    ab = []
    for i in range(n):
        c = o.paillier_encrypt(st.encryption_key_vec[i], st.k_vec[i], st.k_randomness_vec[i])      # MessageA::a_with_predefined_randomness, no proofs
        if c != st.m_a_c[i]:
            bad.append(i)
        row = []
        if not bad:
            for j in range(n - 1):
                ind = j if j < i else j + 1
                ek = st.encryption_key_vec[i]
                beta_tag = st.beta_tag_vec[i][j]
                c_b = o.paillier_add(ek, o.paillier_mul(ek, c, st.gamma_vec[ind]), o.paillier_encrypt(ek, beta_tag, st.beta_randomness_vec[i][j]))
                beta = (-(beta_tag % Q)) % Q
                if c_b != st.m_b_c[i][j]:
                    bad.append(ind)
                row.append(((st.k_vec[i] * st.gamma_vec[ind] - beta) % Q, beta))
        ab.append(row)
    if not bad:
        for i in range(n):
            alpha_sum = sum(x[0] for x in ab[i]) % Q
            beta_sum = 0
            for j in range(n - 1):
                ind1 = j if j < i else j + 1
                ind2 = i - 1 if j < i else i
                beta_sum += ab[ind1][ind2][1]
            if st.delta_vec[i] != (st.k_vec[i] * st.gamma_vec[i] + alpha_sum + beta_sum) % Q:
                bad.append(i)
    return sorted(set(bad))


@dataclass
class GlobalStatePhase6:
    """blame.rs:236-247"""
    k_vec: List[int]
    k_randomness_vec: List[int]
    miu_vec: List[List[int]]             # plaintexts BEFORE reduction mod q
    miu_randomness_vec: List[List[int]]
    g_w_vec: List[o.Point]
    encryption_key_vec: List[o.EncryptionKey]
    proof_vec: List[o.ECDDHProof]
    S_vec: List[o.Point]
    m_a_c: List[int]
    m_b_c: List[List[int]]


def phase6_blame(st: GlobalStatePhase6, R: o.Point) -> List[int]:
    """`GlobalStatePhase6::phase6_blame` (blame.rs:322-431)"""
    n = len(st.k_vec)
    bad: List[int] = []
    for i in range(n):
        for j in range(n - 1):
            if o.paillier_encrypt(st.encryption_key_vec[i], st.miu_vec[i][j], st.miu_randomness_vec[i][j]) != st.m_b_c[i][j]:
                bad.append(i)
    for i in range(n):
        if o.paillier_encrypt(st.encryption_key_vec[i], st.k_vec[i], st.k_randomness_vec[i]) != st.m_a_c[i]:
            bad.append(i)
    if not bad:
        g_ni = [[o.pt_sub(o.pt_mul(st.g_w_vec[j if j < i else j + 1], st.k_vec[i]), o.pt_mul(o.G, st.miu_vec[i][j] % Q)) for j in range(n - 1)] for i in range(n)]
        g_sigma = []
        for i in range(n):
            acc = o.pt_mul(st.g_w_vec[i], st.k_vec[i])
            for x in st.miu_vec[i]:
                acc = o.pt_add(acc, o.pt_mul(o.G, x % Q))
            g_sigma.append(acc)
        for i in range(n):
            for j in range(n - 1):
                ind1 = j if j < i else j + 1
                ind2 = i - 1 if j < i else i
                g_sigma[i] = o.pt_add(g_sigma[i], g_ni[ind1][ind2])
        for i in range(n):
            if not o.ecddh_verify(st.proof_vec[i], o.G, g_sigma[i], R, st.S_vec[i]):
                bad.append(i)
    return sorted(set(bad))


def phase7_blame(s_vec: Sequence[int], r: int, R_dash_vec: Sequence[o.Point], m: int, R: o.Point, S_vec: Sequence[o.Point]) -> List[int]:
    """`GlobalStatePhase7::phase7_blame` (blame.rs:434-454): R s_i == R_dash_i m + S_i r"""
    return [i for i in range(len(s_vec)) if o.pt_mul(R, s_vec[i]) != o.pt_add(o.pt_mul(R_dash_vec[i], m % Q), o.pt_mul(S_vec[i], r))]
