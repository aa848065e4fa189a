"""CPU restatement of the GG18 signing phases that differ from GG20: phase 4 and phases 5a-5d of
/root/reference/src/protocols/multi_party_ecdsa/gg_2018/party_i.rs.  Phases 1-3 reuse the MtA of oracle/gg20_oracle.py with an
empty statement list (`MessageA::a(&k_i, &ek, &[])`, gg_2018/test.rs).

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else.  Parity status: UNPINNED (see oracle/gg20_oracle.py): the reference
has no vectors for these functions; the anchors are algebraic (phase 5d accepts exactly when the signature is valid) and
the final signature verifying under an independent ECDSA implementation.  All randomness is explicit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from oracle.gg20_oracle import (G, Q, DLogProof, HomoElGamalProof, Point, dlog_prove, dlog_verify, hash_commitment, heg_prove,
                                heg_verify, pt_add, pt_compress, pt_mul, pt_neg, pt_sub)
from oracle.lindell17_oracle import sha256_points_bigint

# failure codes, same numbering as include/tecdsa_b200.h
OK, INVALID_KEY, INVALID_SIG, INVALID_COM = 0, 2, 9, 11


def phase1_broadcast(g_gamma_i: Point, blind: int) -> int:
    """`SignKeys::phase1_broadcast` party_i.rs:409-425: commitment to the compressed g^gamma_i."""
    return hash_commitment(int.from_bytes(pt_compress(g_gamma_i), "big"), blind)


def phase2_sum(k_i: int, other: int, a_vec: Sequence[int], b_vec: Sequence[int]) -> int:
    """`phase2_delta_i` / `phase2_sigma_i` party_i.rs:427-445: k_i * other + sum(alpha) + sum(beta)."""
    assert len(a_vec) == len(b_vec)
    return (k_i * other + sum(a_vec) + sum(b_vec)) % Q


def phase3_reconstruct_delta(delta_vec: Sequence[int]) -> Optional[int]:
    """party_i.rs:447-453; None where the reference panics ("sum of deltas is zero")."""
    d = sum(delta_vec) % Q
    return pow(d, -1, Q) if d else None


def phase4(delta_inv: int, b_proof_pks: Sequence[Point], decommits: Sequence[Tuple[int, Point]], coms: Sequence[int]) -> Optional[Point]:
    """`SignKeys::phase4` party_i.rs:455-485: every DLogProof public key equals the decommitted g^gamma and every commitment
    reopens, then R = delta^-1 * sum(g^gamma).  decommits = (blind_factor, g_gamma_i).  None = Err(InvalidKey)."""
    for pk, (blind, gg), com in zip(b_proof_pks, decommits, coms):
        if pk != gg or phase1_broadcast(gg, blind) != com:
            return None
    acc = None
    for _, gg in decommits:
        acc = pt_add(acc, gg)
    return pt_mul(acc, delta_inv)


def phase5_local_sig(k_i: int, message: int, R: Point, sigma_i: int) -> int:
    """`LocalSignature::phase5_local_sig` party_i.rs:489-511: s_i = m k_i + r sigma_i."""
    return ((message % Q) * k_i + (R[0] % Q) * sigma_i) % Q


@dataclass
class Phase5A:
    com: int
    V: Point
    A: Point
    B: Point
    blind: int
    heg: HomoElGamalProof
    dlog: DLogProof


def phase5a(s_i: int, l_i: int, rho_i: int, R: Point, blind: int, heg_s1: int, heg_s2: int, dlog_nonce: int) -> Phase5A:
    """`phase5a_broadcast_5b_zkproof` party_i.rs:513-558: A = rho G, B = (l rho) G, V = s R + l G; the commitment hashes
    (V, A, B) as uncompressed points; HomoELGamalProof for (G=A, H=R, Y=g, D=V, E=B) with witness (x = s_i, r = l_i);
    DLogProof of rho_i."""
    A = pt_mul(G, rho_i)
    B = pt_mul(G, l_i * rho_i % Q)
    V = pt_add(pt_mul(R, s_i), pt_mul(G, l_i))
    com = hash_commitment(sha256_points_bigint([V, A, B]), blind)
    heg = heg_prove(s_i, l_i, A, R, G, V, B, heg_s1, heg_s2)
    return Phase5A(com, V, A, B, blind, heg, dlog_prove(rho_i, dlog_nonce))


def phase5c(message: int, R: Point, y: Point, rho_i: int, l_i: int, others: Sequence[Phase5A], v_i: Point, blind2: int
            ) -> Tuple[int, Optional[Tuple[int, Point, Point]]]:
    """`phase5c` party_i.rs:560-629.  `others` are the other signers' phase-5a messages (commitment, decommitment, ElGamal
    proof, DLog proof of rho).  Returns (code, (com2, u_i, t_i)): InvalidCom unless every commitment reopens and both proofs
    verify.  An identity point among the hashed values has no recallable encoding: reported as INVALID_SIG."""
    ok = True
    for o in others:
        if hash_commitment(sha256_points_bigint([o.V, o.A, o.B]), o.blind) != o.com:
            ok = False
        elif not heg_verify(o.heg, o.A, R, G, o.V, o.B):
            ok = False
        elif not dlog_verify(o.dlog):
            ok = False
    v = v_i
    a = None
    for o in others:
        v = pt_add(v, o.V)
        a = pt_add(a, o.A)
    r = R[0] % Q
    v = pt_sub(pt_sub(v, pt_mul(G, message % Q)), pt_mul(y, r))
    u_i = pt_mul(v, rho_i) if v is not None else None
    t_i = pt_mul(a, l_i) if a is not None else None
    if u_i is None or t_i is None:
        return INVALID_SIG, None
    com2 = hash_commitment(sha256_points_bigint([u_i, t_i]), blind2)
    if not ok:
        return INVALID_COM, None
    return OK, (com2, u_i, t_i)


def phase5d(decom2: Sequence[Tuple[Point, Point, int]], com2: Sequence[int], B_all: Sequence[Point]) -> int:
    """`phase5d` party_i.rs:631-665 over ALL signers' (u_i, t_i, blind), second commitments and B_i: InvalidCom if a commitment
    does not reopen, InvalidKey unless g + sum(t) + sum(B) - sum(u) == g."""
    test_com = all(hash_commitment(sha256_points_bigint([u, t]), bl) == c for (u, t, bl), c in zip(decom2, com2))
    acc = G
    for (_, t, _) in decom2:
        acc = pt_add(acc, t)
    for b in B_all:
        acc = pt_add(acc, b)
    for (u, _, _) in decom2:
        acc = pt_sub(acc, u)
    if not test_com:
        return INVALID_COM
    return OK if acc == G else INVALID_KEY


def output_signature(R: Point, y: Point, message: int, s_all: Sequence[int]) -> Tuple[int, Optional[Tuple[int, int, int]]]:
    """`output_signature` party_i.rs:666-703 + `verify` :706-730: s = sum(s_i), low-s normalisation with recid, then the
    in-tree verification (r == x(u1 G + u2 y) mod q)."""
    s = sum(s_all) % Q
    r = R[0] % Q
    recid = (R[1] % Q) & 1
    if s > Q - s:
        s = Q - s
        recid ^= 1
    if s == 0:
        return INVALID_SIG, None
    b = pow(s, -1, Q)
    pt = pt_add(pt_mul(G, (message % Q) * b % Q), pt_mul(y, r * b % Q))
    if pt is None or r != pt[0] % Q:
        return INVALID_SIG, None
    return OK, (r, s, recid)
