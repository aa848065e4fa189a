"""CPU restatement of the Lindell-2017 two-party ECDSA path and of the interactive PDL proof.

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the product path never routes through this file.
Parity status: UNPINNED, like oracle/gg20_oracle.py — the reference ships no vectors for this path and cannot be built in
this image; every function cites the reference lines it restates, the out-of-tree behaviour ([R], curv-kzen 0.9 /
kzen-paillier 0.4.2) is the one documented at the top of gg20_oracle.py.  All randomness is explicit.

Follows (read-only):
  /root/reference/src/protocols/two_party_ecdsa/lindell_2017/party_one.rs
  /root/reference/src/protocols/two_party_ecdsa/lindell_2017/party_two.rs
  /root/reference/src/utilities/zk_pdl/mod.rs
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Optional, Tuple

from oracle.gg20_oracle import (G, H2 as BASE_POINT2, Q, DLogProof, DecryptionKey, ECDDHProof, EncryptionKey, Point, bn_bytes,
                                dlog_prove, dlog_verify, ecddh_prove, ecddh_verify, hash_commitment, paillier_add,
                                paillier_decrypt, paillier_encrypt, paillier_mul, pt_add, pt_compress, pt_mul, pt_uncompressed)


def _commit_point(p: Point, blind: int) -> int:
    """`HashCommitment::create_commitment_with_user_defined_randomness(&BigInt::from_bytes(P.to_bytes(true)), &blind)`"""
    return hash_commitment(int.from_bytes(pt_compress(p), "big"), blind)


def sha256_points_bigint(points) -> int:
    """`Sha256::new().chain_points([...]).result_bigint()`: 65-byte uncompressed points [R], digest as a BigInt (not reduced)."""
    h = hashlib.sha256()
    for p in points:
        h.update(pt_uncompressed(p))
    return int.from_bytes(h.digest(), "big")


# --------------------------------------------------------------------------- key generation messages
@dataclass
class KeyGenCommit:
    pk_commitment: int
    zk_pok_commitment: int
    public_share: Point
    d_log_proof: DLogProof


def p1_keygen_first(secret_share: int, dlog_nonce: int, pk_blind: int, zk_pok_blind: int) -> KeyGenCommit:
    """`party_one::KeyGenFirstMsg::create_commitments_with_fixed_secret_share` party_one.rs:179-218."""
    pf = dlog_prove(secret_share, dlog_nonce)
    return KeyGenCommit(_commit_point(pf.pk, pk_blind), _commit_point(pf.pk_t_rand_commitment, zk_pok_blind), pf.pk, pf)


def p2_keygen_verify(first: KeyGenCommit, pk_blind: int, zk_pok_blind: int) -> bool:
    """`party_two::KeyGenSecondMsg::verify_commitments_and_dlog_proof` party_two.rs:180-223: both commitments reopen and the
    DLogProof verifies."""
    if first.pk_commitment != _commit_point(first.public_share, pk_blind):
        return False
    if first.zk_pok_commitment != _commit_point(first.d_log_proof.pk_t_rand_commitment, zk_pok_blind):
        return False
    return dlog_verify(first.d_log_proof)


# --------------------------------------------------------------------------- ephemeral key exchange
@dataclass
class EphFirst:
    public_share: Point
    c: Point
    proof: ECDDHProof
    pk_commitment: Optional[int] = None        # party two only
    zk_pok_commitment: Optional[int] = None


def eph_create(secret_share: int, nonce: int, pk_blind: Optional[int] = None, zk_pok_blind: Optional[int] = None) -> EphFirst:
    """`party_one::EphKeyGenFirstMsg::create` party_one.rs:403-433 (no commitments) and
    `party_two::EphKeyGenFirstMsg::create_commitments` party_two.rs:314-371 (with both hash commitments):
    public_share = k G, c = k H (H = base_point2), ECDDHProof over (G, kG, H, kH);
    pk_commitment commits to the compressed public share, zk_pok_commitment to H(a1, a2) of the proof."""
    pub = pt_mul(G, secret_share)
    c = pt_mul(BASE_POINT2, secret_share)
    pf = ecddh_prove(secret_share, G, pub, BASE_POINT2, c, nonce)
    out = EphFirst(pub, c, pf)
    if pk_blind is not None:
        out.pk_commitment = _commit_point(pub, pk_blind)
        out.zk_pok_commitment = hash_commitment(sha256_points_bigint([pf.a1, pf.a2]), zk_pok_blind)
    return out


def p1_eph_verify(msg: EphFirst, pk_blind: int, zk_pok_blind: int) -> bool:
    """`party_one::EphKeyGenSecondMsg::verify_commitments_and_dlog_proof` party_one.rs:436-482."""
    if msg.pk_commitment != _commit_point(msg.public_share, pk_blind):
        return False
    if msg.zk_pok_commitment != hash_commitment(sha256_points_bigint([msg.proof.a1, msg.proof.a2]), zk_pok_blind):
        return False
    return ecddh_verify(msg.proof, G, msg.public_share, BASE_POINT2, msg.c)


def p2_eph_verify(msg: EphFirst) -> bool:
    """`party_two::EphKeyGenSecondMsg::verify_and_decommit` party_two.rs:374-387."""
    return ecddh_verify(msg.proof, G, msg.public_share, BASE_POINT2, msg.c)


# --------------------------------------------------------------------------- signing
def p2_partial_sig(ek: EncryptionKey, encrypted_secret_share: int, x2: int, k2: int, eph_other_public: Point, message: int,
                   rho: int, enc_randomness: int) -> Optional[int]:
    """`party_two::PartialSig::compute` party_two.rs:390-424.  rho < q^2 and the Paillier randomness of `Paillier::encrypt`
    are explicit.  None where the reference panics (k2 = 0: `mod_inv(..).unwrap()`)."""
    if k2 % Q == 0:
        return None
    r = pt_mul(eph_other_public, k2)
    rx = r[0] % Q
    k2_inv = pow(k2, -1, Q)
    partial_sig = rho * Q + (k2_inv * message) % Q
    c1 = paillier_encrypt(ek, partial_sig, enc_randomness)
    v = k2_inv * (rx * x2 % Q) % Q
    c2 = paillier_mul(ek, encrypted_secret_share, v)
    return paillier_add(ek, c2, c1)


def p1_sign(dk: DecryptionKey, c3: int, k1: int, eph_other_public: Point) -> Optional[Tuple[int, int, int]]:
    """`party_one::Signature::compute_with_recid` party_one.rs:519-564 -> (r, s, recid); `compute` (:486-517) returns the same
    (r, s).  None where the reference panics (k1 = 0)."""
    if k1 % Q == 0:
        return None
    r = pt_mul(eph_other_public, k1)
    rx, ry = r[0] % Q, r[1] % Q
    k1_inv = pow(k1, -1, Q)
    s_tag = paillier_decrypt(dk, c3)
    s2 = (s_tag % Q) * k1_inv % Q
    s = min(s2, Q - s2)
    recid = ry & 1
    if s2 > Q - s2:
        recid ^= 1
    return rx, s, recid


def verify(r: int, s: int, pubkey: Point, message: int) -> bool:
    """`party_one::verify` party_one.rs:567-592: x(u1 + u2) compared with r as BYTE STRINGS (so an x >= q never matches) and
    the low-s rule s < q - s.  False where the reference panics (s = 0) or errors."""
    if s % Q == 0:
        return False
    s_inv = pow(s % Q, -1, Q)
    u1 = pt_mul(G, (message % Q) * s_inv % Q)
    u2 = pt_mul(pubkey, (r % Q) * s_inv % Q)
    pt = pt_add(u1, u2)
    if pt is None:
        return False
    return bn_bytes(r) == bn_bytes(pt[0]) and s < Q - s


# --------------------------------------------------------------------------- interactive PDL proof (utilities/zk_pdl/mod.rs)
@dataclass
class PdlVerifierState:
    c_tag: int
    c_tag_tag: int
    a: int
    b: int
    blindness: int
    q_tag: Point


def pdl_verifier_message1(ek: EncryptionKey, ciphertext: int, Qpt: Point, a: int, b: int, enc_randomness: int,
                          blindness: int) -> PdlVerifierState:
    """`Verifier::message1` zk_pdl/mod.rs:111-148: c' = a (x) c (+) Enc(b), c'' = commit(a + (b << bit_length(a)); blindness),
    Q' = a Q + b G (b reduced mod q as a scalar).  a < q, b < q^2; Enc's randomness is explicit."""
    b_enc = paillier_encrypt(ek, b, enc_randomness)
    ac = paillier_mul(ek, ciphertext, a)
    c_tag = paillier_add(ek, ac, b_enc)
    ab_concat = a + (b << a.bit_length())
    c_tag_tag = hash_commitment(ab_concat, blindness)
    q_tag = pt_add(pt_mul(Qpt, a), pt_mul(G, b % Q))
    return PdlVerifierState(c_tag, c_tag_tag, a, b, blindness, q_tag)


def pdl_prover_message1(dk: DecryptionKey, c_tag: int, blindness: int) -> Tuple[int, Point, int]:
    """`Prover::message1` zk_pdl/mod.rs:191-215 without the out-of-tree `RangeProofNi` -> (c_hat, q_hat, alpha)."""
    alpha = paillier_decrypt(dk, c_tag)
    q_hat = pt_mul(G, alpha % Q)
    c_hat = hash_commitment(int.from_bytes(pt_compress(q_hat), "big"), blindness)
    return c_hat, q_hat, alpha


def pdl_prover_message2(x1: int, alpha: int, c_tag_tag: int, a: int, b: int, blindness: int) -> bool:
    """`Prover::message2` zk_pdl/mod.rs:217-243: the verifier's decommitment reopens and a x1 + b == alpha over the integers."""
    ab_concat = a + (b << a.bit_length())
    return a * x1 + b == alpha and c_tag_tag == hash_commitment(ab_concat, blindness)


def pdl_verifier_finalize(c_hat: int, q_hat: Point, blindness: int, q_tag: Point) -> bool:
    """`Verifier::finalize` zk_pdl/mod.rs:170-187."""
    return c_hat == hash_commitment(int.from_bytes(pt_compress(q_hat), "big"), blindness) and q_hat == q_tag
