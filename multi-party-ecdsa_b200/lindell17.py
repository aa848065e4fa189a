"""Host-side mirror of the Lindell-2017 two-party ECDSA interface and of the interactive PDL proof (SURVEY.md section 8(f)
rank 4) over the batched C ABI: /root/reference/src/protocols/two_party_ecdsa/lindell_2017/{party_one,party_two}.rs and
/root/reference/src/utilities/zk_pdl/mod.rs.  Function names follow the reference's (`party_one::KeyGenFirstMsg::
create_commitments` -> `p1_keygen_first`, `party_two::PartialSig::compute` -> `p2_partial_sig`, ...); every argument is a
sequence with one entry per batch element and all randomness is explicit.  No arithmetic happens here — only packing."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import HOST, Engine, ints_to_limbs, limbs_to_ints, _ptr
from . import gg20, keygen
from .gg20 import KeySets, _Screen, _pts, unpack_point

Point = Tuple[int, int]


def _bind(lib):
    if getattr(lib, "_l17_bound", False):
        return
    V, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.tecdsa_l17_eph_create_batch.argtypes = [V] * 10 + [S, I]
    lib.tecdsa_l17_eph_verify_batch.argtypes = [V] * 9 + [S, I]
    lib.tecdsa_l17_partial_sig_batch.argtypes = [V, V, V, S] + [V] * 9 + [S, I]
    lib.tecdsa_l17_sign_batch.argtypes = [V] * 10 + [S, I]
    lib.tecdsa_l17_verify_batch.argtypes = [V] * 6 + [S, I]
    lib.tecdsa_zkpdl_verifier_message1_batch.argtypes = [V, V, V, S] + [V] * 10 + [S, I]
    lib.tecdsa_zkpdl_prover_message1_batch.argtypes = [V] * 9 + [S, I]
    lib.tecdsa_zkpdl_prover_message2_batch.argtypes = [V] * 8 + [S, I]
    lib.tecdsa_zkpdl_verifier_finalize_batch.argtypes = [V] * 6 + [S, I]
    lib._l17_bound = True


def _points(a: np.ndarray):
    return [unpack_point(v) for v in limbs_to_ints(a)]


def _screen_pts(sc: _Screen, pts):
    """Untrusted points: a coordinate that does not fit 256 bits (or a missing point) can never deserialise in the reference; it is
    replaced by the identity encoding and that ONE element is marked rejected after the call (never an OverflowError for the batch)"""
    out = []
    for i, p in enumerate(pts):
        if p is None or p[0] < 0 or p[1] < 0 or p[0] >> 256 or p[1] >> 256:
            sc.bad[i] = True
            p = None
        out.append(p)
    return _pts(out)


# ----------------------------------------------------------------------------- key generation (composition of existing calls)
def p1_keygen_first(eng: Engine, secret_share, dlog_nonce, pk_blind, zk_pok_blind):
    """`party_one::KeyGenFirstMsg::create_commitments_with_fixed_secret_share` (party_one.rs:179-218) ->
    (pk_commitment, zk_pok_commitment, public_share, d_log_proof[n][40])"""
    proof = gg20.dlog_prove(eng, secret_share, dlog_nonce)
    pk = _points(proof[:, :16])
    t = _points(proof[:, 16:32])
    return gg20.hash_commitment(eng, pk, pk_blind), gg20.hash_commitment(eng, t, zk_pok_blind), pk, proof


def p2_keygen_verify(eng: Engine, pk_commitment, zk_pok_commitment, public_share, d_log_proof: np.ndarray, pk_blind, zk_pok_blind) -> np.ndarray:
    """`party_two::KeyGenSecondMsg::verify_commitments_and_dlog_proof` (party_two.rs:180-223) -> status per element
    (0 accept, 11 a commitment does not reopen, 10 the DLogProof fails)"""
    t = _points(d_log_proof[:, 16:32])
    c1 = gg20.hash_commitment(eng, public_share, pk_blind)
    c2 = gg20.hash_commitment(eng, t, zk_pok_blind)
    st = gg20.dlog_verify(eng, d_log_proof).copy()
    for i in range(len(st)):
        if c1[i] != pk_commitment[i] or c2[i] != zk_pok_commitment[i]:
            st[i] = 11
    return st


def generate_h1_h2_n_tilde(eng: Engine, setups):
    """Lindell's `generate_h1_h2_n_tilde` (party_one.rs:594-607) with the samples explicit: setups[i] = (p~, q~, h1, xhi) ->
    [(N~, h1, h2, xhi)] with h2 = (h1^-1)^xhi mod N~  (GG20's variant, gg_2020/party_i.rs:137-156, is keygen.h1_h2_n_tilde)"""
    params, _ = keygen.h1_h2_n_tilde(eng, setups)                 # N~ = p~ q~ on the device
    nt = [prm[0] for prm in params]
    h1 = [s_[2] for s_ in setups]
    h1_inv = eng.mod_inv(h1, nt)
    h2, _ = eng.mod_pow([v or 0 for v in h1_inv], [s_[3] for s_ in setups], nt)
    return [(nt[i], h1[i], h2[i], setups[i][3]) for i in range(len(setups))]


def p1_paillier_and_proofs(eng: Engine, keys: KeySets, key_row, st_row, statements, xhi, x1, randomness, pdl_rand, cdlog_nonce, p_q):
    """`PaillierKeyPair::generate_encrypted_share_from_fixed_paillier_keypair`, `generate_ni_proof_correct_key` and `pdl_proof`
    (party_one.rs:339-400): Paillier key rows `key_row` and (N~, h1, h2) rows `st_row` of an uploaded key set, statements[i] =
    (N~, h1, h2) of row st_row[i] with its secret xhi[i] -> dict(encrypted_share, correct_key_proof (11 sigmas), pdl (PDLwSlackProof
    fields), composite_dlog_proof (x, y), Q).  pdl_rand = (alpha, beta, rho, gamma) sequences; p_q the primes of the Paillier keys."""
    n_list = [p * q for p, q in p_q]
    c_key = eng.paillier_encrypt(n_list, list(range(len(n_list))), list(x1), list(randomness))
    sigma, _ = keygen.correct_key_prove(eng, p_q)
    cd = keygen.composite_dlog_prove(eng, list(statements), list(xhi), list(cdlog_nonce))
    Q = eng.secp_mul(None, list(x1))
    G = eng.secp_mul(None, [1] * len(x1))
    alpha, beta, rho, gamma = pdl_rand
    pdl = gg20.pdl_prove(eng, keys, key_row, st_row, list(x1), list(randomness), c_key, Q, G, alpha, beta, rho, gamma)
    return {"encrypted_share": c_key, "correct_key_proof": sigma, "pdl": pdl, "composite_dlog_proof": cd, "Q": Q, "G": G}


def p2_verify_paillier_and_proofs(eng: Engine, keys: KeySets, key_row, st_row, statements, n_list, msg, q1) -> np.ndarray:
    """`PaillierPublic::verify_ni_proof_correct_key` (party_two.rs:302-311, incl. the bit-length floor of the modulus) and
    `PaillierPublic::pdl_verify` (party_two.rs:275-300) over what p1_paillier_and_proofs produced -> status per element
    (10 = IncorrectProof of the key, 6 = PartyTwoError::PdlVerify)"""
    st = keygen.correct_key_verify(eng, n_list, msg["correct_key_proof"]).copy()
    for i, n in enumerate(n_list):
        if n.bit_length() < 2047 and not st[i]:                  # `ek.n.bit_length() < PAILLIER_KEY_SIZE - 1`
            st[i] = 10
    cd = keygen.composite_dlog_verify(eng, list(statements), msg["composite_dlog_proof"])
    pd = msg["pdl"]
    pv = gg20.pdl_verify(eng, keys, key_row, st_row, msg["encrypted_share"], msg["Q"], msg["G"], pd["z"], pd["u1"], pd["u2"], pd["u3"], pd["s1"], pd["s2"], pd["s3"])
    for i in range(len(st)):
        if not st[i] and (cd[i] or pv[i] or msg["Q"][i] != q1[i]):
            st[i] = 6
    return st


# ----------------------------------------------------------------------------- ephemeral keys
def eph_create(eng: Engine, secret_share, nonce, pk_blind=None, zk_pok_blind=None):
    """`party_one::EphKeyGenFirstMsg::create` (no blinds) / `party_two::EphKeyGenFirstMsg::create_commitments` (with blinds)
    -> dict(public_share, c, proof[n][40], pk_commitment, zk_pok_commitment)"""
    _bind(eng.lib)
    n = len(secret_share)
    k, s = ints_to_limbs(secret_share, 8), ints_to_limbs(nonce, 8)
    b1 = ints_to_limbs(pk_blind, 8) if pk_blind is not None else None
    b2 = ints_to_limbs(zk_pok_blind, 8) if zk_pok_blind is not None else None
    pub, c, pf = np.zeros((n, 16), np.uint32), np.zeros((n, 16), np.uint32), np.zeros((n, 40), np.uint32)
    c1 = np.zeros((n, 8), np.uint32) if b1 is not None else None
    c2 = np.zeros((n, 8), np.uint32) if b1 is not None else None
    eng._ck(eng.lib.tecdsa_l17_eph_create_batch(eng._ctx, _ptr(k), _ptr(s), _ptr(b1), _ptr(b2), _ptr(pub), _ptr(c), _ptr(pf), _ptr(c1), _ptr(c2), n, HOST),
            "l17_eph_create")
    return {"public_share": _points(pub), "c": _points(c), "proof": pf,
            "pk_commitment": limbs_to_ints(c1) if c1 is not None else None, "zk_pok_commitment": limbs_to_ints(c2) if c2 is not None else None}


def eph_verify(eng: Engine, public_share, c, proof: np.ndarray, pk_blind=None, zk_pok_blind=None, pk_commitment=None, zk_pok_commitment=None) -> np.ndarray:
    """`party_one::EphKeyGenSecondMsg::verify_commitments_and_dlog_proof` (with the four commitment arguments) /
    `party_two::EphKeyGenSecondMsg::verify_and_decommit` (without) -> status"""
    _bind(eng.lib)
    n = len(public_share)
    sc = _Screen(n)
    opt = [sc.limbs(v, 8) if v is not None else None for v in (pk_blind, zk_pok_blind, pk_commitment, zk_pok_commitment)]
    ins = [_screen_pts(sc, public_share), _screen_pts(sc, c), np.ascontiguousarray(proof, dtype=np.uint32)]
    st = np.full(n, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_l17_eph_verify_batch(eng._ctx, *[_ptr(a) for a in ins], *[_ptr(a) for a in opt], _ptr(st), n, HOST), "l17_eph_verify")
    return sc.apply(st, 10)                  # TECDSA_ST_PROOF: the message would not deserialise


# ----------------------------------------------------------------------------- signing
def p2_partial_sig(eng: Engine, n_list, key_idx, c_key, x2, k2, eph_other_public, message, rho, randomness):
    """`party_two::PartialSig::compute` (party_two.rs:390-424) -> (c3 list, status).  message = the hashed message as an integer below
    2^256 (reduced mod q on the device); reduce anything wider with Engine.scalar_from_bigint first."""
    _bind(eng.lib)
    cnt = len(c_key)
    N, idx = ints_to_limbs(n_list, 64), np.asarray(key_idx, dtype=np.uint32)
    sc = _Screen(cnt)
    ins = [sc.limbs(c_key, 128), ints_to_limbs(x2, 8), ints_to_limbs(k2, 8), _screen_pts(sc, eph_other_public), ints_to_limbs(message, 8),
           ints_to_limbs(rho, 16), ints_to_limbs(randomness, 64)]
    c3, st = np.zeros((cnt, 128), np.uint32), np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_l17_partial_sig_batch(eng._ctx, _ptr(N), _ptr(idx), len(n_list), *[_ptr(a) for a in ins], _ptr(c3), _ptr(st), cnt, HOST),
            "l17_partial_sig")
    return limbs_to_ints(c3), sc.apply(st, 2)                # TECDSA_ST_INVALID_KEY: party one's material is malformed


def p1_sign(eng: Engine, keys: KeySets, key_row, c3, k1, eph_other_public):
    """`party_one::Signature::compute_with_recid` (party_one.rs:519-564) -> (r, s, recid, status)"""
    _bind(eng.lib)
    cnt = len(c3)
    rows = np.asarray(key_row, dtype=np.uint32)
    sc = _Screen(cnt)
    ins = [sc.limbs(c3, 128), ints_to_limbs(k1, 8), _screen_pts(sc, eph_other_public)]
    r, s = np.zeros((cnt, 8), np.uint32), np.zeros((cnt, 8), np.uint32)
    rec, st = np.zeros(cnt, np.uint8), np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_l17_sign_batch(eng._ctx, keys.handle, _ptr(rows), *[_ptr(a) for a in ins], _ptr(r), _ptr(s), _ptr(rec), _ptr(st), cnt, HOST), "l17_sign")
    return limbs_to_ints(r), limbs_to_ints(s), rec, sc.apply(st, 2)


def verify(eng: Engine, r, s, pubkey, message) -> np.ndarray:
    """`party_one::verify` (party_one.rs:567-592) -> status (0 accept, 9 InvalidSig)"""
    _bind(eng.lib)
    cnt = len(r)
    sc = _Screen(cnt)
    ins = [sc.limbs(r, 8), sc.limbs(s, 8), _screen_pts(sc, pubkey), ints_to_limbs(message, 8)]
    st = np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_l17_verify_batch(eng._ctx, *[_ptr(a) for a in ins], _ptr(st), cnt, HOST), "l17_verify")
    return sc.apply(st, 9)                   # TECDSA_ST_INVALID_SIG: an r or s wider than 256 bits matches no x coordinate / fails s < q - s


# ----------------------------------------------------------------------------- interactive PDL proof
def pdl_verifier_message1(eng: Engine, n_list, key_idx, ciphertext, Q, a, b, randomness, blindness):
    """`zk_pdl::Verifier::message1` (zk_pdl/mod.rs:111-148) -> (c_tag, c_tag_tag, q_tag, status)"""
    _bind(eng.lib)
    cnt = len(ciphertext)
    N, idx = ints_to_limbs(n_list, 64), np.asarray(key_idx, dtype=np.uint32)
    ins = [ints_to_limbs(ciphertext, 128), _pts(Q), ints_to_limbs(a, 8), ints_to_limbs(b, 16), ints_to_limbs(randomness, 64), ints_to_limbs(blindness, 8)]
    ct, ctt, qt = np.zeros((cnt, 128), np.uint32), np.zeros((cnt, 8), np.uint32), np.zeros((cnt, 16), np.uint32)
    st = np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_zkpdl_verifier_message1_batch(eng._ctx, _ptr(N), _ptr(idx), len(n_list), *[_ptr(x) for x in ins], _ptr(ct), _ptr(ctt), _ptr(qt),
                                                         _ptr(st), cnt, HOST), "zkpdl_verifier_message1")
    return limbs_to_ints(ct), limbs_to_ints(ctt), _points(qt), st


def pdl_prover_message1(eng: Engine, keys: KeySets, key_row, c_tag, blindness):
    """`zk_pdl::Prover::message1` (zk_pdl/mod.rs:191-215) without the out-of-tree RangeProofNi -> (c_hat, q_hat, alpha, status)"""
    _bind(eng.lib)
    cnt = len(c_tag)
    rows = np.asarray(key_row, dtype=np.uint32)
    sc = _Screen(cnt)
    ins = [sc.limbs(c_tag, 128), ints_to_limbs(blindness, 8)]
    ch, qh, al = np.zeros((cnt, 8), np.uint32), np.zeros((cnt, 16), np.uint32), np.zeros((cnt, 64), np.uint32)
    st = np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_zkpdl_prover_message1_batch(eng._ctx, keys.handle, _ptr(rows), *[_ptr(x) for x in ins], _ptr(ch), _ptr(qh), _ptr(al), _ptr(st), cnt, HOST),
            "zkpdl_prover_message1")
    return limbs_to_ints(ch), _points(qh), limbs_to_ints(al), sc.apply(st, 2)


def pdl_prover_message2(eng: Engine, x1, alpha, c_tag_tag, a, b, blindness) -> np.ndarray:
    """`zk_pdl::Prover::message2` (zk_pdl/mod.rs:217-243) -> status (0 = decommit, 6 = ZkPdlError::Message2)"""
    _bind(eng.lib)
    cnt = len(x1)
    sc = _Screen(cnt)
    ins = [ints_to_limbs(x1, 8), ints_to_limbs(alpha, 64), sc.limbs(c_tag_tag, 8), sc.limbs(a, 8), sc.limbs(b, 16), sc.limbs(blindness, 8)]
    st = np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_zkpdl_prover_message2_batch(eng._ctx, *[_ptr(x) for x in ins], _ptr(st), cnt, HOST), "zkpdl_prover_message2")
    return sc.apply(st, 6)                   # an over-wide decommitment can match neither alpha nor the commitment


def pdl_verifier_finalize(eng: Engine, c_hat, q_hat, blindness, q_tag) -> np.ndarray:
    """`zk_pdl::Verifier::finalize` (zk_pdl/mod.rs:170-187) -> status (0 accept, 6 = ZkPdlError::Finalize)"""
    _bind(eng.lib)
    cnt = len(c_hat)
    sc = _Screen(cnt)
    ins = [sc.limbs(c_hat, 8), _screen_pts(sc, q_hat), sc.limbs(blindness, 8), _pts(q_tag)]
    st = np.full(cnt, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_zkpdl_verifier_finalize_batch(eng._ctx, *[_ptr(x) for x in ins], _ptr(st), cnt, HOST), "zkpdl_verifier_finalize")
    return sc.apply(st, 6)
