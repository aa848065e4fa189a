"""Host-side mirror of the GG18 signing phases that GG20 replaced over the batched C ABI:
`SignKeys::phase4`, `LocalSignature::{phase5_local_sig, phase5a_broadcast_5b_zkproof, phase5c, phase5d, output_signature}` of
/root/reference/src/protocols/multi_party_ecdsa/gg_2018/party_i.rs:455-730.  A batch is `sessions` signing sessions of `parties`
signers; element u = session * parties + party and every argument is a flat element-major sequence.  Phases 1-3 of GG18 are the
MtA of the hot path with an empty statement list (gg20.mta_message_a / mta_message_b / mta_get_alpha with n_st = 0) plus scalar
sums (Engine.scalar_op).  No arithmetic happens here — only packing."""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import HOST, Engine, ints_to_limbs, limbs_to_ints, _ptr
from .gg20 import _pts, unpack_point


def _bind(lib):
    if getattr(lib, "_gg18_bound", False):
        return
    V, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.tecdsa_gg18_phase4_batch.argtypes = [V, I] + [V] * 7 + [S, I]
    lib.tecdsa_gg18_local_sig_batch.argtypes = [V] * 6 + [S, I]
    lib.tecdsa_gg18_phase5a_batch.argtypes = [V] * 14 + [S, I]
    lib.tecdsa_gg18_phase5c_batch.argtypes = [V, I] + [V] * 14 + [S, I]
    lib.tecdsa_gg18_phase5d_batch.argtypes = [V, I] + [V] * 5 + [S, I]
    lib.tecdsa_gg18_output_signature_batch.argtypes = [V, I] + [V] * 8 + [S, I]
    lib._gg18_bound = True


def _points(a: np.ndarray):
    return [unpack_point(v) for v in limbs_to_ints(a)]


def phase4(eng: Engine, parties: int, delta_inv, b_proof_pk, g_gamma, blind, com):
    """`SignKeys::phase4` (party_i.rs:455-485).  b_proof_pk[u] = the `parties` DLogProof public keys element u holds (entry j from
    signer j; its own entry = its own g^gamma) -> (R per element, status)"""
    _bind(eng.lib)
    count = len(delta_inv)
    flat = [p for row in b_proof_pk for p in row]
    ins = [ints_to_limbs(delta_inv, 8), _pts(flat), _pts(g_gamma), ints_to_limbs(blind, 8), ints_to_limbs(com, 8)]
    R, st = np.zeros((count, 16), np.uint32), np.full(count, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_gg18_phase4_batch(eng._ctx, parties, *[_ptr(a) for a in ins], _ptr(R), _ptr(st), count // parties, HOST), "gg18_phase4")
    return _points(R), st


def local_sig(eng: Engine, message, R, k_i, sigma_i):
    """`LocalSignature::phase5_local_sig` (party_i.rs:489-511): s_i = m k_i + r sigma_i"""
    _bind(eng.lib)
    count = len(k_i)
    ins = [ints_to_limbs(message, 8), _pts(R), ints_to_limbs(k_i, 8), ints_to_limbs(sigma_i, 8)]
    out = np.zeros((count, 8), np.uint32)
    eng._ck(eng.lib.tecdsa_gg18_local_sig_batch(eng._ctx, *[_ptr(a) for a in ins], _ptr(out), count, HOST), "gg18_local_sig")
    return limbs_to_ints(out)


def phase5a(eng: Engine, R, s_i, l_i, rho_i, blind, heg_s1, heg_s2, dlog_nonce):
    """`phase5a_broadcast_5b_zkproof` (party_i.rs:513-558) -> dict(com, decom[n][48] = V|A|B, heg[n][48], dlog[n][40], status)"""
    _bind(eng.lib)
    count = len(s_i)
    ins = [_pts(R)] + [ints_to_limbs(v, 8) for v in (s_i, l_i, rho_i, blind, heg_s1, heg_s2, dlog_nonce)]
    com, decom = np.zeros((count, 8), np.uint32), np.zeros((count, 48), np.uint32)
    heg, dlog = np.zeros((count, 48), np.uint32), np.zeros((count, 40), np.uint32)
    st = np.full(count, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_gg18_phase5a_batch(eng._ctx, *[_ptr(a) for a in ins], _ptr(com), _ptr(decom), _ptr(heg), _ptr(dlog), _ptr(st), count, HOST),
            "gg18_phase5a")
    return {"com": com, "decom": decom, "heg": heg, "dlog": dlog, "status": st}


def phase5c(eng: Engine, parties: int, R, y, message, rho_i, l_i, blind2, com: np.ndarray, decom: np.ndarray, blind, heg: np.ndarray, dlog: np.ndarray):
    """`phase5c` (party_i.rs:560-629) -> dict(com2, decom2[n][32] = u_i|t_i, status)"""
    _bind(eng.lib)
    count = len(rho_i)
    ins = [_pts(R), _pts(y), ints_to_limbs(message, 8), ints_to_limbs(rho_i, 8), ints_to_limbs(l_i, 8), ints_to_limbs(blind2, 8),
           np.ascontiguousarray(com, np.uint32), np.ascontiguousarray(decom, np.uint32), ints_to_limbs(blind, 8),
           np.ascontiguousarray(heg, np.uint32), np.ascontiguousarray(dlog, np.uint32)]
    com2, decom2 = np.zeros((count, 8), np.uint32), np.zeros((count, 32), np.uint32)
    st = np.full(count, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_gg18_phase5c_batch(eng._ctx, parties, *[_ptr(a) for a in ins], _ptr(com2), _ptr(decom2), _ptr(st), count // parties, HOST),
            "gg18_phase5c")
    return {"com2": com2, "decom2": decom2, "status": st}


def phase5d(eng: Engine, parties: int, decom2: np.ndarray, blind2, com2: np.ndarray, decom: np.ndarray) -> np.ndarray:
    """`phase5d` (party_i.rs:631-665) -> status (0 = Ok(s_i), 11 = InvalidCom, 2 = InvalidKey)"""
    _bind(eng.lib)
    count = decom2.shape[0]
    ins = [np.ascontiguousarray(decom2, np.uint32), ints_to_limbs(blind2, 8), np.ascontiguousarray(com2, np.uint32), np.ascontiguousarray(decom, np.uint32)]
    st = np.full(count, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_gg18_phase5d_batch(eng._ctx, parties, *[_ptr(a) for a in ins], _ptr(st), count // parties, HOST), "gg18_phase5d")
    return st


def output_signature(eng: Engine, parties: int, R, y, message, s_i):
    """`output_signature` (party_i.rs:666-703) with the in-tree `verify` (:706-730) -> (r, s, recid, status) per element"""
    _bind(eng.lib)
    count = len(s_i)
    ins = [_pts(R), _pts(y), ints_to_limbs(message, 8), ints_to_limbs(s_i, 8)]
    r, s = np.zeros((count, 8), np.uint32), np.zeros((count, 8), np.uint32)
    rec, st = np.zeros(count, np.uint8), np.full(count, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_gg18_output_signature_batch(eng._ctx, parties, *[_ptr(a) for a in ins], _ptr(r), _ptr(s), _ptr(rec), _ptr(st), count // parties, HOST),
            "gg18_output_signature")
    return limbs_to_ints(r), limbs_to_ints(s), rec, st
