"""Host-side mirror of the key-generation path, batched through the C ABI:
`NiCorrectKeyProof::{proof,verify}`, `CompositeDLogProof::{prove,verify}`, `VerifiableSS::{share,validate_share}`,
`generate_h1_h2_N_tilde` and the per-phase checks of `Keys` as called from
/root/reference/src/protocols/multi_party_ecdsa/gg_2020/party_i.rs:137-156,219-438.  No arithmetic happens here: the functions
pack limbs, call the library and combine status bytes."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import HOST, Engine, ints_to_limbs, _ptr
from . import limbs_to_ints
from .gg20 import pack_point, unpack_point

SALT_STRING = bytes([75, 90, 101, 110])            # zk-paillier `SALT_STRING` [R]


def _bind(lib):
    if getattr(lib, "_keygen_bound", False):
        return
    V, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.tecdsa_correct_key_verify_batch.argtypes = [V, V, V, V, I, V, S, I]
    lib.tecdsa_composite_dlog_verify_batch.argtypes = [V, V, V, V, V, V, I, V, S, I]
    lib.tecdsa_vss_validate_share_batch.argtypes = [V, V, I, V, V, V, S, I]
    lib.tecdsa_correct_key_prove_batch.argtypes = [V, V, V, V, I, V, V, S, I]
    lib.tecdsa_composite_dlog_prove_batch.argtypes = [V, V, V, V, V, I, V, V, V, I, S, I]
    lib.tecdsa_vss_share_batch.argtypes = [V, I, I, V, V, V, S, I]
    lib.tecdsa_h1_h2_n_tilde_batch.argtypes = [V] * 10 + [S, I]
    lib._keygen_bound = True


def correct_key_verify(eng: Engine, n_list: Sequence[int], sigma_vecs: Sequence[Sequence[int]], salt: bytes = SALT_STRING) -> np.ndarray:
    """-> status[count] (0 = accept).  sigma_vecs[i] are the 11 values of proof i; any other length rejects on the host."""
    _bind(eng.lib)
    count = len(n_list)
    N = ints_to_limbs(n_list, 64)
    flat, short = [], np.zeros(count, dtype=bool)
    for i, sv in enumerate(sigma_vecs):
        if len(sv) != 11 or any(s < 0 or s >> 2048 for s in sv):
            short[i] = True
            sv = [0] * 11
        flat += list(sv)
    Sg = ints_to_limbs(flat, 64)
    st = np.full(count, 255, dtype=np.uint8)
    salt_arr = np.frombuffer(salt, dtype=np.uint8).copy() if salt else None
    eng._ck(eng.lib.tecdsa_correct_key_verify_batch(eng._ctx, _ptr(N), _ptr(Sg), _ptr(salt_arr) if salt_arr is not None else None, len(salt),
                                                    _ptr(st), count, HOST), "correct_key_verify")
    st[short] = 10                                   # TECDSA_ST_PROOF
    return st


def composite_dlog_verify(eng: Engine, statements: Sequence[Tuple[int, int, int]], proofs: Sequence[Tuple[int, int]], y_limbs: int = 92) -> np.ndarray:
    """statements[i] = (N, g, ni), proofs[i] = (x, y) -> status[count]"""
    _bind(eng.lib)
    count = len(statements)
    N = ints_to_limbs([s[0] for s in statements], 64)
    G = ints_to_limbs([s[1] % (1 << 2048) for s in statements], 64)
    NI = ints_to_limbs([s[2] % (1 << 2048) for s in statements], 64)
    bad = np.array([p[1] < 0 or p[1] >> (32 * y_limbs) != 0 or p[0] < 0 or p[0] >> 2048 != 0 for p in proofs], dtype=bool)
    X = ints_to_limbs([0 if b else p[0] for p, b in zip(proofs, bad)], 64)
    Y = ints_to_limbs([0 if b else p[1] for p, b in zip(proofs, bad)], y_limbs)
    st = np.full(count, 255, dtype=np.uint8)
    eng._ck(eng.lib.tecdsa_composite_dlog_verify_batch(eng._ctx, _ptr(N), _ptr(G), _ptr(NI), _ptr(X), _ptr(Y), y_limbs, _ptr(st), count, HOST),
            "composite_dlog_verify")
    st[bad] = 10
    return st


def vss_validate_share(eng: Engine, commitments: Sequence[Sequence], shares: Sequence[int], indices: Sequence[int]) -> np.ndarray:
    """commitments[i] = the t+1 coefficient commitments (oracle points) of scheme i -> status[count]"""
    _bind(eng.lib)
    count = len(shares)
    nc = len(commitments[0])
    assert all(len(c) == nc for c in commitments)
    C = ints_to_limbs([pack_point(p) for cs in commitments for p in cs], 16)
    Sh = ints_to_limbs([s % (1 << 256) for s in shares], 8)
    idx = np.asarray(indices, dtype=np.uint32)
    st = np.full(count, 255, dtype=np.uint8)
    eng._ck(eng.lib.tecdsa_vss_validate_share_batch(eng._ctx, _ptr(C), nc, _ptr(Sh), _ptr(idx), _ptr(st), count, HOST), "vss_validate_share")
    return st


# ----------------------------------------------------------------------------- prove side
def correct_key_prove(eng: Engine, pq: Sequence[Tuple[int, int]], salt: bytes = SALT_STRING):
    """`NiCorrectKeyProof::proof(&dk, None)` batched: pq[i] = (p, q) -> (list of 11-element sigma vectors, status)"""
    _bind(eng.lib)
    count = len(pq)
    P, Qv = ints_to_limbs([x[0] for x in pq], 32), ints_to_limbs([x[1] for x in pq], 32)
    sig = np.zeros((count * 11, 64), dtype=np.uint32)
    st = np.full(count, 255, dtype=np.uint8)
    salt_arr = np.frombuffer(salt, dtype=np.uint8).copy() if salt else None
    eng._ck(eng.lib.tecdsa_correct_key_prove_batch(eng._ctx, _ptr(P), _ptr(Qv), _ptr(salt_arr) if salt_arr is not None else None, len(salt),
                                                   _ptr(sig), _ptr(st), count, HOST), "correct_key_prove")
    flat = limbs_to_ints(sig)
    return [flat[11 * i:11 * i + 11] for i in range(count)], st


def composite_dlog_prove(eng: Engine, statements: Sequence[Tuple[int, int, int]], secrets: Sequence[int], nonces: Sequence[int],
                         secret_limbs: int = 64, y_limbs: int = 76):
    """`CompositeDLogProof::prove(&DLogStatement{N, g, ni}, &secret)` with the sampled nonce explicit -> list of (x, y)"""
    _bind(eng.lib)
    count = len(statements)
    N, G, NI = (ints_to_limbs([s[k] for s in statements], 64) for k in range(3))
    Sec, R = ints_to_limbs(secrets, secret_limbs), ints_to_limbs(nonces, 16)
    X, Y = np.zeros((count, 64), dtype=np.uint32), np.zeros((count, y_limbs), dtype=np.uint32)
    eng._ck(eng.lib.tecdsa_composite_dlog_prove_batch(eng._ctx, _ptr(N), _ptr(G), _ptr(NI), _ptr(Sec), secret_limbs, _ptr(R), _ptr(X), _ptr(Y), y_limbs,
                                                      count, HOST), "composite_dlog_prove")
    return list(zip(limbs_to_ints(X), limbs_to_ints(Y)))


def vss_share(eng: Engine, t: int, n: int, polynomials: Sequence[Sequence[int]]):
    """`VerifiableSS::share(t, n, &secret)` with explicit coefficients (polynomials[i][0] = the secret) ->
    (shares[i] = [f(1)..f(n)], commitments[i] = [(x, y) of a_j*G])"""
    _bind(eng.lib)
    count = len(polynomials)
    assert all(len(p) == t + 1 for p in polynomials)
    C = ints_to_limbs([c for p in polynomials for c in p], 8)
    Sh, Cm = np.zeros((count * n, 8), dtype=np.uint32), np.zeros((count * (t + 1), 16), dtype=np.uint32)
    eng._ck(eng.lib.tecdsa_vss_share_batch(eng._ctx, t, n, _ptr(C), _ptr(Sh), _ptr(Cm), count, HOST), "vss_share")
    sh, cm = limbs_to_ints(Sh), [unpack_point(v) for v in limbs_to_ints(Cm)]
    return [sh[i * n:(i + 1) * n] for i in range(count)], [cm[i * (t + 1):(i + 1) * (t + 1)] for i in range(count)]


def h1_h2_n_tilde(eng: Engine, setups: Sequence[Tuple[int, int, int, int]]):
    """`generate_h1_h2_N_tilde` with explicit samples: setups[i] = (p~, q~, h1, xhi) -> ([(N~, h1, h2, phi - xhi, phi - xhi^-1)], status)"""
    _bind(eng.lib)
    count = len(setups)
    P, Qv = ints_to_limbs([s[0] for s in setups], 32), ints_to_limbs([s[1] for s in setups], 32)
    H1, X = ints_to_limbs([s[2] for s in setups], 64), ints_to_limbs([s[3] for s in setups], 64)
    outs = [np.zeros((count, 64), dtype=np.uint32) for _ in range(4)]
    st = np.full(count, 255, dtype=np.uint8)
    eng._ck(eng.lib.tecdsa_h1_h2_n_tilde_batch(eng._ctx, _ptr(P), _ptr(Qv), _ptr(H1), _ptr(X), *[_ptr(o) for o in outs], _ptr(st), count, HOST), "h1_h2_n_tilde")
    nt, h2, xn, xin = (limbs_to_ints(o) for o in outs)
    return [(nt[i], setups[i][2], h2[i], xn[i], xin[i]) for i in range(count)], st


# ----------------------------------------------------------------------------- per-phase checks of `Keys` (batched over messages)
def phase1_verify(eng: Engine, broadcasts: Sequence, decommits: Sequence) -> np.ndarray:
    """The per-sender term of `phase1_verify_com_phase3_verify_correct_key_verify_dlog_phase2_distribute` (party_i.rs:272-305) for a
    batch of (KeyGenBroadcastMessage1, KeyGenDecommitMessage1) pairs (duck-typed like oracle.keygen_oracle's dataclasses):
    commitment opens, NiCorrectKeyProof verifies, both moduli are 2047..2048 bits, both CompositeDLogProofs verify.
    -> bool[count] (False = that sender is a `bad_actor`)."""
    from .gg20 import hash_commitment
    count = len(broadcasts)
    ok = np.ones(count, dtype=bool)
    com = hash_commitment(eng, [d.y_i for d in decommits], [d.blind_factor for d in decommits])
    ok &= np.array([c == b.com for c, b in zip(com, broadcasts)])
    ok &= correct_key_verify(eng, [b.e.n for b in broadcasts], [b.correct_key_proof for b in broadcasts]) == 0
    ok &= np.array([2047 <= b.e.n.bit_length() <= 2048 and 2047 <= b.dlog_statement.N.bit_length() <= 2048 for b in broadcasts])
    st1 = [(b.dlog_statement.N, b.dlog_statement.g, b.dlog_statement.ni) for b in broadcasts]
    st2 = [(b.dlog_statement.N, b.dlog_statement.ni, b.dlog_statement.g) for b in broadcasts]
    pf = [(b.composite_dlog_proof_base_h1.x, b.composite_dlog_proof_base_h1.y) for b in broadcasts] + \
         [(b.composite_dlog_proof_base_h2.x, b.composite_dlog_proof_base_h2.y) for b in broadcasts]
    cd = composite_dlog_verify(eng, st1 + st2, pf) == 0
    return ok & cd[:count] & cd[count:]


def phase2_verify_vss(eng: Engine, y_vec: Sequence, shares_for_me: Sequence[int], vss_vec: Sequence, index: int) -> np.ndarray:
    """The per-sender term of `phase2_verify_vss_construct_keypair_phase3_pok_dlog` (party_i.rs:334-349): validate_share and
    commitments[0] == y_i.  -> bool[count]"""
    ok = vss_validate_share(eng, [v.commitments for v in vss_vec], shares_for_me, [index] * len(shares_for_me)) == 0
    return ok & np.array([v.commitments[0] == y for v, y in zip(vss_vec, y_vec)])
