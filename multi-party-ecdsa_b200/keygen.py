"""Host-side mirror of the key-generation VERIFICATION checks:
`NiCorrectKeyProof::verify`, `CompositeDLogProof::verify`, `VerifiableSS::validate_share` as called from
/root/reference/src/protocols/multi_party_ecdsa/gg_2020/party_i.rs:260-367, batched through the C ABI.

Status: written after the round's GPU budget was spent — NOT yet validated on a GPU (tests/test_keygen_gpu.py is
skipped unless TECDSA_EXPERIMENTAL=1)."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import HOST, Engine, ints_to_limbs, _ptr
from .gg20 import pack_point

SALT_STRING = bytes([75, 90, 101, 110])            # zk-paillier `SALT_STRING` [R]


def _bind(lib):
    if getattr(lib, "_keygen_bound", False):
        return
    V, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.tecdsa_correct_key_verify_batch.argtypes = [V, V, V, V, I, V, S, I]
    lib.tecdsa_composite_dlog_verify_batch.argtypes = [V, V, V, V, V, V, I, V, S, I]
    lib.tecdsa_vss_validate_share_batch.argtypes = [V, V, I, V, V, V, S, I]
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
