"""ctypes front end of the oracle's C twin (oracle/gg20_twin.c, GMP + OpenSSL) — TEST INFRASTRUCTURE ONLY.

`offline_batch` takes the very arrays tecdsa_gg20_offline_batch takes (key tables as the limb arrays of `tecdsa_keys`,
sessions [n,3] uint32, randomness records [2n, 1408] uint32) and returns the same per-unit outputs, so a GPU batch and the CPU
twin are compared array against array.  Used by tests/, __graft_entry__.smoke() and bench.py (parity checker, cpu_baseline,
--impl reference)."""
from __future__ import annotations

import ctypes
import os
from typing import Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgg20_ref.so")
RND_LIMBS = 1408
_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        l = ctypes.CDLL(LIB_PATH)
        l.oracle_gmp_version.restype = ctypes.c_char_p
        V = ctypes.c_void_p
        l.oracle_gg20_offline_batch.argtypes = [ctypes.c_size_t] + [V] * 9 + [ctypes.c_size_t] + [V] * 7 + [ctypes.c_int]
        l.oracle_gg20_offline_batch.restype = ctypes.c_int
        _lib = l
    return _lib


def _limbs(vals: Sequence[int], k: int) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(4 * k, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), k).copy()


def _pt(p) -> int:
    return 0 if p is None else p[0] | (p[1] << 256)


class KeyTables:
    """The limb arrays of `tecdsa_keys` for a list of (t=1,n=3) key sets (oracle.LocalKey triples)."""

    def __init__(self, keysets: Sequence[Sequence]):
        rows = [lk for ks in keysets for lk in ks]
        self.n_keysets = len(keysets)
        self.p = _limbs([lk.dk.p for lk in rows], 32)
        self.q = _limbs([lk.dk.q for lk in rows], 32)
        self.nt = _limbs([lk.h1_h2_n_tilde_vec[lk.i - 1].N for lk in rows], 64)
        self.h1 = _limbs([lk.h1_h2_n_tilde_vec[lk.i - 1].g for lk in rows], 64)
        self.h2 = _limbs([lk.h1_h2_n_tilde_vec[lk.i - 1].ni for lk in rows], 64)
        self.x = _limbs([lk.x_i for lk in rows], 8)
        self.pk = _limbs([_pt(lk.pk_vec[lk.i - 1]) for lk in rows], 16)
        self.y = _limbs([_pt(ks[0].y_sum_s) for ks in keysets], 16)


class CpuResult:
    def __init__(self, status, R, sigma, k, t_vec, digest):
        self.status, self.R, self.sigma, self.k, self.t_vec, self.digest = status, R, sigma, k, t_vec, digest


def offline_batch(keys: KeyTables, sessions: np.ndarray, rnd: np.ndarray, threads: int = 1) -> CpuResult:
    sessions = np.ascontiguousarray(sessions, dtype=np.uint32).reshape(-1, 3)
    n = sessions.shape[0]
    rnd = np.ascontiguousarray(rnd, dtype=np.uint32)
    assert rnd.shape == (2 * n, RND_LIMBS)
    U = 2 * n
    status = np.full(U, 255, dtype=np.uint8)
    R, sigma, k = np.zeros((U, 16), np.uint32), np.zeros((U, 8), np.uint32), np.zeros((U, 8), np.uint32)
    tvec, digest = np.zeros((U, 32), np.uint32), np.zeros((U, 8), np.uint32)
    a = lambda x: x.ctypes.data
    rc = lib().oracle_gg20_offline_batch(keys.n_keysets, a(keys.p), a(keys.q), a(keys.nt), a(keys.h1), a(keys.h2), a(keys.x), a(keys.pk), a(keys.y),
                                         a(sessions), n, a(rnd), a(status), a(R), a(sigma), a(k), a(tvec), a(digest), threads)
    if rc != 0:
        raise ValueError("oracle_gg20_offline_batch: bad session descriptor")
    return CpuResult(status, R, sigma, k, tvec, digest)
