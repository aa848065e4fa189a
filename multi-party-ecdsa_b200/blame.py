"""Host-side mirror of GG20 identifiable abort: `GlobalStatePhase5::phase5_blame`,
`GlobalStatePhase6::{extract_paillier_randomness, ecddh_proof, phase6_blame}`, `GlobalStatePhase7::phase7_blame`
(/root/reference/src/protocols/multi_party_ecdsa/gg_2020/blame.rs:116-224, 252-271, 322-431, 434-454), written over the batched
C-ABI calls: every opened value of every signer is re-derived on the GPU in a handful of batch calls (Paillier encryptions,
homomorphic multiplications, point and scalar arithmetic, ECDDH proofs) and compared with what was broadcast.  Like the
reference, each function returns the list of bad signer positions (`ErrorType.bad_actors`); no arithmetic happens on the host —
only index bookkeeping and equality tests of the returned values."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np

from . import HOST, Engine, ints_to_limbs, limbs_to_ints, _ptr
from .gg20 import KeySets, _pts, unpack_point

Point = Tuple[int, int]


def _bind(lib):
    if getattr(lib, "_blame_bound", False):
        return
    V, S, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.tecdsa_paillier_open_batch.argtypes = [V, V, V, V, V, V, S, I]
    lib.tecdsa_ecddh_prove_batch.argtypes = [V] * 8 + [S, I]
    lib.tecdsa_ecddh_verify_batch.argtypes = [V] * 7 + [S, I]
    lib._blame_bound = True


# ----------------------------------------------------------------------------- new primitives
def paillier_open(eng: Engine, keys: KeySets, key_rows: Sequence[int], c: Sequence[int]):
    """Batched `Paillier::open(dk, c)` (blame.rs:252-256): -> (plaintexts, randomness) with c = (1 + m n) r^n mod n^2"""
    _bind(eng.lib)
    rows = np.asarray(key_rows, dtype=np.uint32)
    C = ints_to_limbs(c, 128)
    m, r = np.zeros((len(c), 64), np.uint32), np.zeros((len(c), 64), np.uint32)
    eng._ck(eng.lib.tecdsa_paillier_open_batch(eng._ctx, keys.handle, _ptr(rows), _ptr(C), _ptr(m), _ptr(r), len(c), HOST), "paillier_open")
    return limbs_to_ints(m), limbs_to_ints(r)


def ecddh_prove(eng: Engine, x: Sequence[int], g1, h1, g2, h2, nonce: Sequence[int]) -> np.ndarray:
    """Batched curv `ECDDHProof::prove` for statements (g1, h1 = x g1, g2, h2 = x g2) with the sampled nonce explicit
    (blame.rs:258-271) -> [n][40] uint32: a1 16 | a2 16 | z 8"""
    _bind(eng.lib)
    n = len(x)
    ins = [ints_to_limbs(x, 8), _pts(g1), _pts(h1), _pts(g2), _pts(h2), ints_to_limbs(nonce, 8)]
    out = np.zeros((n, 40), np.uint32)
    eng._ck(eng.lib.tecdsa_ecddh_prove_batch(eng._ctx, *[_ptr(a) for a in ins], _ptr(out), n, HOST), "ecddh_prove")
    return out


def ecddh_verify(eng: Engine, proofs: np.ndarray, g1, h1, g2, h2) -> np.ndarray:
    """Batched curv `ECDDHProof::verify` -> status[n] (0 = accept)"""
    _bind(eng.lib)
    n = proofs.shape[0]
    ins = [np.ascontiguousarray(proofs, dtype=np.uint32), _pts(g1), _pts(h1), _pts(g2), _pts(h2)]
    st = np.full(n, 255, np.uint8)
    eng._ck(eng.lib.tecdsa_ecddh_verify_batch(eng._ctx, *[_ptr(a) for a in ins], _ptr(st), n, HOST), "ecddh_verify")
    return st


# ----------------------------------------------------------------------------- phase 5
def phase5_blame(eng: Engine, n_list: Sequence[int], k_vec, k_randomness_vec, gamma_vec, beta_randomness_vec, beta_tag_vec, delta_vec,
                 g_gamma_vec, m_a_c, m_b_c) -> List[int]:
    """`GlobalStatePhase5::phase5_blame` (blame.rs:116-224).  n_list[i] = Paillier modulus of signer i; the *_vec arguments are
    the fields of GlobalStatePhase5 (beta_* already re-indexed per Alice, :79-101); m_a_c[i] / m_b_c[i][j] the broadcast
    ciphertexts."""
    n = len(delta_vec)
    bad: List[int] = []
    # g_gamma_i == gamma_i * G
    gg = eng.secp_mul(None, list(gamma_vec))
    bad += [i for i in range(n) if gg[i] != g_gamma_vec[i]]
    # every ciphertext is re-derived up front in a few batch calls; the reference's sequential bookkeeping (rows are only
    # examined while nobody has been blamed yet, :139) is then replayed on the precomputed values
    c_a = eng.paillier_encrypt(list(n_list), list(range(n)), list(k_vec), list(k_randomness_vec))
    pairs = [(i, j) for i in range(n) for j in range(n - 1)]
    ind = [j if j < i else j + 1 for i, j in pairs]
    enc = eng.paillier_encrypt(list(n_list), [i for i, _ in pairs], [beta_tag_vec[i][j] for i, j in pairs], [beta_randomness_vec[i][j] for i, j in pairs])
    mul = eng.paillier_mul(list(n_list), [i for i, _ in pairs], [c_a[i] for i, _ in pairs], [gamma_vec[x] for x in ind], 8)
    c_b = eng.paillier_add(list(n_list), [i for i, _ in pairs], mul, enc)
    beta_fe = eng.scalar_from_bigint([beta_tag_vec[i][j] for i, j in pairs], 64)
    beta = eng.scalar_op("sub", [0] * len(pairs), beta_fe)
    kg = eng.scalar_op("mul", [k_vec[i] for i, _ in pairs], [gamma_vec[x] for x in ind])
    alpha = eng.scalar_op("sub", kg, beta)
    ab = {}
    for i in range(n):
        if c_a[i] != m_a_c[i]:
            bad.append(i)
        if not bad:
            for t, (pi, j) in enumerate(pairs):
                if pi != i:
                    continue
                if c_b[t] != m_b_c[i][j]:
                    bad.append(ind[t])
                ab[(i, j)] = (alpha[t], beta[t])
    if not bad:
        kgi = eng.scalar_op("mul", list(k_vec), list(gamma_vec))
        for i in range(n):
            acc = [kgi[i]] + [ab[(i, j)][0] for j in range(n - 1)]
            for j in range(n - 1):
                ind1, ind2 = (j, i - 1) if j < i else (j + 1, i)
                acc.append(ab[(ind1, ind2)][1])
            total = acc[0]
            for v in acc[1:]:
                total = eng.scalar_op("add", [total], [v])[0]
            if total != delta_vec[i]:
                bad.append(i)
    return sorted(set(bad))


# ----------------------------------------------------------------------------- phase 6
def phase6_blame(eng: Engine, n_list: Sequence[int], k_vec, k_randomness_vec, miu_vec, miu_randomness_vec, g_w_vec, proof_vec: np.ndarray, S_vec,
                 m_a_c, m_b_c, R) -> List[int]:
    """`GlobalStatePhase6::phase6_blame(&R)` (blame.rs:322-431); proof_vec = [n][40] ECDDH proofs (ecddh_prove layout)"""
    n = len(k_vec)
    bad: List[int] = []
    pairs = [(i, j) for i in range(n) for j in range(n - 1)]
    enc = eng.paillier_encrypt(list(n_list), [i for i, _ in pairs], [miu_vec[i][j] for i, j in pairs], [miu_randomness_vec[i][j] for i, j in pairs])
    bad += [i for t, (i, j) in enumerate(pairs) if enc[t] != m_b_c[i][j]]
    c_a = eng.paillier_encrypt(list(n_list), list(range(n)), list(k_vec), list(k_randomness_vec))
    bad += [i for i in range(n) if c_a[i] != m_a_c[i]]
    if not bad:
        miu_fe = eng.scalar_from_bigint([miu_vec[i][j] for i, j in pairs], 64)
        g_miu = eng.secp_mul(None, miu_fe)
        gwk = eng.secp_mul([g_w_vec[j if j < i else j + 1] for i, j in pairs], [k_vec[i] for i, _ in pairs])
        g_ni = dict(zip(pairs, eng.point_add(gwk, g_miu, subtract=True)))
        g_miu_d = dict(zip(pairs, g_miu))
        g_sigma = eng.secp_mul(list(g_w_vec), list(k_vec))
        for i in range(n):
            terms = [g_miu_d[(i, j)] for j in range(n - 1)]
            for j in range(n - 1):
                ind1, ind2 = (j, i - 1) if j < i else (j + 1, i)
                terms.append(g_ni[(ind1, ind2)])
            for t in terms:
                g_sigma[i] = eng.point_add([g_sigma[i]], [t])[0]
        G = eng.secp_mul(None, [1])[0]
        st = ecddh_verify(eng, proof_vec, [G] * n, g_sigma, [R] * n, list(S_vec))
        bad += [i for i in range(n) if st[i] != 0]
    return sorted(set(bad))


# ----------------------------------------------------------------------------- phase 7
def phase7_blame(eng: Engine, s_vec, r: int, R_dash_vec, m: int, R, S_vec) -> List[int]:
    """`GlobalStatePhase7::phase7_blame` (blame.rs:434-454): R s_i == R_dash_i m + S_i r"""
    n = len(s_vec)
    left = eng.secp_mul([R] * n, list(s_vec))
    m_fe = eng.scalar_from_bigint([m], 64)[0]
    right = eng.point_add(eng.secp_mul(list(R_dash_vec), [m_fe] * n), eng.secp_mul(list(S_vec), [r] * n))
    return [i for i in range(n) if left[i] != right[i]]
