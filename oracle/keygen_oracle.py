"""CPU oracle for SURVEY.md section 8(f) rank 1: the GG20 key-generation VERIFICATION path.

TEST INFRASTRUCTURE ONLY (same rule as gg20_oracle.py): nothing outside tests/ may import this.  The CUDA entry
points built against it: csrc/keygen.cu, multi-party-ecdsa_b200/keygen.py, tests/test_keygen_gpu.py.

PARITY UNPINNED.  The in-tree callers are cited by file:line under /root/reference; the proofs themselves live in
crates that are NOT vendored (zk-paillier 0.4.3: `NiCorrectKeyProof`, `CompositeDLogProof`; curv-kzen 0.9:
`VerifiableSS`), so their bodies are restated from the published algorithms and marked [R] (recalled, not
verifiable in this container).  What the tests pin instead: prove -> verify accepts, every tampered field rejects,
Feldman shares interpolate to the secret and match the committed fixtures (tests/golden/keys_t1n3.json).

Randomness is always an explicit argument (the reference samples inside the functions).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gcd
from typing import List, Optional, Sequence, Tuple

from oracle import gg20_oracle as o

Q = o.Q
Point = o.Point

# zk-paillier 0.4.3 zkproofs/correct_key_ni.rs [R]
SALT_STRING = bytes([75, 90, 101, 110])         # "KZen"
M2 = 11
DIGEST_SIZE = 256
ALPHA = 6379                                     # eprint 2018/987 section 6.2.3: alpha = 6379, m2 = 11


def _primorial(alpha: int) -> int:
    sieve = bytearray([1]) * (alpha + 1)
    sieve[0:2] = b"\0\0"
    for i in range(2, int(alpha ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    out = 1
    for i in range(alpha + 1):
        if sieve[i]:
            out *= i
    return out


# `const P: &str = "1824183726245393467247644231302244136...8238796796690"` [R]: the product of all primes <= 6379 (2746 decimal
# digits); verify() requires gcd(P, n) == 1, which is what bounds the soundness error of the proof (an N with a prime factor
# below alpha is rejected outright).  Round 1 of this repo wrongly used the number 6370 itself.
PRIMORIAL = _primorial(ALPHA)
assert str(PRIMORIAL).startswith("182418372624539346724764423130") and str(PRIMORIAL).endswith("88238796796690")
# zk-paillier 0.4.3 zkproofs/composite_dlog_proof.rs [R]
K_BITS, K_PRIME_BITS, SAMPLE_S_BITS = 128, 128, 256
# gg_2020/party_i.rs:49-50
PAILLIER_MIN_BIT_LENGTH, PAILLIER_MAX_BIT_LENGTH = 2047, 2048


def compute_digest(items: Sequence[int]) -> int:
    """zk-paillier `compute_digest`: SHA-256 over the concatenated `BigInt::to_bytes()` of the items [R]"""
    import hashlib
    h = hashlib.sha256()
    for x in items:
        h.update(o.bn_bytes(x))
    return int.from_bytes(h.digest(), "big")


def mask_generation(out_length: int, seed: int) -> int:
    """correct_key_ni.rs `mask_generation` [R]: sum_j H(seed, j) << (256 j) for j < out_length/256 + 1"""
    msklen = out_length // DIGEST_SIZE + 1
    return sum(compute_digest([seed, j]) << (j * DIGEST_SIZE) for j in range(msklen))


def _rho_vec(n: int, salt: bytes) -> List[int]:
    key_length = n.bit_length()
    salt_bn = int.from_bytes(salt, "big")
    return [mask_generation(key_length, compute_digest([n, salt_bn, i])) % n for i in range(M2)]


# ------------------------------------------------------------------------------------------- NiCorrectKeyProof
def correct_key_proof(dk: o.DecryptionKey, salt: bytes = SALT_STRING) -> List[int]:
    """`NiCorrectKeyProof::proof(&dk, None)` (call site gg_2020/party_i.rs:225): sigma_i = rho_i^(N^-1 mod phi) mod N,
    i < 11 — eleven full-width modular exponentiations modulo N [R]"""
    n = dk.p * dk.q
    phi = (dk.p - 1) * (dk.q - 1)
    n_inv = pow(n, -1, phi)
    return [pow(rho, n_inv, n) for rho in _rho_vec(n, salt)]


def correct_key_verify(sigma_vec: Sequence[int], ek: o.EncryptionKey, salt: bytes = SALT_STRING) -> bool:
    """`NiCorrectKeyProof::verify(&ek, SALT_STRING)` (call site party_i.rs:288-291): gcd(P, N) == 1 and
    sigma_i^N == rho_i (mod N) for all i [R]; P = the primorial of all primes <= 6379"""
    if len(sigma_vec) != M2 or gcd(PRIMORIAL, ek.n) != 1:
        return False
    return all(pow(s, ek.n, ek.n) == rho for s, rho in zip(sigma_vec, _rho_vec(ek.n, salt)))


# ------------------------------------------------------------------------------------------- CompositeDLogProof
@dataclass
class CompositeDLogProof:
    x: int
    y: int


def composite_dlog_prove(st: o.DLogStatement, secret: int, r: int) -> CompositeDLogProof:
    """`CompositeDLogProof::prove(&statement, &secret)` (call sites party_i.rs:238-241); r < 2^(128+128+256) is the
    sampled nonce.  x = g^r mod N, e = H(x, g, N, ni), y = r + e*secret over the integers [R]"""
    assert 0 <= r < 1 << (K_BITS + K_PRIME_BITS + SAMPLE_S_BITS)
    x = pow(st.g, r, st.N)
    e = compute_digest([x, st.g, st.N, st.ni])
    return CompositeDLogProof(x, r + e * secret)


def composite_dlog_verify(pf: CompositeDLogProof, st: o.DLogStatement) -> bool:
    """`CompositeDLogProof::verify(&statement)` (call sites party_i.rs:296-303): N > 2^128, gcd(g, N) = gcd(ni, N) = 1,
    x == g^y * ni^e mod N [R] (the reference `assert!`s the first three; a batch engine reports them as a reject)"""
    if st.N <= 1 << K_BITS or gcd(st.g, st.N) != 1 or gcd(st.ni, st.N) != 1 or pf.y < 0:
        return False
    e = compute_digest([pf.x, st.g, st.N, st.ni])
    return pf.x == pow(st.g, pf.y, st.N) * pow(st.ni, e, st.N) % st.N


def h1_h2_n_tilde(p_t: int, q_t: int, h1: int, xhi: int) -> Tuple[int, int, int, int, int]:
    """`generate_h1_h2_N_tilde` (party_i.rs:137-156) with its samples explicit: returns (N_tilde, h1, h2, xhi', xhi_inv')
    where the primed values are the NEGATED exponents the function hands back (phi - xhi, phi - xhi^-1)"""
    nt = p_t * q_t
    phi = (p_t - 1) * (q_t - 1)
    xhi_inv = pow(xhi, -1, phi)
    h2 = pow(h1, xhi, nt)
    return nt, h1, h2, phi - xhi, phi - xhi_inv


# ------------------------------------------------------------------------------------------- Feldman VSS (curv) [R]
@dataclass
class VerifiableSS:
    threshold: int
    share_count: int
    commitments: List[Point]


def vss_share(t: int, n: int, secret: int, coefficients: Sequence[int]) -> Tuple[VerifiableSS, List[int]]:
    """`VerifiableSS::share(t, n, &secret)` (call site party_i.rs:313): f(x) = secret + sum a_j x^j (j = 1..t), shares f(1..n),
    commitments [f_j * G]"""
    assert len(coefficients) == t
    poly = [secret % Q] + [c % Q for c in coefficients]
    shares = [sum(c * pow(i, j, Q) for j, c in enumerate(poly)) % Q for i in range(1, n + 1)]
    return VerifiableSS(t, n, [o.pt_mul(o.G, c) for c in poly]), shares


def vss_point_commitment(vss: VerifiableSS, index: int) -> Point:
    """`get_point_commitment(index)`: sum_j index^j * C_j  (Horner from the top, as curv does)"""
    acc = None
    for c in reversed(vss.commitments):
        acc = o.pt_add(o.pt_mul(acc, index % Q), c) if acc is not None else c
    return acc


def vss_validate_share(vss: VerifiableSS, share: int, index: int) -> bool:
    """`validate_share(&share, index)` (call site party_i.rs:337-339)"""
    return o.pt_mul(o.G, share) == vss_point_commitment(vss, index)


# ------------------------------------------------------------------------------------------- keygen phases
@dataclass
class KeyGenBroadcast1:
    """`KeyGenBroadcastMessage1` (party_i.rs:96-104)"""
    e: o.EncryptionKey
    dlog_statement: o.DLogStatement
    com: int
    correct_key_proof: List[int]
    composite_dlog_proof_base_h1: CompositeDLogProof
    composite_dlog_proof_base_h2: CompositeDLogProof


@dataclass
class KeyGenDecommit1:
    blind_factor: int
    y_i: Point


def phase1_broadcast(dk: o.DecryptionKey, nt: int, h1: int, h2: int, xhi: int, xhi_inv: int, y_i: Point, blind: int,
                     r1: int, r2: int) -> Tuple[KeyGenBroadcast1, KeyGenDecommit1]:
    """`phase1_broadcast_phase3_proof_of_correct_key_proof_of_correct_h1h2` (party_i.rs:219-258); xhi, xhi_inv are the
    (negated) exponents returned by h1_h2_n_tilde, r1/r2 the two CompositeDLog nonces"""
    st1 = o.DLogStatement(nt, h1, h2)
    st2 = o.DLogStatement(nt, h2, h1)
    n = dk.p * dk.q
    bc = KeyGenBroadcast1(o.EncryptionKey(n, n * n), st1,
                          o.hash_commitment(o.bn_from_bytes(o.pt_compress(y_i)), blind),
                          correct_key_proof(dk), composite_dlog_prove(st1, xhi, r1), composite_dlog_prove(st2, xhi_inv, r2))
    return bc, KeyGenDecommit1(blind, y_i)


def phase1_verify(bc: KeyGenBroadcast1, dec: KeyGenDecommit1) -> bool:
    """one party's term of `phase1_verify_com_phase3_verify_correct_key_verify_dlog_phase2_distribute` (party_i.rs:272-305)"""
    st = bc.dlog_statement
    st2 = o.DLogStatement(st.N, st.ni, st.g)
    return (o.hash_commitment(o.bn_from_bytes(o.pt_compress(dec.y_i)), dec.blind_factor) == bc.com
            and correct_key_verify(bc.correct_key_proof, bc.e)
            and PAILLIER_MIN_BIT_LENGTH <= bc.e.n.bit_length() <= PAILLIER_MAX_BIT_LENGTH
            and PAILLIER_MIN_BIT_LENGTH <= st.N.bit_length() <= PAILLIER_MAX_BIT_LENGTH
            and composite_dlog_verify(bc.composite_dlog_proof_base_h1, st)
            and composite_dlog_verify(bc.composite_dlog_proof_base_h2, st2))


def phase2_verify_vss(y_vec: Sequence[Point], shares_for_me: Sequence[int], vss_vec: Sequence[VerifiableSS], index: int,
                      nonce: int) -> Optional[Tuple[Point, int, o.DLogProof]]:
    """`phase2_verify_vss_construct_keypair_phase3_pok_dlog` (party_i.rs:322-367): returns (y, x_i, DLogProof(x_i)) or None
    ("invalid vss")"""
    for y_i, s, vss in zip(y_vec, shares_for_me, vss_vec):
        if not (vss_validate_share(vss, s, index) and vss.commitments[0] == y_i):
            return None
    y = None
    for p in y_vec:
        y = o.pt_add(y, p) if y is not None else p
    x_i = sum(shares_for_me) % Q
    return y, x_i, o.dlog_prove(x_i, nonce)


def commitments_to_xi(vss_vec: Sequence[VerifiableSS]) -> List[Point]:
    """`get_commitments_to_xi` (party_i.rs:369-388): X_i = sum over parties of their polynomial commitments at i"""
    n = len(vss_vec)
    glob = list(vss_vec[0].commitments)
    for vss in vss_vec[1:]:
        glob = [o.pt_add(a, b) for a, b in zip(glob, vss.commitments)]
    g = VerifiableSS(vss_vec[0].threshold, vss_vec[0].share_count, glob)
    return [vss_point_commitment(g, i) for i in range(1, n + 1)]


def verify_dlog_proofs_check_against_vss(proofs: Sequence[o.DLogProof], vss_vec: Sequence[VerifiableSS]) -> bool:
    """`verify_dlog_proofs_check_against_vss` (party_i.rs:405-438)"""
    xi = commitments_to_xi(vss_vec)
    return all(o.dlog_verify(pf) and xi[i] == pf.pk for i, pf in enumerate(proofs))
