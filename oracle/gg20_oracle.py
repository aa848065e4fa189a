"""CPU oracle for the GG20 offline-signing arithmetic path — TEST INFRASTRUCTURE ONLY.

This file restates, with Python integers, the algorithm of the reference
(ZenGo-X/multi-party-ecdsa @ 7d8bd41, /root/reference) for the one hot path this repo
accelerates.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import it; the product (multi-party-ecdsa_b200/) never does.

PARITY STATUS: "parity unpinned" for encodings.  The reference ships no golden vectors, KATs
or fixtures for this path and cannot be built here (no cargo/rustc, no
vendored crates).  What IS pinned: every arithmetic primitive below is cross-checked against
independent implementations in tests/ (GMP mpz_powm via ctypes, hashlib SHA-256, OpenSSL /
`cryptography` secp256k1, FIPS/SEC known answers, `base_point2` = SHA256^3(compressed G)), and
the protocol-level invariants the reference's own tests assert (proof generate->verify,
alpha+beta = a*b, sum R_dash = G, sum S_i = y, final ECDSA signature verifies under an
independent verifier).  What is recalled, not verified ([R] in SURVEY.md Appendix C): the
byte encodings used by the out-of-tree crates (curv-kzen 0.9, kzen-paillier 0.4.2) —
`BigInt::to_bytes` of zero, `chain_point` = 65-byte uncompressed SEC1, `result_scalar`
= digest mod q, sign conventions of the sigma-proof responses.  Each such assumption is
marked [R] at its definition and isolated so it can be flipped in one place.

Every function cites the reference lines it follows.  All randomness is an explicit input
(the reference samples inside; its `*_with_predefined_randomness` variants show the seam:
src/utilities/mta/mod.rs:62,111).
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------- secp256k1
P = 2**256 - 2**32 - 977
Q = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
# curv `Point::<Secp256k1>::base_point2()` (used at gg_2020/party_i.rs:629,787,812).
# x = SHA256(SHA256(SHA256(compressed G))) — verified in tests/test_oracle.py.
H2X = 0x08D13221E3A7326A34DD45214BA80116DD142E4B5FF3CE66A8DC7BFA0378B795
H2Y = 0x5D41AC1477614B5C0848D50DBD565EA2807BCBA1DF0DF07A8217E9F7F7C2BE88

Point = Optional[Tuple[int, int]]          # None = identity
G: Point = (GX, GY)
H2: Point = (H2X, H2Y)


def pt_add(a: Point, b: Point) -> Point:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def pt_neg(a: Point) -> Point:
    return None if a is None else (a[0], (-a[1]) % P)


def pt_sub(a: Point, b: Point) -> Point:
    return pt_add(a, pt_neg(b))


def pt_mul(a: Point, k: int) -> Point:
    """`Point * Scalar` (scalar already reduced mod q)."""
    k %= Q
    if k == 0 or a is None:
        return None
    # Jacobian double-and-add
    X, Y, Z = a[0], a[1], 1
    RX, RY, RZ = 0, 1, 0
    for bit in bin(k)[2:]:
        # double R
        if RZ:
            S = 4 * RX * RY * RY % P
            M = 3 * RX * RX % P
            nx = (M * M - 2 * S) % P
            ny = (M * (S - nx) - 8 * pow(RY, 4, P)) % P
            nz = 2 * RY * RZ % P
            RX, RY, RZ = nx, ny, nz
        if bit == "1":
            if not RZ:
                RX, RY, RZ = X, Y, Z
            else:
                # mixed add R + (X,Y)
                Z2 = RZ * RZ % P
                U2 = X * Z2 % P
                S2 = Y * Z2 * RZ % P
                Hh = (U2 - RX) % P
                Rr = (S2 - RY) % P
                if Hh == 0:
                    if Rr == 0:
                        S = 4 * RX * RY * RY % P
                        M = 3 * RX * RX % P
                        nx = (M * M - 2 * S) % P
                        ny = (M * (S - nx) - 8 * pow(RY, 4, P)) % P
                        nz = 2 * RY * RZ % P
                        RX, RY, RZ = nx, ny, nz
                    else:
                        RX, RY, RZ = 0, 1, 0
                    continue
                H2_ = Hh * Hh % P
                H3 = H2_ * Hh % P
                nx = (Rr * Rr - H3 - 2 * RX * H2_) % P
                ny = (Rr * (RX * H2_ - nx) - RY * H3) % P
                nz = RZ * Hh % P
                RX, RY, RZ = nx, ny, nz
    if not RZ:
        return None
    zi = pow(RZ, -1, P)
    return (RX * zi * zi % P, RY * zi * zi * zi % P)


def pt_compress(a: Point) -> bytes:
    """`Point::to_bytes(true)`: 33-byte SEC1 compressed."""
    assert a is not None
    return bytes([2 + (a[1] & 1)]) + a[0].to_bytes(32, "big")


def pt_uncompressed(a: Point) -> bytes:
    """`Point::to_bytes(false)`: 65-byte SEC1 uncompressed."""
    assert a is not None
    return b"\x04" + a[0].to_bytes(32, "big") + a[1].to_bytes(32, "big")


# --------------------------------------------------------------------------- BigInt encodings
def bn_bytes(x: int) -> bytes:
    """curv `BigInt::to_bytes()`: big-endian magnitude, minimal length; zero -> one 0x00 byte
    [R: rust-gmp `From<&Mpz> for Vec<u8>` sizes by mpz_sizeinbase(x,2), which is 1 for 0]."""
    x = abs(x)
    return x.to_bytes(max(1, (x.bit_length() + 7) // 8), "big")


def bn_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "big")


def sha256_bigints(items: Sequence[int]) -> int:
    """`Sha256::new().chain_bigint(a).chain_bigint(b)....result_bigint()` — plain concatenation of
    to_bytes() outputs, digest read as a big-endian integer [R] (curv DigestExt)."""
    h = hashlib.sha256()
    for it in items:
        h.update(bn_bytes(it))
    return int.from_bytes(h.digest(), "big")


def sha256_points_scalar(points: Sequence[Point]) -> int:
    """`H::new().chain_points([...]).result_scalar()`: each point as 65-byte uncompressed SEC1 [R],
    digest reduced mod q [R]."""
    h = hashlib.sha256()
    for p in points:
        h.update(pt_uncompressed(p))
    return int.from_bytes(h.digest(), "big") % Q


def hash_commitment(message: int, blind: int) -> int:
    """curv `HashCommitment::<Sha256>::create_commitment_with_user_defined_randomness` [R]:
    SHA256(bytes(m) || bytes(r)) as BigInt.  Call sites gg_2020/party_i.rs:577-580,654-659."""
    return int.from_bytes(hashlib.sha256(bn_bytes(message) + bn_bytes(blind)).digest(), "big")


# --------------------------------------------------------------------------- Paillier (kzen-paillier 0.4.2 [R])
@dataclass
class EncryptionKey:
    n: int
    nn: int


@dataclass
class DecryptionKey:
    p: int
    q: int


def paillier_encrypt(ek: EncryptionKey, m: int, r: int) -> int:
    """`Paillier::encrypt_with_chosen_randomness`: c = (1 + m*n) * r^n mod n^2 (mta/mod.rs:68,133)."""
    rn = pow(r, ek.n, ek.nn)
    gm = (m * ek.n + 1) % ek.nn
    return gm * rn % ek.nn


def paillier_mul(ek: EncryptionKey, c: int, m: int) -> int:
    """`Paillier::mul`: c^m mod n^2 (mta/mod.rs:140)."""
    return pow(c, m, ek.nn)


def paillier_add(ek: EncryptionKey, c1: int, c2: int) -> int:
    """`Paillier::add`: c1*c2 mod n^2 (mta/mod.rs:145)."""
    return c1 * c2 % ek.nn


def paillier_decrypt(dk: DecryptionKey, c: int) -> int:
    """`Paillier::decrypt` (CRT form, kzen-paillier core.rs [R]); call site mta/mod.rs:165.
    The canonical plaintext does not depend on the representative choices inside."""
    p, q = dk.p, dk.q
    pp, qq, n = p * p, q * q, p * q

    def l(u: int, m: int) -> int:
        return (u - 1) // m

    def h(pr: int, prpr: int) -> int:
        gp = (1 - n) % prpr
        return pow(l(gp, pr), -1, pr)

    hp, hq = h(p, pp), h(q, qq)
    mp = l(pow(c % pp, p - 1, pp), p) * hp % p
    mq = l(pow(c % qq, q - 1, qq), q) * hq % q
    pinv = pow(p, -1, q)
    u = (mq - mp) % q * pinv % q
    return mp + u * p


# --------------------------------------------------------------------------- zk-paillier statement
@dataclass
class DLogStatement:
    """zk_paillier::zkproofs::DLogStatement {N, g, ni} = (N_tilde, h1, h2)."""
    N: int
    g: int
    ni: int


# --------------------------------------------------------------------------- MtA range proof (Alice)
@dataclass
class AliceProof:
    z: int
    e: int
    s: int
    s1: int
    s2: int


def alice_proof_generate(a: int, cipher: int, ek: EncryptionKey, st: DLogStatement, r: int,
                         alpha: int, beta: int, gamma: int, ro: int) -> AliceProof:
    """`AliceProof::generate` range_proofs.rs:160-193 (round1 :39-67, round2 :77-91).
    alpha < q^3, beta in Z*_N, gamma < q^3*N_tilde, ro < q*N_tilde are the values the reference samples."""
    h1, h2, Nt = st.g, st.ni, st.N
    z = pow(h1, a, Nt) * pow(h2, ro, Nt) % Nt                           # :52
    u = (alpha * ek.n + 1) * pow(beta, ek.n, ek.nn) % ek.nn             # :53-55
    w = pow(h1, alpha, Nt) * pow(h2, gamma, Nt) % Nt                    # :56-57
    e = sha256_bigints([ek.n, ek.n + 1, cipher, z, u, w])               # :174-182
    s = pow(r, e, ek.n) * beta % ek.n                                   # :86
    s1 = e * a + alpha                                                  # :87
    s2 = e * ro + gamma                                                 # :88
    return AliceProof(z, e, s, s1, s2)


def _mod_inv(x: int, m: int) -> Optional[int]:
    """`BigInt::mod_inv` -> Option (None when gcd != 1)."""
    try:
        return pow(x, -1, m)
    except ValueError:
        return None


def alice_proof_verify(pf: AliceProof, cipher: int, ek: EncryptionKey, st: DLogStatement) -> bool:
    """`AliceProof::verify` range_proofs.rs:105-156."""
    N, NN, Nt, h1, h2 = ek.n, ek.nn, st.N, st.g, st.ni
    if pf.s1 > Q**3:                                                    # :118
        return False
    z_e_inv = _mod_inv(pow(pf.z, pf.e, Nt), Nt)                         # :122
    if z_e_inv is None:
        return False
    w = pow(h1, pf.s1, Nt) * pow(h2, pf.s2, Nt) * z_e_inv % Nt          # :129-132
    gs1 = (pf.s1 * N + 1) % NN                                          # :134
    c_e_inv = _mod_inv(pow(cipher, pf.e, NN), NN)                       # :135
    if c_e_inv is None:
        return False
    u = gs1 * pow(pf.s, N, NN) * c_e_inv % NN                           # :141
    e = sha256_bigints([N, N + 1, cipher, pf.z, u, w])                  # :143-150
    return e == pf.e


# --------------------------------------------------------------------------- MtA range proof (Bob) — off-protocol API
@dataclass
class BobProof:
    t: int
    z: int
    e: int
    s: int
    s1: int
    s2: int
    t1: int
    t2: int


def bob_proof_generate(a_enc: int, mta_enc: int, b: int, beta_prim: int, ek: EncryptionKey, st: DLogStatement,
                       r: int, check: bool, alpha: int, beta: int, gamma: int, ro: int, ro_prim: int,
                       sigma: int, tau: int) -> Tuple[BobProof, Optional[Point]]:
    """`BobProof::generate` range_proofs.rs:414-487 (round1 :214-264, round2 :276-297)."""
    h1, h2, Nt = st.g, st.ni, st.N
    z = pow(h1, b, Nt) * pow(h2, ro, Nt) % Nt
    z_prim = pow(h1, alpha, Nt) * pow(h2, ro_prim, Nt) % Nt
    t = pow(h1, beta_prim, Nt) * pow(h2, sigma, Nt) % Nt
    w = pow(h1, gamma, Nt) * pow(h2, tau, Nt) % Nt
    v = pow(a_enc, alpha, ek.nn) * (gamma * ek.n + 1) * pow(beta, ek.n, ek.nn) % ek.nn
    items = [ek.n, ek.n + 1, a_enc, mta_enc, z, z_prim, t, v, w]
    u_pt = None
    if check:
        X = pt_mul(G, b)
        u_pt = pt_mul(G, alpha % Q)
        items += [X[0], X[1], u_pt[0], u_pt[1]]
    e = sha256_bigints(items)
    return BobProof(t=t, z=z, e=e, s=pow(r, e, ek.n) * beta % ek.n, s1=e * b + alpha, s2=e * ro + ro_prim,
                    t1=e * beta_prim + gamma, t2=e * sigma + tau), u_pt


def bob_proof_verify(pf: BobProof, a_enc: int, mta_out: int, ek: EncryptionKey, st: DLogStatement,
                     check: Optional[Tuple[Point, Point]] = None) -> bool:
    """`BobProof::verify` range_proofs.rs:321-412; `check` = (u, X) as in BobCheck."""
    N, NN, Nt, h1, h2 = ek.n, ek.nn, st.N, st.g, st.ni
    if pf.s1 > Q**3:
        return False
    z_e_inv = _mod_inv(pow(pf.z, pf.e, Nt), Nt)
    if z_e_inv is None:
        return False
    z_prim = pow(h1, pf.s1, Nt) * pow(h2, pf.s2, Nt) * z_e_inv % Nt
    mta_e_inv = _mod_inv(pow(mta_out, pf.e, NN), NN)
    if mta_e_inv is None:
        return False
    v = pow(a_enc, pf.s1, NN) * pow(pf.s, N, NN) * (pf.t1 * N + 1) * mta_e_inv % NN
    t_e_inv = _mod_inv(pow(pf.t, pf.e, Nt), Nt)
    if t_e_inv is None:
        return False
    w = pow(h1, pf.t1, Nt) * pow(h2, pf.t2, Nt) * t_e_inv % Nt
    items = [N, N + 1, a_enc, mta_out, pf.z, z_prim, pf.t, v, w]
    if check is not None:
        u_pt, X = check
        items += [X[0], X[1], u_pt[0], u_pt[1]]
    return sha256_bigints(items) == pf.e


def bob_proof_ext_verify(pf: BobProof, u_pt: Point, a_enc: int, mta_out: int, ek: EncryptionKey,
                         st: DLogStatement, X: Point) -> bool:
    """`BobProofExt::verify` range_proofs.rs:499-534."""
    if not bob_proof_verify(pf, a_enc, mta_out, ek, st, (u_pt, X)):
        return False
    x1 = pt_mul(G, pf.s1 % Q)
    x2 = pt_add(pt_mul(X, pf.e % Q), u_pt)
    return x1 == x2


# --------------------------------------------------------------------------- curv sigma proofs [R]
@dataclass
class DLogProof:
    pk: Point
    pk_t_rand_commitment: Point
    challenge_response: int


def dlog_prove(sk: int, nonce: int) -> DLogProof:
    """curv `DLogProof::prove` (sigma_dlog.rs [R]); call sites mta/mod.rs:147-148."""
    T = pt_mul(G, nonce)
    pk = pt_mul(G, sk)
    e = sha256_points_scalar([T, G, pk])
    return DLogProof(pk, T, (nonce - e * sk) % Q)


def dlog_verify(pf: DLogProof) -> bool:
    """curv `DLogProof::verify` [R]; call sites mta/mod.rs:170-171."""
    e = sha256_points_scalar([pf.pk_t_rand_commitment, G, pf.pk])
    return pt_add(pt_mul(G, pf.challenge_response), pt_mul(pf.pk, e)) == pf.pk_t_rand_commitment


@dataclass
class PedersenProof:
    e: int
    a1: Point
    a2: Point
    com: Point
    z1: int
    z2: int


def pedersen_prove(m: int, r: int, s1: int, s2: int) -> PedersenProof:
    """curv `PedersenProof::prove` (sigma_valid_pedersen.rs [R]); call site party_i.rs:631."""
    a1 = pt_mul(G, s1)
    a2 = pt_mul(H2, s2)
    com = pt_add(pt_mul(G, m), pt_mul(H2, r))
    e = sha256_points_scalar([G, H2, com, a1, a2])
    return PedersenProof(e, a1, a2, com, (s1 + e * m) % Q, (s2 + e * r) % Q)


def pedersen_verify(pf: PedersenProof) -> bool:
    """curv `PedersenProof::verify` [R]; call site sign/rounds.rs:371-378."""
    e = sha256_points_scalar([G, H2, pf.com, pf.a1, pf.a2])
    lhs = pt_add(pt_mul(G, pf.z1), pt_mul(H2, pf.z2))
    rhs = pt_add(pt_add(pf.a1, pf.a2), pt_mul(pf.com, e))
    return lhs == rhs


@dataclass
class HomoElGamalProof:
    T: Point
    A3: Point
    z1: int
    z2: int


def heg_prove(x: int, r: int, Gp: Point, Hp: Point, Y: Point, D: Point, E: Point, s1: int, s2: int) -> HomoElGamalProof:
    """curv `HomoELGamalProof::prove` (sigma_correct_homomorphic_elgamal_enc.rs [R]); call site party_i.rs:796."""
    A1 = pt_mul(Hp, s1)
    A2 = pt_mul(Y, s2)
    A3 = pt_mul(Gp, s2)
    T = pt_add(A1, A2)
    e = sha256_points_scalar([T, A3, Gp, Hp, Y, D, E])
    z1 = (s1 + x * e) % Q if x % Q != 0 else s1
    z2 = (s2 + r * e) % Q
    return HomoElGamalProof(T, A3, z1, z2)


def heg_verify(pf: HomoElGamalProof, Gp: Point, Hp: Point, Y: Point, D: Point, E: Point) -> bool:
    """curv `HomoELGamalProof::verify` [R]; call site party_i.rs:816."""
    e = sha256_points_scalar([pf.T, pf.A3, Gp, Hp, Y, D, E])
    ok1 = pt_add(pt_mul(Hp, pf.z1), pt_mul(Y, pf.z2)) == pt_add(pf.T, pt_mul(D, e))
    ok2 = pt_mul(Gp, pf.z2) == pt_add(pf.A3, pt_mul(E, e))
    return ok1 and ok2


# --------------------------------------------------------------------------- MtA messages
@dataclass
class MessageA:
    c: int
    range_proofs: List[AliceProof]


@dataclass
class MessageB:
    c: int
    b_proof: DLogProof
    beta_tag_proof: DLogProof


def message_a(a: int, ek: EncryptionKey, randomness: int, stmts: Sequence[DLogStatement],
              proof_rand: Sequence[Tuple[int, int, int, int]]) -> MessageA:
    """`MessageA::a_with_predefined_randomness` mta/mod.rs:62-87; proof_rand[x] = (alpha,beta,gamma,ro)."""
    c = paillier_encrypt(ek, a, randomness)
    proofs = [alice_proof_generate(a, c, ek, st, randomness, *pr) for st, pr in zip(stmts, proof_rand)]
    return MessageA(c, proofs)


def message_b(b: int, ek: EncryptionKey, m_a: MessageA, randomness: int, beta_tag: int,
              stmts: Sequence[DLogStatement], nonce_b: int, nonce_beta: int) -> Optional[Tuple[MessageB, int]]:
    """`MessageB::b_with_predefined_randomness` mta/mod.rs:111-158.  None == Err(InvalidKey)."""
    if len(m_a.range_proofs) != len(stmts):                             # :119
        return None
    if not all(alice_proof_verify(pf, m_a.c, ek, st) for pf, st in zip(m_a.range_proofs, stmts)):   # :123-131
        return None
    beta_tag_fe = beta_tag % Q                                          # :132
    c_beta_tag = paillier_encrypt(ek, beta_tag, randomness)             # :133
    b_c_a = paillier_mul(ek, m_a.c, b)                                  # :140
    c_b = paillier_add(ek, b_c_a, c_beta_tag)                           # :145
    beta = (-beta_tag_fe) % Q                                           # :146
    return MessageB(c_b, dlog_prove(b, nonce_b), dlog_prove(beta_tag_fe, nonce_beta)), beta


def verify_proofs_get_alpha(m_b: MessageB, dk: DecryptionKey, a: int) -> Optional[Tuple[int, int]]:
    """`MessageB::verify_proofs_get_alpha` mta/mod.rs:160-179.  None == Err(InvalidKey)."""
    alice_share = paillier_decrypt(dk, m_b.c)
    alpha = alice_share % Q
    g_alpha = pt_mul(G, alpha)
    ba_btag = pt_add(pt_mul(m_b.b_proof.pk, a), m_b.beta_tag_proof.pk)
    if dlog_verify(m_b.b_proof) and dlog_verify(m_b.beta_tag_proof) and ba_btag == g_alpha:
        return alpha, alice_share
    return None


# --------------------------------------------------------------------------- PDL with slack
@dataclass
class PDLwSlackProof:
    z: int
    u1: Point
    u2: int
    u3: int
    s1: int
    s2: int
    s3: int


def commitment_unknown_order(h1: int, h2: int, Nt: int, x: int, r: int) -> int:
    """zk_pdl_with_slack/mod.rs:182-199 (negative r -> invert h2; the reference unwraps/panics when
    h2 is not invertible — here a ValueError)."""
    h1_x = pow(h1, x, Nt)
    if r < 0:
        h2_r = pow(pow(h2, -1, Nt), -r, Nt)
    else:
        h2_r = pow(h2, r, Nt)
    return h1_x * h2_r % Nt


def pdl_prove(x: int, r: int, cipher: int, ek: EncryptionKey, Qp: Point, Gp: Point, h1: int, h2: int, Nt: int,
              alpha: int, beta: int, rho: int, gamma: int) -> PDLwSlackProof:
    """`PDLwSlackProof::prove` zk_pdl_with_slack/mod.rs:68-125."""
    z = commitment_unknown_order(h1, h2, Nt, x, rho)                    # :78-84
    u1 = pt_mul(Gp, alpha % Q)                                          # :85
    u2 = commitment_unknown_order(ek.n + 1, beta, ek.nn, alpha, ek.n)   # :86-92
    u3 = commitment_unknown_order(h1, h2, Nt, alpha, gamma)             # :93-99
    e = sha256_bigints([bn_from_bytes(pt_compress(Gp)), bn_from_bytes(pt_compress(Qp)), cipher, z,
                        bn_from_bytes(pt_compress(u1)), u2, u3])        # :101-109
    s1 = e * x + alpha
    s2 = commitment_unknown_order(r, beta, ek.n, e, 1)
    s3 = e * rho + gamma
    return PDLwSlackProof(z, u1, u2, u3, s1, s2, s3)


def pdl_verify(pf: PDLwSlackProof, cipher: int, ek: EncryptionKey, Qp: Point, Gp: Point, h1: int, h2: int, Nt: int) -> bool:
    """`PDLwSlackProof::verify` zk_pdl_with_slack/mod.rs:127-179."""
    e = sha256_bigints([bn_from_bytes(pt_compress(Gp)), bn_from_bytes(pt_compress(Qp)), cipher, pf.z,
                        bn_from_bytes(pt_compress(pf.u1)), pf.u2, pf.u3])
    g_s1 = pt_mul(Gp, pf.s1 % Q)
    y_minus_e = pt_mul(Qp, (Q - e) % Q)
    u1_test = pt_add(g_s1, y_minus_e)
    try:
        u2_tmp = commitment_unknown_order(ek.n + 1, pf.s2, ek.nn, pf.s1, ek.n)
        u2_test = commitment_unknown_order(u2_tmp, cipher, ek.nn, 1, -e)
        u3_tmp = commitment_unknown_order(h1, h2, Nt, pf.s1, pf.s3)
        u3_test = commitment_unknown_order(u3_tmp, pf.z, Nt, 1, -e)
    except ValueError:          # reference: .unwrap() panic on a non-invertible base (:192)
        return False
    return pf.u1 == u1_test and pf.u2 == u2_test and pf.u3 == u3_test


# --------------------------------------------------------------------------- GG20 party-level helpers
def lagrange_at_zero(index: int, s: Sequence[int]) -> int:
    """curv `VerifiableSS::map_share_to_new_params(params, index, s)` [R]: Lagrange basis at 0 over
    the points x_m = s[m] + 1.  Call sites party_i.rs:536-540,553-557."""
    j = list(s).index(index)
    xs = [(x + 1) % Q for x in s]
    num, den = 1, 1
    for m, xm in enumerate(xs):
        if m != j:
            num = num * xm % Q
            den = den * ((xm - xs[j]) % Q) % Q
    return num * pow(den, -1, Q) % Q


@dataclass
class LocalKey:
    """The fields of `LocalKey<Secp256k1>` (keygen/rounds.rs:310-322) the offline stage reads."""
    i: int                                  # 1-based keygen index
    t: int
    n: int
    x_i: int                                # keys_linear.x_i
    dk: DecryptionKey                       # paillier_dk
    pk_vec: List[Point]                     # X_j = x_j * G
    paillier_key_vec: List[EncryptionKey]
    h1_h2_n_tilde_vec: List[DLogStatement]
    y_sum_s: Point


@dataclass
class UnitRandomness:
    """Every value one party's OfflineStage samples, in order of use (t=1: one peer)."""
    gamma_i: int = 0
    k_i: int = 0
    blind: int = 0
    r_k: int = 0
    alice: List[Tuple[int, int, int, int]] = field(default_factory=list)   # per statement: alpha,beta,gamma,ro
    beta_tag_gamma: int = 0
    r_gamma: int = 0
    nonce_gamma_b: int = 0
    nonce_gamma_beta: int = 0
    beta_tag_w: int = 0
    r_w: int = 0
    nonce_w_b: int = 0
    nonce_w_beta: int = 0
    l: int = 0
    ped_s1: int = 0
    ped_s2: int = 0
    pdl: Tuple[int, int, int, int] = (0, 0, 0, 0)                          # alpha,beta,rho,gamma
    heg_s1: int = 0
    heg_s2: int = 0


@dataclass
class UnitResult:
    """What `CompletedOfflineStage` holds (sign/rounds.rs:647-654) plus the emitted messages."""
    status: int
    R: Point = None
    sigma_i: int = 0
    k_i: int = 0
    t_vec: List[Point] = field(default_factory=list)
    transcript: bytes = b""


ST_OK, ST_INVALID_KEY, ST_PDL, ST_PHASE5, ST_PHASE6, ST_PROOF, ST_COMMIT = 0, 2, 6, 7, 8, 10, 11


def _enc_point(p: Point) -> bytes:
    return pt_compress(p)


def _enc_int(x: int, width: int) -> bytes:
    return x.to_bytes(width, "big")


def offline_session(keys: Sequence[LocalKey], s_l: Sequence[int], rnd: Sequence[UnitRandomness]) -> List[UnitResult]:
    """One two-signer GG20 offline session: both parties' `OfflineStage` Round0..Round6
    (sign/rounds.rs:68-636) run in lock-step with their messages exchanged in memory.
    keys[p], rnd[p] belong to signer position p (0/1); s_l[p] is its 1-based keygen index.
    The transcript of a unit is the canonical concatenation of every message it EMITS."""
    ttag = len(s_l)
    assert ttag == 2, "this restatement covers the t=1 (two signers) configuration"
    l_s = [x - 1 for x in s_l]
    tr: List[List[bytes]] = [[], []]
    # ---- Round 0 (rounds.rs:68-104)
    w, gamma, k, g_gamma, com, m_a = [], [], [], [], [], []
    for p in range(2):
        lk, r = keys[p], rnd[p]
        li = lagrange_at_zero(l_s[p], l_s)
        w_i = li * lk.x_i % Q
        w.append(w_i); gamma.append(r.gamma_i % Q); k.append(r.k_i % Q)
        gg = pt_mul(G, gamma[p]); g_gamma.append(gg)
        com.append(hash_commitment(bn_from_bytes(pt_compress(gg)), r.blind))          # party_i.rs:573-580
        ek = lk.paillier_key_vec[lk.i - 1]
        m_a.append(message_a(k[p], ek, r.r_k, lk.h1_h2_n_tilde_vec, r.alice))
        tr[p].append(_enc_int(m_a[p].c, 512) + b"".join(
            _enc_int(pf.z, 256) + _enc_int(pf.e, 32) + _enc_int(pf.s, 256) + _enc_int(pf.s1, 128) + _enc_int(pf.s2, 384)
            for pf in m_a[p].range_proofs) + _enc_int(com[p], 32))
    # ---- Round 1 (rounds.rs:122-206)
    m_b_gamma, m_b_w, beta_v, ni_v = [None, None], [None, None], [0, 0], [0, 0]
    res = [UnitResult(ST_OK), UnitResult(ST_OK)]
    for p in range(2):
        o = 1 - p
        lk, r = keys[p], rnd[p]
        ek_o = lk.paillier_key_vec[l_s[o]]
        rb = message_b(gamma[p], ek_o, m_a[o], r.r_gamma, r.beta_tag_gamma, lk.h1_h2_n_tilde_vec, r.nonce_gamma_b, r.nonce_gamma_beta)
        rw = message_b(w[p], ek_o, m_a[o], r.r_w, r.beta_tag_w, lk.h1_h2_n_tilde_vec, r.nonce_w_b, r.nonce_w_beta)
        if rb is None or rw is None:
            res[p].status = ST_INVALID_KEY
            continue
        m_b_gamma[p], beta_v[p] = rb
        m_b_w[p], ni_v[p] = rw
        for mb in (m_b_gamma[p], m_b_w[p]):
            tr[p].append(_enc_int(mb.c, 512) + b"".join(
                _enc_point(d.pk) + _enc_point(d.pk_t_rand_commitment) + _enc_int(d.challenge_response, 32)
                for d in (mb.b_proof, mb.beta_tag_proof)))
    if any(r_.status for r_ in res):
        return _finish(res, tr)
    # ---- Round 2 (rounds.rs:234-317)
    delta, sigma, T, l_v, t_proof = [0, 0], [0, 0], [None, None], [0, 0], [None, None]
    for p in range(2):
        o = 1 - p
        lk, r = keys[p], rnd[p]
        g_w_vec = [pt_mul(lk.pk_vec[l_s[x]], lagrange_at_zero(l_s[x], l_s)) for x in range(2)]   # party_i.rs:527-544
        ra = verify_proofs_get_alpha(m_b_gamma[o], lk.dk, k[p])
        rm = verify_proofs_get_alpha(m_b_w[o], lk.dk, k[p])
        if ra is None or rm is None or m_b_w[o].b_proof.pk != g_w_vec[o]:                          # :281 (assert_eq!)
            res[p].status = ST_INVALID_KEY
            continue
        delta[p] = (k[p] * gamma[p] + ra[0] + beta_v[p]) % Q                                         # party_i.rs:591-604
        sigma[p] = (k[p] * w[p] + rm[0] + ni_v[p]) % Q                                               # :606-618
        l_v[p] = r.l % Q
        T[p] = pt_add(pt_mul(G, sigma[p]), pt_mul(H2, l_v[p]))                                       # :620-634
        t_proof[p] = pedersen_prove(sigma[p], l_v[p], r.ped_s1 % Q, r.ped_s2 % Q)
        pf = t_proof[p]
        tr[p].append(_enc_int(delta[p], 32) + _enc_point(T[p]) + _enc_int(pf.e, 32) + _enc_point(pf.a1) +
                     _enc_point(pf.a2) + _enc_point(pf.com) + _enc_int(pf.z1, 32) + _enc_int(pf.z2, 32))
    if any(r_.status for r_ in res):
        return _finish(res, tr)
    # ---- Round 3 (rounds.rs:347-402)
    delta_inv = pow((delta[0] + delta[1]) % Q, -1, Q)                                                # party_i.rs:635-640
    for p in range(2):
        if any(T[x] != t_proof[x].com for x in range(2)) or not all(pedersen_verify(t_proof[x]) for x in range(2)):
            res[p].status = ST_PROOF
            continue
        tr[p].append(_enc_int(rnd[p].blind, 32) + _enc_point(g_gamma[p]))
    if any(r_.status for r_ in res):
        return _finish(res, tr)
    # ---- Round 4 (rounds.rs:431-498)
    R, R_dash, pdl = [None, None], [None, None], [None, None]
    for p in range(2):
        o = 1 - p
        lk, r = keys[p], rnd[p]
        ok = (m_b_gamma[o].b_proof.pk == g_gamma[o] and
              hash_commitment(bn_from_bytes(pt_compress(g_gamma[o])), rnd[o].blind) == com[o])     # party_i.rs:650-674
        if not ok:
            res[p].status = ST_COMMIT
            continue
        R[p] = pt_mul(pt_add(g_gamma[0], g_gamma[1]), delta_inv)                                     # :684-686
        R_dash[p] = pt_mul(R[p], k[p])                                                               # rounds.rs:452
        st = lk.h1_h2_n_tilde_vec[l_s[o]]
        pdl[p] = pdl_prove(k[p], r.r_k, m_a[p].c, lk.paillier_key_vec[l_s[p]], R_dash[p], R[p], st.g, st.ni, st.N, *r.pdl)
        pf = pdl[p]
        tr[p].append(_enc_point(R_dash[p]) + _enc_int(pf.z, 256) + _enc_point(pf.u1) + _enc_int(pf.u2, 512) +
                     _enc_int(pf.u3, 256) + _enc_int(pf.s1, 128) + _enc_int(pf.s2, 256) + _enc_int(pf.s3, 384))
    if any(r_.status for r_ in res):
        return _finish(res, tr)
    # ---- Round 5 (rounds.rs:525-592)
    S, heg = [None, None], [None, None]
    for p in range(2):
        lk, r = keys[p], rnd[p]
        ok = True
        for x in range(2):          # every signer's proof list, own included (party_i.rs:719-766)
            y = 1 - x               # the single proof of signer x was made against the statement of its peer
            st = lk.h1_h2_n_tilde_vec[l_s[y]]
            ok = ok and pdl_verify(pdl[x], m_a[x].c, lk.paillier_key_vec[l_s[x]], R_dash[x], R[p], st.g, st.ni, st.N)
        if not ok:
            res[p].status = ST_PDL
            continue
        ssum = pt_add(pt_add(G, R_dash[0]), R_dash[1])                                               # party_i.rs:768-776
        if pt_sub(ssum, G) != G:
            res[p].status = ST_PHASE5
            continue
        S[p] = pt_mul(R[p], sigma[p])                                                                # :784
        heg[p] = heg_prove(l_v[p], sigma[p], R[p], H2, G, T[p], S[p], r.heg_s1 % Q, r.heg_s2 % Q)
        pf = heg[p]
        tr[p].append(_enc_point(S[p]) + _enc_point(pf.T) + _enc_point(pf.A3) + _enc_int(pf.z1, 32) + _enc_int(pf.z2, 32))
    if any(r_.status for r_ in res):
        return _finish(res, tr)
    # ---- Round 6 (rounds.rs:612-636)
    for p in range(2):
        lk = keys[p]
        if not all(heg_verify(heg[x], R[p], H2, G, T[x], S[x]) for x in range(2)):                   # party_i.rs:801-833
            res[p].status = ST_PHASE6
            continue
        ssum = pt_sub(pt_add(pt_add(G, S[0]), S[1]), G)                                              # :835-848
        if ssum != lk.y_sum_s:
            res[p].status = ST_PHASE6
            continue
        res[p].R, res[p].sigma_i, res[p].k_i, res[p].t_vec = R[p], sigma[p], k[p], [T[0], T[1]]
    return _finish(res, tr)


def _finish(res: List[UnitResult], tr: List[List[bytes]]) -> List[UnitResult]:
    for p in range(2):
        res[p].transcript = hashlib.sha256(b"".join(tr[p])).digest()
    return res


# --------------------------------------------------------------------------- online step (end-to-end validity only)
def local_sig(k_i: int, message: int, R: Point, sigma_i: int) -> int:
    """`LocalSignature::phase7_local_sig` party_i.rs:850-871: s_i = m*k_i + r*sigma_i."""
    r = R[0] % Q
    return (message % Q * k_i + r * sigma_i) % Q


def output_signature(R: Point, s_parts: Sequence[int]) -> Tuple[int, int, int]:
    """`LocalSignature::output_signature` party_i.rs:873-910 (low-s normalisation, recid)."""
    s = sum(s_parts) % Q
    r = R[0] % Q
    recid = (R[1] % Q) & 1
    if s > Q - s:
        s = Q - s
        recid ^= 1
    return r, s, recid


def ecdsa_verify(r: int, s: int, y: Point, message: int) -> bool:
    """`verify` party_i.rs:913-936."""
    b = pow(s, -1, Q)
    u1 = message % Q * b % Q
    u2 = r * b % Q
    pt = pt_add(pt_mul(G, u1), pt_mul(y, u2))
    return pt is not None and r == pt[0] % Q


# --------------------------------------------------------------------------- identifiable abort helpers (gg_2020/blame.rs)
def paillier_open(dk: DecryptionKey, c: int) -> Tuple[int, int]:
    """kzen-paillier `Paillier::open(dk, c)` [R] (call site gg_2020/blame.rs:252-256): the plaintext and the randomness r with
    c = (1 + m n) r^n mod n^2.  r is the unique n-th root of c modulo n: r = (c mod n)^(n^-1 mod phi) mod n."""
    n = dk.p * dk.q
    phi = (dk.p - 1) * (dk.q - 1)
    return paillier_decrypt(dk, c), pow(c % n, pow(n, -1, phi), n)


@dataclass
class ECDDHProof:
    a1: Point
    a2: Point
    z: int


def ecddh_prove(x: int, g1: Point, h1: Point, g2: Point, h2: Point, s: int) -> ECDDHProof:
    """curv `ECDDHProof::prove(&w, &delta)` (sigma_ec_ddh.rs [R]; call site blame.rs:258-271): a1 = s g1, a2 = s g2,
    e = H(g1, h1, g2, h2, a1, a2), z = s + e x"""
    a1, a2 = pt_mul(g1, s), pt_mul(g2, s)
    e = sha256_points_scalar([g1, h1, g2, h2, a1, a2])
    return ECDDHProof(a1, a2, (s + e * x) % Q)


def ecddh_verify(pf: ECDDHProof, g1: Point, h1: Point, g2: Point, h2: Point) -> bool:
    """curv `ECDDHProof::verify(&delta)` [R] (call site blame.rs:410-413): z g1 == a1 + e h1 and z g2 == a2 + e h2"""
    e = sha256_points_scalar([g1, h1, g2, h2, pf.a1, pf.a2])
    return pt_mul(g1, pf.z) == pt_add(pf.a1, pt_mul(h1, e)) and pt_mul(g2, pf.z) == pt_add(pf.a2, pt_mul(h2, e))
