// Per-element kernels of the Lindell-2017 / zk_pdl entry points (lindell17.cu): secp256k1 and SHA-256 steps around the jobs modulo
// N^2, one thread per element.  Device code only (no runtime calls), so that tests/host_harness can compile it for the CPU.
#pragma once
#include "gg20_glue.cuh"

namespace tecdsa {
namespace l17 {

__device__ __forceinline__ bool good_point(const Affine& P) { return !P.inf && on_curve(P); }
// a > b for 8-limb values
__device__ __forceinline__ bool u256_gt(const U256& a, const U256& b) {
    for (int i = 7; i >= 0; i--) { if (a.v[i] != b.v[i]) return a.v[i] > b.v[i]; }
    return false;
}
// 1 + m * N as a 128-limb value (m < N, so no reduction modulo N^2 is needed)
__device__ __forceinline__ void lin_of(uint32_t* out128, const uint32_t* m, int m_limbs, const uint32_t* N) {
    uint32_t one = 1;
    st::mul_add(out128, 128, m, m_limbs, N, 64, &one, 1);
}
// `HashCommitment::create_commitment_with_user_defined_randomness(&m, &blind)`, m given as limbs
__device__ __forceinline__ void hash_commit_bigint(uint32_t* out8, const uint32_t* m, int m_limbs, const uint32_t* blind8) {
    Sha256 h; h.init();
    h.put_bigint(m, m_limbs);
    h.put_bigint(blind8, 8);
    h.finish(out8);
}

// ---- party two: PartialSig::compute (party_two.rs:390-424), the scalar part -------------------------------------------
// r = k2 * R1, rx = r.x mod q, k2_inv, partial_sig = rho q + (k2_inv m mod q)  [24 limbs], v = k2_inv (rx x2) mod q; lin = 1 + partial_sig N
__global__ void k_l17_p2_pre(const uint32_t* n_tab, const uint32_t* key_idx, const uint32_t* x2, const uint32_t* k2, const uint32_t* R1,
                             const uint32_t* msg, const uint32_t* rho16, uint32_t* v8, uint32_t* lin128, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U256 k = load_scalar(k2 + (size_t)i * 8);
    const Affine P = affine_load(R1 + (size_t)i * 16);
    uint8_t st = TECDSA_ST_OK;
    if (u256_is_zero(k)) st = TECDSA_ST_NOT_INVERTIBLE;            // `mod_inv(..).unwrap()` panics in the reference
    else if (!good_point(P)) st = TECDSA_ST_INVALID_KEY;
    uint32_t ps[24];
    U256 v = u256_zero();
    for (int j = 0; j < 24; j++) ps[j] = 0;
    if (st == TECDSA_ST_OK) {
        const Affine r = jac_to_affine(jac_mul(jac_from_affine(P), k));
        const U256 rx = sc_reduce_once(r.x, 0);
        const U256 kinv = sc_inv(k);
        const U256 t = sc_mul(kinv, sc_from_limbs(msg + (size_t)i * 8, 8));
        st::mul_add(ps, 24, rho16 + (size_t)i * 16, 16, Q_LIMBS, 8, t.v, 8);
        v = sc_mul(kinv, sc_mul(rx, load_scalar(x2 + (size_t)i * 8)));
    }
    u256_store(v8 + (size_t)i * 8, v);
    lin_of(lin128 + (size_t)i * 128, ps, 24, n_tab + (size_t)(key_idx ? key_idx[i] : i) * 64);
    status[i] = st;
}

// ---- party one: Signature::compute_with_recid (party_one.rs:519-564) after the decrypt ----------------------------------
__global__ void k_l17_p1_post(const uint32_t* s_tag64, const uint32_t* k1, const uint32_t* R2, uint32_t* sig_r, uint32_t* sig_s, uint8_t* recid,
                              uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U256 k = load_scalar(k1 + (size_t)i * 8);
    const Affine P = affine_load(R2 + (size_t)i * 16);
    uint8_t st = TECDSA_ST_OK, rid = 0;
    U256 r = u256_zero(), s = u256_zero();
    if (u256_is_zero(k)) st = TECDSA_ST_NOT_INVERTIBLE;
    else if (!good_point(P)) st = TECDSA_ST_INVALID_KEY;
    else {
        const Affine Rp = jac_to_affine(jac_mul(jac_from_affine(P), k));
        r = sc_reduce_once(Rp.x, 0);
        rid = (uint8_t)(sc_reduce_once(Rp.y, 0).v[0] & 1u);
        const U256 s2 = sc_mul(sc_from_limbs(s_tag64 + (size_t)i * 64, 64), sc_inv(k));
        const U256 neg = sc_neg(s2);
        if (u256_gt(s2, neg)) { s = neg; rid ^= 1; } else s = s2;      // s = min(s'', q - s''), recid flips with it
    }
    u256_store(sig_r + (size_t)i * 8, r); u256_store(sig_s + (size_t)i * 8, s);
    recid[i] = rid;
    status[i] = st;
}

// ---- party_one::verify (party_one.rs:567-592) ----------------------------------------------------------------------------
__global__ void k_l17_verify(const uint32_t* sig_r, const uint32_t* sig_s, const uint32_t* pubkey, const uint32_t* msg, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U256 r_raw = u256_load(sig_r + (size_t)i * 8), s_raw = u256_load(sig_s + (size_t)i * 8);
    const U256 s = sc_reduce_once(s_raw, 0), r = sc_reduce_once(r_raw, 0);
    const Affine Y = affine_load(pubkey + (size_t)i * 16);
    bool ok = good_point(Y) && !u256_is_zero(s);
    if (ok) {
        const U256 b = sc_inv(s);
        const Affine P = lin_GP(sc_mul(sc_from_limbs(msg + (size_t)i * 8, 8), b), Y, sc_mul(r, b));
        // the reference compares the BYTES of signature.r with those of the unreduced x coordinate, and wants s < q - s
        // (on the signature's s as given: a non-canonical s >= q fails the comparison like the BigInt arithmetic does)
        ok = !P.inf && u256_eq(P.x, r_raw) && u256_eq(s, s_raw) && u256_gt(sc_neg(s), s);
    }
    status[i] = ok ? TECDSA_ST_OK : TECDSA_ST_INVALID_SIG;
}

// ---- ephemeral key exchange (party_one.rs:403-433, party_two.rs:314-371) -----------------------------------------------
// public_share = k G, c = k H, ECDDHProof over (G, kG, H, kH) with nonce s: a1 = s G, a2 = s H, z = s + e k; party two adds
// pk_commitment = commit(compressed public_share; pk_blind) and zk_pok_commitment = commit(H(a1, a2); zk_blind)
__global__ void k_l17_eph_create(const uint32_t* k8, const uint32_t* nonce8, const uint32_t* pk_blind, const uint32_t* zk_blind, uint32_t* pub16,
                                 uint32_t* c16, uint32_t* proof40, uint32_t* pk_com8, uint32_t* zk_com8, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U256 k = load_scalar(k8 + (size_t)i * 8), s = load_scalar(nonce8 + (size_t)i * 8);
    Affine pts[6];
    pts[0] = affine_G(); pts[2] = affine_H();
    jac_to_affine2(pts[1], pts[3], jac_mul_fixed(0, k), jac_mul_fixed(1, k));
    jac_to_affine2(pts[4], pts[5], jac_mul_fixed(0, s), jac_mul_fixed(1, s));
    const U256 e = hash_points_scalar(pts, 6);
    affine_store(pub16 + (size_t)i * 16, pts[1]); affine_store(c16 + (size_t)i * 16, pts[3]);
    uint32_t* o = proof40 + (size_t)i * 40;
    affine_store(o, pts[4]); affine_store(o + 16, pts[5]); u256_store(o + 32, sc_add(s, sc_mul(e, k)));
    if (pk_com8) {
        hash_commit_point(pk_com8 + (size_t)i * 8, pts[1], pk_blind + (size_t)i * 8);
        commit_points(zk_com8 + (size_t)i * 8, pts + 4, 2, zk_blind + (size_t)i * 8);
    }
}
// party_one::EphKeyGenSecondMsg::verify_commitments_and_dlog_proof (party_one.rs:436-482) when the commitments are given,
// party_two::EphKeyGenSecondMsg::verify_and_decommit (party_two.rs:374-387) when they are not
__global__ void k_l17_eph_verify(const uint32_t* pub16, const uint32_t* c16, const uint32_t* proof40, const uint32_t* pk_blind, const uint32_t* zk_blind,
                                 const uint32_t* pk_com8, const uint32_t* zk_com8, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t* p = proof40 + (size_t)i * 40;
    Affine pts[6];
    pts[0] = affine_G(); pts[1] = affine_load(pub16 + (size_t)i * 16); pts[2] = affine_H(); pts[3] = affine_load(c16 + (size_t)i * 16);
    pts[4] = affine_load(p); pts[5] = affine_load(p + 16);
    bool pts_ok = good_point(pts[1]) && good_point(pts[3]) && good_point(pts[4]) && good_point(pts[5]);
    uint8_t st = TECDSA_ST_OK;
    if (pk_com8 && pts_ok) {
        uint32_t t[8];
        hash_commit_point(t, pts[1], pk_blind + (size_t)i * 8);
        bool same = true;
        for (int j = 0; j < 8; j++) same = same && t[j] == pk_com8[(size_t)i * 8 + j];
        commit_points(t, pts + 4, 2, zk_blind + (size_t)i * 8);
        for (int j = 0; j < 8; j++) same = same && t[j] == zk_com8[(size_t)i * 8 + j];
        if (!same) st = TECDSA_ST_COMMITMENT;
    }
    if (st == TECDSA_ST_OK) {
        bool ok = pts_ok;
        if (ok) {
            const U256 z = load_scalar(p + 32);
            const U256 e = hash_points_scalar(pts, 6);
            const bool ok1 = jac_eq(jac_mul_fixed(0, z), jac_madd(jac_mul(jac_from_affine(pts[1]), e), pts[4]));
            const bool ok2 = jac_eq(jac_mul_fixed(1, z), jac_madd(jac_mul(jac_from_affine(pts[3]), e), pts[5]));
            ok = ok1 && ok2;
        }
        if (!ok) st = TECDSA_ST_PROOF;
    }
    status[i] = st;
}

// ---- interactive PDL proof (utilities/zk_pdl/mod.rs) -------------------------------------------------------------------
// Verifier::message1 (:111-148), scalar part: lin = 1 + b N, c'' = commit(a + (b << bit_length(a)); blindness), Q' = a Q + b G
__global__ void k_zkpdl_v1_pre(const uint32_t* n_tab, const uint32_t* key_idx, const uint32_t* Qpt, const uint32_t* a8, const uint32_t* b16,
                               const uint32_t* blind8, uint32_t* lin128, uint32_t* ctt8, uint32_t* qtag16, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t *a = a8 + (size_t)i * 8, *b = b16 + (size_t)i * 16;
    lin_of(lin128 + (size_t)i * 128, b, 16, n_tab + (size_t)(key_idx ? key_idx[i] : i) * 64);
    // ab_concat = a + (b << bit_length(a))
    int bl = 0;
    for (int j = 7; j >= 0; j--) { if (a[j]) { bl = 32 * j + 32 - __clz(a[j]); break; } }
    uint32_t cat[25];
    for (int j = 0; j < 25; j++) cat[j] = j < 8 ? a[j] : 0;
    const int ws = bl >> 5, bs = bl & 31;
    for (int j = 0; j < 16; j++) {
        const uint64_t w = (uint64_t)b[j] << bs;
        cat[j + ws] |= (uint32_t)w;
        if (j + ws + 1 < 25) cat[j + ws + 1] |= (uint32_t)(w >> 32);
    }
    hash_commit_bigint(ctt8 + (size_t)i * 8, cat, 25, blind8 + (size_t)i * 8);
    const Affine Qp = affine_load(Qpt + (size_t)i * 16);
    const U256 as = load_scalar(a), bs256 = sc_from_limbs(b, 16);
    uint8_t st = TECDSA_ST_OK;
    Affine qt = affine_inf();
    if (!good_point(Qp)) st = TECDSA_ST_INVALID_KEY;
    else {
        qt = jac_to_affine(jac_add(jac_mul(jac_from_affine(Qp), as), jac_mul_fixed(0, bs256)));
        if (qt.inf) st = TECDSA_ST_INVALID_KEY;                   // a Q + b G = identity: not representable in the messages
    }
    affine_store(qtag16 + (size_t)i * 16, qt);
    status[i] = st;
}
// Prover::message1 (:191-215) after the decrypt: q_hat = (alpha mod q) G, c_hat = commit(compressed q_hat; blindness)
__global__ void k_zkpdl_p1_post(const uint32_t* alpha64, const uint32_t* blind8, uint32_t* chat8, uint32_t* qhat16, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U256 al = sc_from_limbs(alpha64 + (size_t)i * 64, 64);
    uint8_t st = TECDSA_ST_OK;
    Affine qh = affine_inf();
    if (u256_is_zero(al)) st = TECDSA_ST_INVALID_KEY;             // identity point: no encoding to commit to
    else qh = mul_G(al);
    affine_store(qhat16 + (size_t)i * 16, qh);
    if (st == TECDSA_ST_OK) hash_commit_point(chat8 + (size_t)i * 8, qh, blind8 + (size_t)i * 8);
    else for (int j = 0; j < 8; j++) chat8[(size_t)i * 8 + j] = 0;
    status[i] = st;
}
// Prover::message2 (:217-243): a x1 + b == alpha over the integers, and the verifier's commitment reopens
__global__ void k_zkpdl_p2(const uint32_t* x1, const uint32_t* alpha64, const uint32_t* ctt8, const uint32_t* a8, const uint32_t* b16,
                           const uint32_t* blind8, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t *a = a8 + (size_t)i * 8, *b = b16 + (size_t)i * 16, *al = alpha64 + (size_t)i * 64;
    uint32_t t[17];
    st::mul_add(t, 17, a, 8, x1 + (size_t)i * 8, 8, b, 16);
    bool ok = true;
    for (int j = 0; j < 64; j++) ok = ok && al[j] == (j < 17 ? t[j] : 0u);
    int bl = 0;
    for (int j = 7; j >= 0; j--) { if (a[j]) { bl = 32 * j + 32 - __clz(a[j]); break; } }
    uint32_t cat[25], d[8];
    for (int j = 0; j < 25; j++) cat[j] = j < 8 ? a[j] : 0;
    const int ws = bl >> 5, bs = bl & 31;
    for (int j = 0; j < 16; j++) {
        const uint64_t w = (uint64_t)b[j] << bs;
        cat[j + ws] |= (uint32_t)w;
        if (j + ws + 1 < 25) cat[j + ws + 1] |= (uint32_t)(w >> 32);
    }
    hash_commit_bigint(d, cat, 25, blind8 + (size_t)i * 8);
    for (int j = 0; j < 8; j++) ok = ok && d[j] == ctt8[(size_t)i * 8 + j];
    status[i] = ok ? TECDSA_ST_OK : TECDSA_ST_PDL_VERIFY;
}
// Verifier::finalize (:170-187)
__global__ void k_zkpdl_finalize(const uint32_t* chat8, const uint32_t* qhat16, const uint32_t* blind8, const uint32_t* qtag16, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Affine qh = affine_load(qhat16 + (size_t)i * 16), qt = affine_load(qtag16 + (size_t)i * 16);
    bool ok = good_point(qh) && good_point(qt) && u256_eq(qh.x, qt.x) && u256_eq(qh.y, qt.y);
    if (ok) {
        uint32_t d[8];
        hash_commit_point(d, qh, blind8 + (size_t)i * 8);
        for (int j = 0; j < 8; j++) ok = ok && d[j] == chat8[(size_t)i * 8 + j];
    }
    status[i] = ok ? TECDSA_ST_OK : TECDSA_ST_PDL_VERIFY;
}

__global__ void k_l17_decrypt_finish(Arena A, uint32_t* out, const uint32_t* dp, const uint32_t* dq, const uint32_t* rows, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    decrypt_finish(out + (size_t)i * 64, A, rows[i], dp + (size_t)i * 64, dq + (size_t)i * 64);
}

}  // namespace l17
}  // namespace tecdsa
