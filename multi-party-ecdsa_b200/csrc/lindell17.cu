// Lindell-2017 two-party ECDSA and the interactive PDL proof as batched drivers over the same primitives (SURVEY.md section 8(f)
// rank 4).  Reference (read-only): /root/reference/src/protocols/two_party_ecdsa/lindell_2017/{party_one,party_two}.rs,
// /root/reference/src/utilities/zk_pdl/mod.rs.  The heavy steps are the jobs modulo N^2 of nadic.cuh (Paillier encrypt, scalar
// multiply, add: ONE job per element computes c_key^v * r^N * (1 + m N)) and the p-adic CRT decrypt; the kernels in this file are
// the per-element secp256k1 / SHA-256 steps around them, one thread per element.  Key generation of this protocol is a
// composition of existing entry points (DLogProof, hash commitments, NiCorrectKeyProof, PDL-with-slack, CompositeDLogProof) and
// lives in the host layer (multi-party-ecdsa_b200/lindell17.py).  Oracle: oracle/lindell17_oracle.py.
#include "stage.cuh"
#include "lindell17_kernels.cuh"

using namespace tecdsa;

int tecdsa_internal_fb_points_set_lindell17(const uint32_t* table) {
    CK(cudaMemcpyToSymbol(secp::g_fb_points, &table, sizeof(table)));
    return 0;
}

using namespace tecdsa::l17;

namespace {

// `Paillier::decrypt` (CRT, kzen-paillier [R]) of 128-limb ciphertexts under key rows of an uploaded key set -> 64-limb plaintexts:
// c^(p-1) mod p^2 and c^(q-1) mod q^2 as p-adic jobs, then the L-function / recombination per element
int decrypt_dev(tecdsa_ctx* c, const tecdsa_keyset* ks, const uint32_t* rows, const uint32_t* ct, uint32_t* dp, uint32_t* dq, uint32_t* m64, int count) {
    Launches L;
    const Operand cw = arr(ct, 128);
    add_nn(L.e128, count, tab(ks->tab[KT_P], rows, 32), tab(ks->nadic_p, rows, NADIC_ROW * 32), 1, cw, tab(ks->tab[KT_PM1], rows, 32), 32, NONE, NONE, 0, 0, NONE, NONE, dp, 64);
    add_nn(L.e128, count, tab(ks->tab[KT_Q], rows, 32), tab(ks->nadic_q, rows, NADIC_ROW * 32), 1, cw, tab(ks->tab[KT_QM1], rows, 32), 32, NONE, NONE, 0, 0, NONE, NONE, dq, 64);
    int rc = run_nn(c, L.e128, 32);
    if (rc) return rc;
    k_l17_decrypt_finish<<<grid_for((size_t)count), 64, 0, c->stream>>>(key_arena(ks), m64, dp, dq, rows, count);
    c->count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : tecdsa_fail(TECDSA_E_CUDA, "paillier decrypt", e);
}

}  // namespace

extern "C" int tecdsa_l17_partial_sig_batch(tecdsa_ctx* c, const uint32_t* n, const uint32_t* key_idx, size_t n_keys, const uint32_t* c_key,
                                            const uint32_t* x2, const uint32_t* k2, const uint32_t* eph_other_public, const uint32_t* message,
                                            const uint32_t* rho, const uint32_t* randomness, uint32_t* c3, uint8_t* status, size_t count, int mem) {
    if (!c || !n || !c_key || !x2 || !k2 || !eph_other_public || !message || !rho || !randomness || !c3 || !status)
        return tecdsa_fail(TECDSA_E_ARG, "l17_partial_sig: null argument");
    if (count == 0) return 0;
    CK(cudaSetDevice(c->device));
    const size_t nk = key_idx ? n_keys : count;
    Stage S(c, mem);
    const uint32_t *dn = S.in(n, nk * 64), *di = S.in(key_idx, count), *dck = S.in(c_key, count * 128), *dx2 = S.in(x2, count * 8), *dk2 = S.in(k2, count * 8),
                   *dR = S.in(eph_other_public, count * 16), *dm = S.in(message, count * 8), *drho = S.in(rho, count * 16), *dr = S.in(randomness, count * 64);
    uint32_t* dc3 = S.out(c3, count * 128);
    uint8_t* dst = S.out(status, count);
    uint32_t *nd = S.tmp<uint32_t>(nk * NADIC_ROW * 64), *lin = S.tmp<uint32_t>(count * 128), *v = S.tmp<uint32_t>(count * 8);
    if (S.err) return S.finish();
    RUN(c->nadic_setup(dn, nd, (int)nk, 64));
    k_l17_p2_pre<<<grid_for(count), 64, 0, c->stream>>>(dn, di, dx2, dk2, dR, dm, drho, v, lin, dst, (int)count);
    KCHECK();
    Launches L;
    const Operand N = di ? tab(dn, di, 64) : arr(dn, 64), ND = di ? tab(nd, di, NADIC_ROW * 64) : arr(nd, NADIC_ROW * 64);
    // c3 = c_key^v * r^N * (1 + partial_sig N) mod N^2: Paillier::mul, ::encrypt and ::add of party_two.rs:407-423 as one job
    add_nn(L.e128, (int)count, N, ND, 2, arr(dr, 64), N, 64, arr(dck, 128), arr(v, 8), 8, 1, arr(lin, 128), NONE, dc3, 128);
    RUN(run_nn(c, L.e128));
    return S.finish();
}

extern "C" int tecdsa_l17_sign_batch(tecdsa_ctx* c, const tecdsa_keyset* ks, const uint32_t* key_row, const uint32_t* c3, const uint32_t* k1,
                                     const uint32_t* eph_other_public, uint32_t* sig_r, uint32_t* sig_s, uint8_t* recid, uint8_t* status,
                                     size_t count, int mem) {
    if (!c || !ks || !key_row || !c3 || !k1 || !eph_other_public || !sig_r || !sig_s || !recid || !status)
        return tecdsa_fail(TECDSA_E_ARG, "l17_sign: null argument");
    if (count == 0) return 0;
    CK(cudaSetDevice(c->device));
    Stage S(c, mem);
    const uint32_t *drow = S.in(key_row, count), *dct = S.in(c3, count * 128), *dk = S.in(k1, count * 8), *dR = S.in(eph_other_public, count * 16);
    uint32_t *dr = S.out(sig_r, count * 8), *ds = S.out(sig_s, count * 8);
    uint8_t *drec = S.out(recid, count), *dst = S.out(status, count);
    uint32_t *dp = S.tmp<uint32_t>(count * 64), *dq = S.tmp<uint32_t>(count * 64), *dm = S.tmp<uint32_t>(count * 64);
    if (S.err) return S.finish();
    RUN(decrypt_dev(c, ks, drow, dct, dp, dq, dm, (int)count));
    k_l17_p1_post<<<grid_for(count), 64, 0, c->stream>>>(dm, dk, dR, dr, ds, drec, dst, (int)count);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_l17_verify_batch(tecdsa_ctx* c, const uint32_t* sig_r, const uint32_t* sig_s, const uint32_t* pubkey, const uint32_t* message,
                                       uint8_t* status, size_t count, int mem) {
    if (!sig_r || !sig_s || !pubkey || !message || !status) return tecdsa_fail(TECDSA_E_ARG, "l17_verify: null argument");
    SIMPLE_PROLOGUE("l17_verify")
    const uint32_t *dr = S.in(sig_r, count * 8), *ds = S.in(sig_s, count * 8), *dy = S.in(pubkey, count * 16), *dm = S.in(message, count * 8);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_l17_verify<<<grid_for(count), 64, 0, c->stream>>>(dr, ds, dy, dm, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_l17_eph_create_batch(tecdsa_ctx* c, const uint32_t* secret_share, const uint32_t* nonce, const uint32_t* pk_blind,
                                           const uint32_t* zk_pok_blind, uint32_t* public_share, uint32_t* c_point, uint32_t* proof,
                                           uint32_t* pk_commitment, uint32_t* zk_pok_commitment, size_t count, int mem) {
    if (!secret_share || !nonce || !public_share || !c_point || !proof) return tecdsa_fail(TECDSA_E_ARG, "l17_eph_create: null argument");
    const bool with_com = pk_blind || zk_pok_blind || pk_commitment || zk_pok_commitment;
    if (with_com && !(pk_blind && zk_pok_blind && pk_commitment && zk_pok_commitment))
        return tecdsa_fail(TECDSA_E_ARG, "l17_eph_create: the four commitment buffers come together (party two) or not at all (party one)");
    SIMPLE_PROLOGUE("l17_eph_create")
    const uint32_t *dk = S.in(secret_share, count * 8), *dn = S.in(nonce, count * 8), *db1 = S.in(pk_blind, count * 8), *db2 = S.in(zk_pok_blind, count * 8);
    uint32_t *dpub = S.out(public_share, count * 16), *dc = S.out(c_point, count * 16), *dpf = S.out(proof, count * 40);
    uint32_t *dc1 = with_com ? S.out(pk_commitment, count * 8) : nullptr, *dc2 = with_com ? S.out(zk_pok_commitment, count * 8) : nullptr;
    if (S.err) return S.finish();
    k_l17_eph_create<<<grid_for(count), 64, 0, c->stream>>>(dk, dn, db1, db2, dpub, dc, dpf, dc1, dc2, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_l17_eph_verify_batch(tecdsa_ctx* c, const uint32_t* public_share, const uint32_t* c_point, const uint32_t* proof,
                                           const uint32_t* pk_blind, const uint32_t* zk_pok_blind, const uint32_t* pk_commitment,
                                           const uint32_t* zk_pok_commitment, uint8_t* status, size_t count, int mem) {
    if (!public_share || !c_point || !proof || !status) return tecdsa_fail(TECDSA_E_ARG, "l17_eph_verify: null argument");
    const bool with_com = pk_blind || zk_pok_blind || pk_commitment || zk_pok_commitment;
    if (with_com && !(pk_blind && zk_pok_blind && pk_commitment && zk_pok_commitment))
        return tecdsa_fail(TECDSA_E_ARG, "l17_eph_verify: the four commitment buffers come together or not at all");
    SIMPLE_PROLOGUE("l17_eph_verify")
    const uint32_t *dpub = S.in(public_share, count * 16), *dc = S.in(c_point, count * 16), *dpf = S.in(proof, count * 40), *db1 = S.in(pk_blind, count * 8),
                   *db2 = S.in(zk_pok_blind, count * 8), *dc1 = S.in(pk_commitment, count * 8), *dc2 = S.in(zk_pok_commitment, count * 8);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_l17_eph_verify<<<grid_for(count), 64, 0, c->stream>>>(dpub, dc, dpf, db1, db2, dc1, dc2, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_zkpdl_verifier_message1_batch(tecdsa_ctx* c, const uint32_t* n, const uint32_t* key_idx, size_t n_keys, const uint32_t* ciphertext,
                                                    const uint32_t* Q, const uint32_t* a, const uint32_t* b, const uint32_t* randomness,
                                                    const uint32_t* blindness, uint32_t* c_tag, uint32_t* c_tag_tag, uint32_t* q_tag, uint8_t* status,
                                                    size_t count, int mem) {
    if (!c || !n || !ciphertext || !Q || !a || !b || !randomness || !blindness || !c_tag || !c_tag_tag || !q_tag || !status)
        return tecdsa_fail(TECDSA_E_ARG, "zkpdl_verifier_message1: null argument");
    if (count == 0) return 0;
    CK(cudaSetDevice(c->device));
    const size_t nk = key_idx ? n_keys : count;
    Stage S(c, mem);
    const uint32_t *dn = S.in(n, nk * 64), *di = S.in(key_idx, count), *dct = S.in(ciphertext, count * 128), *dQ = S.in(Q, count * 16), *da = S.in(a, count * 8),
                   *db = S.in(b, count * 16), *dr = S.in(randomness, count * 64), *dbl = S.in(blindness, count * 8);
    uint32_t *dc = S.out(c_tag, count * 128), *dctt = S.out(c_tag_tag, count * 8), *dqt = S.out(q_tag, count * 16);
    uint8_t* dst = S.out(status, count);
    uint32_t *nd = S.tmp<uint32_t>(nk * NADIC_ROW * 64), *lin = S.tmp<uint32_t>(count * 128);
    if (S.err) return S.finish();
    RUN(c->nadic_setup(dn, nd, (int)nk, 64));
    k_zkpdl_v1_pre<<<grid_for(count), 64, 0, c->stream>>>(dn, di, dQ, da, db, dbl, lin, dctt, dqt, dst, (int)count);
    KCHECK();
    Launches L;
    const Operand N = di ? tab(dn, di, 64) : arr(dn, 64), ND = di ? tab(nd, di, NADIC_ROW * 64) : arr(nd, NADIC_ROW * 64);
    // c' = c^a * r^N * (1 + b N) mod N^2   (Paillier::mul, ::encrypt, ::add of zk_pdl/mod.rs:118-124 as one job)
    add_nn(L.e128, (int)count, N, ND, 2, arr(dr, 64), N, 64, arr(dct, 128), arr(da, 8), 8, 1, arr(lin, 128), NONE, dc, 128);
    RUN(run_nn(c, L.e128));
    return S.finish();
}

extern "C" int tecdsa_zkpdl_prover_message1_batch(tecdsa_ctx* c, const tecdsa_keyset* ks, const uint32_t* key_row, const uint32_t* c_tag,
                                                  const uint32_t* blindness, uint32_t* c_hat, uint32_t* q_hat, uint32_t* alpha, uint8_t* status,
                                                  size_t count, int mem) {
    if (!c || !ks || !key_row || !c_tag || !blindness || !c_hat || !q_hat || !alpha || !status)
        return tecdsa_fail(TECDSA_E_ARG, "zkpdl_prover_message1: null argument");
    if (count == 0) return 0;
    CK(cudaSetDevice(c->device));
    Stage S(c, mem);
    const uint32_t *drow = S.in(key_row, count), *dct = S.in(c_tag, count * 128), *dbl = S.in(blindness, count * 8);
    uint32_t *dch = S.out(c_hat, count * 8), *dqh = S.out(q_hat, count * 16), *dal = S.out(alpha, count * 64);
    uint8_t* dst = S.out(status, count);
    uint32_t *dp = S.tmp<uint32_t>(count * 64), *dq = S.tmp<uint32_t>(count * 64);
    if (S.err) return S.finish();
    RUN(decrypt_dev(c, ks, drow, dct, dp, dq, dal, (int)count));
    k_zkpdl_p1_post<<<grid_for(count), 64, 0, c->stream>>>(dal, dbl, dch, dqh, dst, (int)count);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_zkpdl_prover_message2_batch(tecdsa_ctx* c, const uint32_t* x1, const uint32_t* alpha, const uint32_t* c_tag_tag, const uint32_t* a,
                                                  const uint32_t* b, const uint32_t* blindness, uint8_t* status, size_t count, int mem) {
    if (!x1 || !alpha || !c_tag_tag || !a || !b || !blindness || !status) return tecdsa_fail(TECDSA_E_ARG, "zkpdl_prover_message2: null argument");
    SIMPLE_PROLOGUE("zkpdl_prover_message2")
    const uint32_t *dx = S.in(x1, count * 8), *dal = S.in(alpha, count * 64), *dctt = S.in(c_tag_tag, count * 8), *da = S.in(a, count * 8), *db = S.in(b, count * 16),
                   *dbl = S.in(blindness, count * 8);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_zkpdl_p2<<<grid_for(count), 64, 0, c->stream>>>(dx, dal, dctt, da, db, dbl, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_zkpdl_verifier_finalize_batch(tecdsa_ctx* c, const uint32_t* c_hat, const uint32_t* q_hat, const uint32_t* blindness,
                                                    const uint32_t* q_tag, uint8_t* status, size_t count, int mem) {
    if (!c_hat || !q_hat || !blindness || !q_tag || !status) return tecdsa_fail(TECDSA_E_ARG, "zkpdl_verifier_finalize: null argument");
    SIMPLE_PROLOGUE("zkpdl_verifier_finalize")
    const uint32_t *dch = S.in(c_hat, count * 8), *dqh = S.in(q_hat, count * 16), *dbl = S.in(blindness, count * 8), *dqt = S.in(q_tag, count * 16);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_zkpdl_finalize<<<grid_for(count), 64, 0, c->stream>>>(dch, dqh, dbl, dqt, o, n);
    KCHECK();
    return S.finish();
}
