// GG18 signing phases that GG20 replaced, as batched kernels: phase 4 and phases 5a-5d plus
// `output_signature` of /root/reference/src/protocols/multi_party_ecdsa/gg_2018/party_i.rs:455-730.  Phases 1-3 are the MtA of the
// hot path with an empty statement list (tecdsa_mta_message_{a,b}_batch with n_st = 0, tecdsa_mta_get_alpha_batch) and scalar sums.
// Element layout: a batch holds `sessions` signing sessions of `parties` signers each; element u = session * parties + party.
// Every array is element-major; "the other signers" of an element are the other elements of its session.  One thread per
// element: the work is secp256k1 and SHA-256 only.  Oracle: oracle/gg18_oracle.py.
#include "stage.cuh"
#include "gg18_kernels.cuh"

using namespace tecdsa;

int tecdsa_internal_fb_points_set_gg18(const uint32_t* table) {
    CK(cudaMemcpyToSymbol(secp::g_fb_points, &table, sizeof(table)));
    return 0;
}

using namespace tecdsa::gg18;


#define GG18_PROLOGUE(name)                                                                                           \
    if (!c) return tecdsa_fail(TECDSA_E_ARG, name ": null ctx");                                                      \
    if (parties < 2 || parties > 64) return tecdsa_fail(TECDSA_E_ARG, name ": parties must be in 2..64");             \
    if (sessions == 0) return 0;                                                                                      \
    if (sessions > ((size_t)1 << 24)) return tecdsa_fail(TECDSA_E_ARG, name ": too many sessions");                   \
    CK(cudaSetDevice(c->device));                                                                                     \
    const size_t count = sessions * (size_t)parties;                                                                  \
    const int n = (int)count;                                                                                         \
    Stage S(c, mem);

extern "C" int tecdsa_gg18_phase4_batch(tecdsa_ctx* c, int parties, const uint32_t* delta_inv, const uint32_t* b_proof_pk, const uint32_t* g_gamma,
                                        const uint32_t* blind, const uint32_t* com, uint32_t* R, uint8_t* status, size_t sessions, int mem) {
    if (!delta_inv || !b_proof_pk || !g_gamma || !blind || !com || !R || !status) return tecdsa_fail(TECDSA_E_ARG, "gg18_phase4: null argument");
    GG18_PROLOGUE("gg18_phase4")
    const uint32_t *dd = S.in(delta_inv, count * 8), *dpk = S.in(b_proof_pk, count * parties * 16), *dg = S.in(g_gamma, count * 16),
                   *db = S.in(blind, count * 8), *dc = S.in(com, count * 8);
    uint32_t* dR = S.out(R, count * 16);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_gg18_phase4<<<grid_for(count), 64, 0, c->stream>>>(parties, dd, dpk, dg, db, dc, dR, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_gg18_local_sig_batch(tecdsa_ctx* c, const uint32_t* message, const uint32_t* R, const uint32_t* k_i, const uint32_t* sigma_i,
                                           uint32_t* s_i, size_t count, int mem) {
    if (!message || !R || !k_i || !sigma_i || !s_i) return tecdsa_fail(TECDSA_E_ARG, "gg18_local_sig: null argument");
    SIMPLE_PROLOGUE("gg18_local_sig")
    const uint32_t *dm = S.in(message, count * 8), *dR = S.in(R, count * 16), *dk = S.in(k_i, count * 8), *dsg = S.in(sigma_i, count * 8);
    uint32_t* o = S.out(s_i, count * 8);
    if (S.err) return S.finish();
    k_gg18_local_sig<<<grid_for(count), 64, 0, c->stream>>>(dm, dR, dk, dsg, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_gg18_phase5a_batch(tecdsa_ctx* c, const uint32_t* R, const uint32_t* s_i, const uint32_t* l_i, const uint32_t* rho_i,
                                         const uint32_t* blind, const uint32_t* heg_s1, const uint32_t* heg_s2, const uint32_t* dlog_nonce,
                                         uint32_t* com, uint32_t* decom, uint32_t* heg_proof, uint32_t* dlog_proof, uint8_t* status, size_t count, int mem) {
    if (!R || !s_i || !l_i || !rho_i || !blind || !heg_s1 || !heg_s2 || !dlog_nonce || !com || !decom || !heg_proof || !dlog_proof || !status)
        return tecdsa_fail(TECDSA_E_ARG, "gg18_phase5a: null argument");
    SIMPLE_PROLOGUE("gg18_phase5a")
    const uint32_t *dR = S.in(R, count * 16), *ds = S.in(s_i, count * 8), *dl = S.in(l_i, count * 8), *drho = S.in(rho_i, count * 8), *db = S.in(blind, count * 8),
                   *d1 = S.in(heg_s1, count * 8), *d2 = S.in(heg_s2, count * 8), *dn = S.in(dlog_nonce, count * 8);
    uint32_t *oc = S.out(com, count * 8), *od = S.out(decom, count * 48), *oh = S.out(heg_proof, count * 48), *og = S.out(dlog_proof, count * 40);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    CK(cudaMemsetAsync(oc, 0, count * 32, c->stream)); CK(cudaMemsetAsync(od, 0, count * 192, c->stream));
    CK(cudaMemsetAsync(oh, 0, count * 192, c->stream)); CK(cudaMemsetAsync(og, 0, count * 160, c->stream));
    k_gg18_phase5a<<<grid_for(count), 64, 0, c->stream>>>(dR, ds, dl, drho, db, d1, d2, dn, oc, od, oh, og, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_gg18_phase5c_batch(tecdsa_ctx* c, int parties, const uint32_t* R, const uint32_t* y, const uint32_t* message, const uint32_t* rho_i,
                                         const uint32_t* l_i, const uint32_t* blind2, const uint32_t* com, const uint32_t* decom, const uint32_t* blind,
                                         const uint32_t* heg_proof, const uint32_t* dlog_proof, uint32_t* com2, uint32_t* decom2, uint8_t* status,
                                         size_t sessions, int mem) {
    if (!R || !y || !message || !rho_i || !l_i || !blind2 || !com || !decom || !blind || !heg_proof || !dlog_proof || !com2 || !decom2 || !status)
        return tecdsa_fail(TECDSA_E_ARG, "gg18_phase5c: null argument");
    GG18_PROLOGUE("gg18_phase5c")
    const uint32_t *dR = S.in(R, count * 16), *dy = S.in(y, count * 16), *dm = S.in(message, count * 8), *drho = S.in(rho_i, count * 8), *dl = S.in(l_i, count * 8),
                   *db2 = S.in(blind2, count * 8), *dc = S.in(com, count * 8), *dd = S.in(decom, count * 48), *db = S.in(blind, count * 8),
                   *dh = S.in(heg_proof, count * 48), *dg = S.in(dlog_proof, count * 40);
    uint32_t *oc = S.out(com2, count * 8), *od = S.out(decom2, count * 32);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_gg18_phase5c<<<grid_for(count), 64, 0, c->stream>>>(parties, dR, dy, dm, drho, dl, db2, dc, dd, db, dh, dg, oc, od, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_gg18_phase5d_batch(tecdsa_ctx* c, int parties, const uint32_t* decom2, const uint32_t* blind2, const uint32_t* com2,
                                         const uint32_t* decom, uint8_t* status, size_t sessions, int mem) {
    if (!decom2 || !blind2 || !com2 || !decom || !status) return tecdsa_fail(TECDSA_E_ARG, "gg18_phase5d: null argument");
    GG18_PROLOGUE("gg18_phase5d")
    const uint32_t *dd2 = S.in(decom2, count * 32), *db2 = S.in(blind2, count * 8), *dc2 = S.in(com2, count * 8), *dd = S.in(decom, count * 48);
    uint8_t* o = S.out(status, count);
    if (S.err) return S.finish();
    k_gg18_phase5d<<<grid_for(count), 64, 0, c->stream>>>(parties, dd2, db2, dc2, dd, o, n);
    KCHECK();
    return S.finish();
}

extern "C" int tecdsa_gg18_output_signature_batch(tecdsa_ctx* c, int parties, const uint32_t* R, const uint32_t* y, const uint32_t* message,
                                                  const uint32_t* s_i, uint32_t* sig_r, uint32_t* sig_s, uint8_t* recid, uint8_t* status,
                                                  size_t sessions, int mem) {
    if (!R || !y || !message || !s_i || !sig_r || !sig_s || !recid || !status) return tecdsa_fail(TECDSA_E_ARG, "gg18_output_signature: null argument");
    GG18_PROLOGUE("gg18_output_signature")
    const uint32_t *dR = S.in(R, count * 16), *dy = S.in(y, count * 16), *dm = S.in(message, count * 8), *ds = S.in(s_i, count * 8);
    uint32_t *orr = S.out(sig_r, count * 8), *os = S.out(sig_s, count * 8);
    uint8_t *orec = S.out(recid, count), *o = S.out(status, count);
    if (S.err) return S.finish();
    k_gg18_output<<<grid_for(count), 64, 0, c->stream>>>(parties, dR, dy, dm, ds, orr, os, orec, o, n);
    KCHECK();
    return S.finish();
}
