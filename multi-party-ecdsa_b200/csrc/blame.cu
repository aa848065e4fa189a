// Identifiable-abort primitives of GG20 (/root/reference/src/protocols/multi_party_ecdsa/gg_2020/
// blame.rs): `Paillier::open` (plaintext AND encryption randomness of a ciphertext, :252-256) and curv's `ECDDHProof`
// (:258-271, 405-417).  The blame procedures themselves re-derive every opened value with the existing batch calls
// (multi-party-ecdsa_b200/blame.py); oracle: oracle/blame_oracle.py.
#include "stage.cuh"
#include "even_inverse.cuh"

using namespace tecdsa;

int tecdsa_internal_fb_points_set_blame(const uint32_t* table) {
    CK(cudaMemcpyToSymbol(secp::g_fb_points, &table, sizeof(table)));
    return 0;
}

namespace {

// phi(N) = (p - 1)(q - 1) of the key row of instance i
__global__ void k_phi(Arena A, const uint32_t* rows, uint32_t* phi64, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    st::mul(phi64 + (size_t)i * 64, A.k(KT_PM1, rows[i]), 32, A.k(KT_QM1, rows[i]), 32);
}
// rows[i]-th 64-limb row of a key table -> dense array
__global__ void k_gather_rows(const uint32_t* table, const uint32_t* rows, uint32_t* out64, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    st::copy(out64 + (size_t)i * 64, table + (size_t)rows[i] * 64, 64);
}
// curv ECDDHProof [R] (sigma_ec_ddh.rs): e = H(g1, h1, g2, h2, a1, a2) over 65-byte uncompressed points, reduced mod q
__device__ __forceinline__ U256 ecddh_hash(const Affine& g1, const Affine& h1, const Affine& g2, const Affine& h2, const Affine& a1, const Affine& a2) {
    Affine pts[6] = {g1, h1, g2, h2, a1, a2};
    return hash_points_scalar(pts, 6);
}
// prove: a1 = s g1, a2 = s g2, z = s + e x.  out: a1 16 | a2 16 | z 8
__global__ void k_ecddh_prove(const uint32_t* x8, const uint32_t* g1, const uint32_t* h1, const uint32_t* g2, const uint32_t* h2, const uint32_t* nonce8,
                              uint32_t* out40, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Affine G1 = affine_load(g1 + (size_t)i * 16), H1 = affine_load(h1 + (size_t)i * 16), G2 = affine_load(g2 + (size_t)i * 16), H2p = affine_load(h2 + (size_t)i * 16);
    const U256 s = load_scalar(nonce8 + (size_t)i * 8), x = load_scalar(x8 + (size_t)i * 8);
    Affine a1, a2;
    jac_to_affine2(a1, a2, jac_mul(jac_from_affine(G1), s), jac_mul(jac_from_affine(G2), s));
    const U256 e = ecddh_hash(G1, H1, G2, H2p, a1, a2);
    uint32_t* o = out40 + (size_t)i * 40;
    affine_store(o, a1); affine_store(o + 16, a2); u256_store(o + 32, sc_add(s, sc_mul(e, x)));
}
// verify: z g1 == a1 + e h1 and z g2 == a2 + e h2 (compared projectively)
__global__ void k_ecddh_verify(const uint32_t* pf40, const uint32_t* g1, const uint32_t* h1, const uint32_t* g2, const uint32_t* h2, uint8_t* status, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t* p = pf40 + (size_t)i * 40;
    const Affine G1 = affine_load(g1 + (size_t)i * 16), H1 = affine_load(h1 + (size_t)i * 16), G2 = affine_load(g2 + (size_t)i * 16), H2p = affine_load(h2 + (size_t)i * 16);
    const Affine a1 = affine_load(p), a2 = affine_load(p + 16);
    bool ok = on_curve(G1) && on_curve(H1) && on_curve(G2) && on_curve(H2p) && on_curve(a1) && on_curve(a2) && !a1.inf && !a2.inf && !G1.inf && !G2.inf && !H1.inf && !H2p.inf;
    if (ok) {
        const U256 z = load_scalar(p + 32);
        const U256 e = ecddh_hash(G1, H1, G2, H2p, a1, a2);
        const bool ok1 = jac_eq(jac_mul(jac_from_affine(G1), z), jac_madd(jac_mul(jac_from_affine(H1), e), a1));
        const bool ok2 = jac_eq(jac_mul(jac_from_affine(G2), z), jac_madd(jac_mul(jac_from_affine(H2p), e), a2));
        ok = ok1 && ok2;
    }
    status[i] = ok ? TECDSA_ST_OK : TECDSA_ST_PROOF;
}

}  // namespace

// Paillier::open(dk, c) [R] (kzen-paillier; call site blame.rs:252-256) over an uploaded key set: m = Dec(c) and the unique
// r in [0, N) with c = (1 + m N) r^N mod N^2, i.e. r = (c mod N)^(N^-1 mod phi(N)) mod N.
extern "C" int tecdsa_paillier_open_batch(tecdsa_ctx* c, const tecdsa_keyset* ks, const uint32_t* key_row, const uint32_t* cipher, uint32_t* m,
                                          uint32_t* r, size_t count, int mem) {
    if (!ks || !key_row || !cipher || !m || !r) return tecdsa_fail(TECDSA_E_ARG, "paillier_open: null argument");
    SIMPLE_PROLOGUE("paillier_open")
    const uint32_t *dr = S.in(key_row, count), *dc = S.in(cipher, count * 128);
    uint32_t *dm = S.out(m, count * 64), *drr = S.out(r, count * 64);
    uint32_t *phi = S.tmp<uint32_t>(count * 64), *phinv = S.tmp<uint32_t>(count * 64), *d = S.tmp<uint32_t>(count * 64);
    uint8_t *ok = S.tmp<uint8_t>(count), *st = S.tmp<uint8_t>(count);
    if (S.err) return S.finish();
    int rc = tecdsa_paillier_decrypt_batch(c, ks, dr, dc, dm, count, TECDSA_DEVICE);
    if (rc) { S.finish(); return rc; }
    if (cudaMemsetAsync(st, 0, count, c->stream) != cudaSuccess) { S.finish(); return tecdsa_fail(TECDSA_E_CUDA, "paillier_open: memset"); }
    k_phi<<<grid_for(count), 64, 0, c->stream>>>(key_arena(ks), dr, phi, n);
    KCHECK();
    Launches L;
    const Operand N = tab(ks->tab[KT_N], dr, 64);
    add_inv(L.i64, 64, n, N, arr(phi, 64), phinv, ok);                          // phi^-1 mod N
    RUN(run(c, L.i64, 64));
    // d = N^-1 mod phi from it (even_inverse.cuh); the moduli of an uploaded key set are products of two odd primes, so it exists
    {
        uint32_t* Nrows = S.tmp<uint32_t>(count * 64);
        if (S.err) return S.finish();
        k_gather_rows<<<grid_for(count), 64, 0, c->stream>>>(ks->tab[KT_N], dr, Nrows, n);
        KCHECK();
        k_inv_even_post<<<grid_for(count), 64, 0, c->stream>>>(Nrows, phi, phinv, ok, d, 0, st, n);
        KCHECK();
    }
    // r = (c mod N)^d mod N: the double-width ciphertext is reduced by the job itself (wide0)
    add_exp(L.e64, 64, n, N, 1, arr(dc, 128), arr(d, 64), 64, NONE, NONE, 0, 0, NONE, NONE, drr, 64);
    L.e64.cls[L.e64.n_classes - 1].wide0 = 1;
    RUN(run(c, L.e64, 64));
    return S.finish();
}

extern "C" int tecdsa_ecddh_prove_batch(tecdsa_ctx* c, const uint32_t* x, const uint32_t* g1, const uint32_t* h1, const uint32_t* g2, const uint32_t* h2,
                                        const uint32_t* nonce, uint32_t* proof, size_t count, int mem) {
    if (!x || !g1 || !h1 || !g2 || !h2 || !nonce || !proof) return tecdsa_fail(TECDSA_E_ARG, "ecddh_prove: null argument");
    SIMPLE_PROLOGUE("ecddh_prove")
    const uint32_t *dx = S.in(x, count * 8), *a = S.in(g1, count * 16), *b = S.in(h1, count * 16), *cc = S.in(g2, count * 16), *dd = S.in(h2, count * 16), *dn = S.in(nonce, count * 8);
    uint32_t* o = S.out(proof, count * 40);
    if (S.err) return S.finish();
    k_ecddh_prove<<<grid_for(count), 64, 0, c->stream>>>(dx, a, b, cc, dd, dn, o, n);
    KCHECK();
    return S.finish();
}
extern "C" int tecdsa_ecddh_verify_batch(tecdsa_ctx* c, const uint32_t* proof, const uint32_t* g1, const uint32_t* h1, const uint32_t* g2, const uint32_t* h2,
                                         uint8_t* status, size_t count, int mem) {
    if (!proof || !g1 || !h1 || !g2 || !h2 || !status) return tecdsa_fail(TECDSA_E_ARG, "ecddh_verify: null argument");
    SIMPLE_PROLOGUE("ecddh_verify")
    const uint32_t *p = S.in(proof, count * 40), *a = S.in(g1, count * 16), *b = S.in(h1, count * 16), *cc = S.in(g2, count * 16), *dd = S.in(h2, count * 16);
    uint8_t* st = S.out(status, count);
    if (S.err) return S.finish();
    k_ecddh_verify<<<grid_for(count), 64, 0, c->stream>>>(p, a, b, cc, dd, st, n);
    KCHECK();
    return S.finish();
}
