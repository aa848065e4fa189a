/* C twin of oracle/gg20_oracle.py `offline_session` over GMP + OpenSSL — TEST INFRASTRUCTURE ONLY.
 *
 * The whole GG20 OfflineStage (Round0..Round6, /root/reference/src/protocols/multi_party_ecdsa/gg_2020/state_machine/sign/
 * rounds.rs:68-636) of a batch of two-signer sessions, executed the way the REFERENCE executes it on a CPU: scalar calls,
 * GMP `mpz_powm` / `mpz_invert` behind `BigInt::mod_pow / mod_inv` (curv-kzen's default backend, Cargo.toml:29), libsecp256k1
 * replaced by OpenSSL's secp256k1, every redundant step kept (each `MessageB::b` re-verifies the three AliceProofs,
 * mta/mod.rs:123-131; Round5 verifies the party's own PDL proof, party_i.rs:719-766; Paillier decrypt recomputes its CRT
 * constants per call).  It takes the SAME packed inputs as tecdsa_gg20_offline_batch (include/tecdsa_b200.h) and produces
 * the same outputs, so it serves as (1) a second, independent implementation the Python restatement is pinned against
 * (tests/test_oracle.py), (2) the in-run parity checker of bench.py, (3) the timed CPU arm of bench.py ("port": the Rust
 * reference itself cannot be built in this image).  "parity unpinned" applies exactly as stated in gg20_oracle.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's checker / cpu_baseline / --impl reference legs may load it.
 */
#include "gmp_decl.h"
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <openssl/sha.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* randomness record layout: include/tecdsa_b200.h TECDSA_RND_* */
enum { RND_GAMMA = 0, RND_K = 8, RND_BLIND = 16, RND_RK = 24, RND_AL = 88, RND_AL_STRIDE = 248, RND_AL_ALPHA = 0, RND_AL_BETA = 24,
       RND_AL_GAMMA = 88, RND_AL_RHO = 176, RND_BT_G = 832, RND_R_G = 896, RND_NB_G = 960, RND_NBT_G = 968, RND_BT_W = 976,
       RND_R_W = 1040, RND_NB_W = 1104, RND_NBT_W = 1112, RND_L = 1120, RND_PED_S1 = 1128, RND_PED_S2 = 1136, RND_PDL_ALPHA = 1144,
       RND_PDL_BETA = 1168, RND_PDL_RHO = 1232, RND_PDL_GAMMA = 1304, RND_HEG_S1 = 1392, RND_HEG_S2 = 1400, RND_LIMBS = 1408 };
enum { ST_OK = 0, ST_INVALID_KEY = 2, ST_PDL = 6, ST_PHASE5 = 7, ST_PHASE6 = 8, ST_PROOF = 10, ST_COMMIT = 11 };

static void imp(mpz_t z, const uint32_t* limbs, int k) { mpz_import(z, (size_t)k, -1, 4, 0, 0, limbs); }
static void expo(uint32_t* limbs, int k, mpz_srcptr z) {
    size_t cnt = 0;
    memset(limbs, 0, (size_t)k * 4);
    mpz_export(limbs, &cnt, -1, 4, 0, 0, z);
}
/* curv `BigInt::to_bytes()`: big-endian magnitude, minimal length, zero -> one 0x00 byte [R] */
static size_t bn_bytes(unsigned char* out, mpz_srcptr z) {
    size_t cnt = 0;
    if (mpz_cmp_ui(z, 0) == 0) { out[0] = 0; return 1; }
    mpz_export(out, &cnt, 1, 1, 1, 0, z);
    return cnt;
}
static void fixed_be(unsigned char* out, size_t width, mpz_srcptr z) {
    unsigned char tmp[1024];
    size_t cnt = 0;
    memset(out, 0, width);
    if (mpz_cmp_ui(z, 0) == 0) return;
    mpz_export(tmp, &cnt, 1, 1, 1, 0, z);
    if (cnt > width) { memcpy(out, tmp + (cnt - width), width); return; }
    memcpy(out + (width - cnt), tmp, cnt);
}

/* ------------------------------------------------------------------------------------------------ thread context */
typedef struct {
    EC_GROUP* grp;
    BN_CTX* bctx;
    mpz_t q, q3;                /* group order, q^3 */
    EC_POINT* H2;               /* curv base_point2 */
    BIGNUM* bn_tmp;
    mpz_t t0, t1, t2, t3, t4;   /* scratch */
} tctx;

static void mpz_to_bn(BIGNUM* b, mpz_srcptr z) {
    unsigned char buf[80];
    size_t n = bn_bytes(buf, z);
    BN_bin2bn(buf, (int)n, b);
}
static void bn_to_mpz(mpz_t z, const BIGNUM* b) {
    unsigned char buf[80];
    int n = BN_bn2bin(b, buf);
    mpz_import(z, (size_t)n, 1, 1, 1, 0, buf);
}
static EC_POINT* pt_new(tctx* T) { return EC_POINT_new(T->grp); }
/* r = k * P (P == NULL: the generator); k any non-negative integer, reduced mod q */
static void pt_mul(tctx* T, EC_POINT* r, const EC_POINT* P, mpz_srcptr k) {
    mpz_mod(T->t4, k, T->q);
    mpz_to_bn(T->bn_tmp, T->t4);
    if (P) EC_POINT_mul(T->grp, r, NULL, P, T->bn_tmp, T->bctx);
    else EC_POINT_mul(T->grp, r, T->bn_tmp, NULL, NULL, T->bctx);
}
static void pt_add(tctx* T, EC_POINT* r, const EC_POINT* a, const EC_POINT* b) { EC_POINT_add(T->grp, r, a, b, T->bctx); }
static int pt_eq(tctx* T, const EC_POINT* a, const EC_POINT* b) { return EC_POINT_cmp(T->grp, a, b, T->bctx) == 0; }
static size_t pt_bytes(tctx* T, unsigned char* out, const EC_POINT* p, int compressed) {
    return EC_POINT_point2oct(T->grp, p, compressed ? POINT_CONVERSION_COMPRESSED : POINT_CONVERSION_UNCOMPRESSED, out, 80, T->bctx);
}
static void pt_from_limbs(tctx* T, EC_POINT* p, const uint32_t* xy16) {
    unsigned char buf[65];
    buf[0] = 4;
    for (int i = 0; i < 32; i++) { buf[1 + i] = (unsigned char)(xy16[7 - (i >> 2)] >> (8 * (3 - (i & 3)))); buf[33 + i] = (unsigned char)(xy16[15 - (i >> 2)] >> (8 * (3 - (i & 3)))); }
    int zero = 1;
    for (int i = 0; i < 16; i++) if (xy16[i]) zero = 0;
    if (zero) EC_POINT_set_to_infinity(T->grp, p);
    else EC_POINT_oct2point(T->grp, p, buf, 65, T->bctx);
}
static void pt_to_limbs(tctx* T, uint32_t* xy16, const EC_POINT* p) {
    unsigned char buf[80];
    memset(xy16, 0, 64);
    if (EC_POINT_is_at_infinity(T->grp, p)) return;
    pt_bytes(T, buf, p, 0);
    for (int i = 0; i < 32; i++) { xy16[7 - (i >> 2)] |= (uint32_t)buf[1 + i] << (8 * (3 - (i & 3))); xy16[15 - (i >> 2)] |= (uint32_t)buf[33 + i] << (8 * (3 - (i & 3))); }
}
/* `Sha256::new().chain_bigint(..)...result_bigint()` (curv DigestExt [R]) */
typedef struct { SHA256_CTX c; } hasher;
static void h_init(hasher* h) { SHA256_Init(&h->c); }
static void h_bigint(hasher* h, mpz_srcptr z) { unsigned char buf[1100]; size_t n = bn_bytes(buf, z); SHA256_Update(&h->c, buf, n); }
static void h_point_compressed(tctx* T, hasher* h, const EC_POINT* p) { unsigned char buf[80]; size_t n = pt_bytes(T, buf, p, 1); SHA256_Update(&h->c, buf, n); }
static void h_point_uncompressed(tctx* T, hasher* h, const EC_POINT* p) { unsigned char buf[80]; size_t n = pt_bytes(T, buf, p, 0); SHA256_Update(&h->c, buf, n); }
static void h_result_bigint(hasher* h, mpz_t out) { unsigned char d[32]; SHA256_Final(d, &h->c); mpz_import(out, 32, 1, 1, 1, 0, d); }
/* `H::new().chain_points([...]).result_scalar()`: 65-byte uncompressed points [R], digest mod q [R] */
static void hash_points_scalar(tctx* T, mpz_t out, int n, const EC_POINT** pts) {
    hasher h; h_init(&h);
    for (int i = 0; i < n; i++) h_point_uncompressed(T, &h, pts[i]);
    h_result_bigint(&h, out);
    mpz_mod(out, out, T->q);
}
/* HashCommitment::create_commitment_with_user_defined_randomness(from_bytes(P.to_bytes(true)), blind) (party_i.rs:577-580) */
static void hash_commitment(tctx* T, mpz_t out, const EC_POINT* P, mpz_srcptr blind) {
    hasher h; h_init(&h);
    h_point_compressed(T, &h, P);          /* 33 bytes starting with 02/03: already the minimal big-endian magnitude */
    h_bigint(&h, blind);
    h_result_bigint(&h, out);
}

/* ------------------------------------------------------------------------------------------------ keys */
typedef struct { mpz_t p, q, n, nn, nt, h1, h2, x_i; EC_POINT* pk; } key_row;
typedef struct { size_t rows; key_row* row; EC_POINT** y; } keyring;

/* ------------------------------------------------------------------------------------------------ Paillier (kzen-paillier [R]) */
static void paillier_encrypt(tctx* T, mpz_t c, const key_row* ek, mpz_srcptr m, mpz_srcptr r) {
    mpz_powm(c, r, ek->n, ek->nn);
    mpz_mul(T->t0, m, ek->n); mpz_add_ui(T->t0, T->t0, 1); mpz_mod(T->t0, T->t0, ek->nn);
    mpz_mul(c, c, T->t0); mpz_mod(c, c, ek->nn);
}
/* CRT decrypt, constants recomputed on every call as the crate does */
static void paillier_decrypt(tctx* T, mpz_t m, const key_row* dk, mpz_srcptr c) {
    mpz_t pp, qq, hp, hq, mp, mq, pinv, t;
    mpz_init(pp); mpz_init(qq); mpz_init(hp); mpz_init(hq); mpz_init(mp); mpz_init(mq); mpz_init(pinv); mpz_init(t);
    mpz_mul(pp, dk->p, dk->p); mpz_mul(qq, dk->q, dk->q);
    mpz_set_ui(t, 1); mpz_sub(t, t, dk->n); mpz_mod(t, t, pp); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, dk->p); mpz_invert(hp, t, dk->p);
    mpz_set_ui(t, 1); mpz_sub(t, t, dk->n); mpz_mod(t, t, qq); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, dk->q); mpz_invert(hq, t, dk->q);
    mpz_invert(pinv, dk->p, dk->q);
    mpz_mod(t, c, pp); mpz_sub_ui(mp, dk->p, 1); mpz_powm(t, t, mp, pp); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, dk->p); mpz_mul(t, t, hp); mpz_mod(mp, t, dk->p);
    mpz_mod(t, c, qq); mpz_sub_ui(mq, dk->q, 1); mpz_powm(t, t, mq, qq); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, dk->q); mpz_mul(t, t, hq); mpz_mod(mq, t, dk->q);
    mpz_sub(t, mq, mp); mpz_mod(t, t, dk->q); mpz_mul(t, t, pinv); mpz_mod(t, t, dk->q);
    mpz_mul(t, t, dk->p); mpz_add(m, t, mp);
    mpz_clear(pp); mpz_clear(qq); mpz_clear(hp); mpz_clear(hq); mpz_clear(mp); mpz_clear(mq); mpz_clear(pinv); mpz_clear(t);
    (void)T;
}

/* ------------------------------------------------------------------------------------------------ AliceProof (mta/range_proofs.rs) */
typedef struct { mpz_t z, e, s, s1, s2; } alice_pf;
static void alice_init(alice_pf* p) { mpz_init(p->z); mpz_init(p->e); mpz_init(p->s); mpz_init(p->s1); mpz_init(p->s2); }
static void alice_hash(mpz_t e, const key_row* ek, mpz_srcptr cipher, mpz_srcptr z, mpz_srcptr u, mpz_srcptr w, mpz_t tmp) {
    hasher h; h_init(&h);
    h_bigint(&h, ek->n);
    mpz_add_ui(tmp, ek->n, 1); h_bigint(&h, tmp);
    h_bigint(&h, cipher); h_bigint(&h, z); h_bigint(&h, u); h_bigint(&h, w);
    h_result_bigint(&h, e);
}
/* AliceProof::generate (range_proofs.rs:160-193) */
static void alice_generate(tctx* T, alice_pf* pf, mpz_srcptr a, mpz_srcptr cipher, const key_row* ek, const key_row* st, mpz_srcptr r,
                           mpz_srcptr alpha, mpz_srcptr beta, mpz_srcptr gamma, mpz_srcptr ro) {
    mpz_t u, w;
    mpz_init(u); mpz_init(w);
    mpz_powm(T->t0, st->h1, a, st->nt); mpz_powm(T->t1, st->h2, ro, st->nt); mpz_mul(pf->z, T->t0, T->t1); mpz_mod(pf->z, pf->z, st->nt);          /* :52 */
    mpz_mul(T->t0, alpha, ek->n); mpz_add_ui(T->t0, T->t0, 1); mpz_powm(T->t1, beta, ek->n, ek->nn); mpz_mul(u, T->t0, T->t1); mpz_mod(u, u, ek->nn); /* :53-55 */
    mpz_powm(T->t0, st->h1, alpha, st->nt); mpz_powm(T->t1, st->h2, gamma, st->nt); mpz_mul(w, T->t0, T->t1); mpz_mod(w, w, st->nt);                /* :56-57 */
    alice_hash(pf->e, ek, cipher, pf->z, u, w, T->t2);                                                                                              /* :174-182 */
    mpz_powm(T->t0, r, pf->e, ek->n); mpz_mul(pf->s, T->t0, beta); mpz_mod(pf->s, pf->s, ek->n);                                                   /* :86 */
    mpz_mul(pf->s1, pf->e, a); mpz_add(pf->s1, pf->s1, alpha);                                                                                     /* :87 */
    mpz_mul(pf->s2, pf->e, ro); mpz_add(pf->s2, pf->s2, gamma);                                                                                    /* :88 */
    mpz_clear(u); mpz_clear(w);
}
/* AliceProof::verify (range_proofs.rs:105-156) */
static int alice_verify(tctx* T, const alice_pf* pf, mpz_srcptr cipher, const key_row* ek, const key_row* st) {
    int ok = 0;
    mpz_t zei, cei, w, u, e;
    mpz_init(zei); mpz_init(cei); mpz_init(w); mpz_init(u); mpz_init(e);
    if (mpz_cmp(pf->s1, T->q3) > 0) goto done;                                                     /* :118 */
    mpz_powm(T->t0, pf->z, pf->e, st->nt);
    if (!mpz_invert(zei, T->t0, st->nt)) goto done;                                                 /* :122-127 */
    mpz_powm(T->t0, st->h1, pf->s1, st->nt); mpz_powm(T->t1, st->h2, pf->s2, st->nt);
    mpz_mul(w, T->t0, T->t1); mpz_mul(w, w, zei); mpz_mod(w, w, st->nt);                            /* :129-132 */
    mpz_powm(T->t0, cipher, pf->e, ek->nn);
    if (!mpz_invert(cei, T->t0, ek->nn)) goto done;                                                 /* :135-139 */
    mpz_mul(T->t0, pf->s1, ek->n); mpz_add_ui(T->t0, T->t0, 1); mpz_mod(T->t0, T->t0, ek->nn);      /* :134 */
    mpz_powm(T->t1, pf->s, ek->n, ek->nn);
    mpz_mul(u, T->t0, T->t1); mpz_mul(u, u, cei); mpz_mod(u, u, ek->nn);                            /* :141 */
    alice_hash(e, ek, cipher, pf->z, u, w, T->t2);                                                  /* :143-150 */
    ok = mpz_cmp(e, pf->e) == 0;
done:
    mpz_clear(zei); mpz_clear(cei); mpz_clear(w); mpz_clear(u); mpz_clear(e);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ curv sigma proofs [R] */
typedef struct { EC_POINT *pk, *Tc; mpz_t resp; } dlog_pf;
static void dlog_init(tctx* T, dlog_pf* p) { p->pk = pt_new(T); p->Tc = pt_new(T); mpz_init(p->resp); }
static void dlog_prove(tctx* T, dlog_pf* pf, mpz_srcptr sk, mpz_srcptr nonce) {
    pt_mul(T, pf->Tc, NULL, nonce);
    pt_mul(T, pf->pk, NULL, sk);
    const EC_POINT* pts[3] = {pf->Tc, EC_GROUP_get0_generator(T->grp), pf->pk};
    hash_points_scalar(T, T->t0, 3, pts);
    mpz_mul(T->t0, T->t0, sk); mpz_sub(pf->resp, nonce, T->t0); mpz_mod(pf->resp, pf->resp, T->q);
}
static int dlog_verify(tctx* T, const dlog_pf* pf) {
    const EC_POINT* pts[3] = {pf->Tc, EC_GROUP_get0_generator(T->grp), pf->pk};
    hash_points_scalar(T, T->t0, 3, pts);
    EC_POINT *a = pt_new(T), *b = pt_new(T);
    pt_mul(T, a, NULL, pf->resp); pt_mul(T, b, pf->pk, T->t0); pt_add(T, a, a, b);
    int ok = pt_eq(T, a, pf->Tc);
    EC_POINT_free(a); EC_POINT_free(b);
    return ok;
}
typedef struct { mpz_t e, z1, z2; EC_POINT *a1, *a2, *com; } ped_pf;
static void ped_init(tctx* T, ped_pf* p) { mpz_init(p->e); mpz_init(p->z1); mpz_init(p->z2); p->a1 = pt_new(T); p->a2 = pt_new(T); p->com = pt_new(T); }
static void pedersen_prove(tctx* T, ped_pf* pf, mpz_srcptr m, mpz_srcptr r, mpz_srcptr s1, mpz_srcptr s2) {
    EC_POINT* t = pt_new(T);
    pt_mul(T, pf->a1, NULL, s1); pt_mul(T, pf->a2, T->H2, s2);
    pt_mul(T, pf->com, NULL, m); pt_mul(T, t, T->H2, r); pt_add(T, pf->com, pf->com, t);
    const EC_POINT* pts[5] = {EC_GROUP_get0_generator(T->grp), T->H2, pf->com, pf->a1, pf->a2};
    hash_points_scalar(T, pf->e, 5, pts);
    mpz_mul(T->t0, pf->e, m); mpz_add(pf->z1, s1, T->t0); mpz_mod(pf->z1, pf->z1, T->q);
    mpz_mul(T->t0, pf->e, r); mpz_add(pf->z2, s2, T->t0); mpz_mod(pf->z2, pf->z2, T->q);
    EC_POINT_free(t);
}
static int pedersen_verify(tctx* T, const ped_pf* pf) {
    const EC_POINT* pts[5] = {EC_GROUP_get0_generator(T->grp), T->H2, pf->com, pf->a1, pf->a2};
    mpz_t e; mpz_init(e);
    hash_points_scalar(T, e, 5, pts);
    EC_POINT *l = pt_new(T), *t = pt_new(T), *r = pt_new(T);
    pt_mul(T, l, NULL, pf->z1); pt_mul(T, t, T->H2, pf->z2); pt_add(T, l, l, t);
    pt_add(T, r, pf->a1, pf->a2); pt_mul(T, t, pf->com, e); pt_add(T, r, r, t);
    int ok = pt_eq(T, l, r);
    EC_POINT_free(l); EC_POINT_free(t); EC_POINT_free(r); mpz_clear(e);
    return ok;
}
typedef struct { EC_POINT *Tt, *A3; mpz_t z1, z2; } heg_pf;
static void heg_init(tctx* T, heg_pf* p) { p->Tt = pt_new(T); p->A3 = pt_new(T); mpz_init(p->z1); mpz_init(p->z2); }
/* statement (G, H = base_point2, Y = generator, D, E), witness (x, r) — party_i.rs:778-799 */
static void heg_prove(tctx* T, heg_pf* pf, mpz_srcptr x, mpz_srcptr r, const EC_POINT* Gp, const EC_POINT* D, const EC_POINT* E, mpz_srcptr s1, mpz_srcptr s2) {
    EC_POINT* t = pt_new(T);
    pt_mul(T, pf->Tt, T->H2, s1); pt_mul(T, t, NULL, s2); pt_add(T, pf->Tt, pf->Tt, t);      /* A1 + A2 */
    pt_mul(T, pf->A3, Gp, s2);
    const EC_POINT* pts[7] = {pf->Tt, pf->A3, Gp, T->H2, EC_GROUP_get0_generator(T->grp), D, E};
    mpz_t e; mpz_init(e);
    hash_points_scalar(T, e, 7, pts);
    mpz_mod(T->t0, x, T->q);
    if (mpz_cmp_ui(T->t0, 0) != 0) { mpz_mul(T->t0, x, e); mpz_add(pf->z1, s1, T->t0); mpz_mod(pf->z1, pf->z1, T->q); } else mpz_set(pf->z1, s1);
    mpz_mul(T->t0, r, e); mpz_add(pf->z2, s2, T->t0); mpz_mod(pf->z2, pf->z2, T->q);
    EC_POINT_free(t); mpz_clear(e);
}
static int heg_verify(tctx* T, const heg_pf* pf, const EC_POINT* Gp, const EC_POINT* D, const EC_POINT* E) {
    const EC_POINT* pts[7] = {pf->Tt, pf->A3, Gp, T->H2, EC_GROUP_get0_generator(T->grp), D, E};
    mpz_t e; mpz_init(e);
    hash_points_scalar(T, e, 7, pts);
    EC_POINT *l = pt_new(T), *t = pt_new(T), *r = pt_new(T);
    pt_mul(T, l, T->H2, pf->z1); pt_mul(T, t, NULL, pf->z2); pt_add(T, l, l, t);
    pt_mul(T, t, D, e); pt_add(T, r, pf->Tt, t);
    int ok1 = pt_eq(T, l, r);
    pt_mul(T, l, Gp, pf->z2); pt_mul(T, t, E, e); pt_add(T, r, pf->A3, t);
    int ok2 = pt_eq(T, l, r);
    EC_POINT_free(l); EC_POINT_free(t); EC_POINT_free(r); mpz_clear(e);
    return ok1 && ok2;
}

/* ------------------------------------------------------------------------------------------------ MtA messages (mta/mod.rs) */
typedef struct { mpz_t c; alice_pf pf[3]; } msg_a;
typedef struct { mpz_t c; dlog_pf b, bt; } msg_b;
/* MessageB::b_with_predefined_randomness (mod.rs:111-158): 0 == Err(InvalidKey) */
static int message_b(tctx* T, msg_b* out, mpz_t beta_out, mpz_srcptr b, const key_row* ek, const msg_a* ma, mpz_srcptr randomness, mpz_srcptr beta_tag,
                     const key_row* const* stmts, mpz_srcptr nonce_b, mpz_srcptr nonce_beta) {
    for (int x = 0; x < 3; x++)
        if (!alice_verify(T, &ma->pf[x], ma->c, ek, stmts[x])) return 0;                            /* :123-131 */
    mpz_t fe, cbt, bca;
    mpz_init(fe); mpz_init(cbt); mpz_init(bca);
    mpz_mod(fe, beta_tag, T->q);                                                                     /* :132 */
    paillier_encrypt(T, cbt, ek, beta_tag, randomness);                                              /* :133 */
    mpz_powm(bca, ma->c, b, ek->nn);                                                                 /* :140 Paillier::mul */
    mpz_mul(out->c, bca, cbt); mpz_mod(out->c, out->c, ek->nn);                                      /* :145 Paillier::add */
    mpz_sub(beta_out, T->q, fe); mpz_mod(beta_out, beta_out, T->q);                                  /* :146 */
    dlog_prove(T, &out->b, b, nonce_b);
    dlog_prove(T, &out->bt, fe, nonce_beta);
    mpz_clear(fe); mpz_clear(cbt); mpz_clear(bca);
    return 1;
}
/* MessageB::verify_proofs_get_alpha (mod.rs:160-179) */
static int get_alpha(tctx* T, mpz_t alpha, const msg_b* mb, const key_row* dk, mpz_srcptr a) {
    mpz_t share; mpz_init(share);
    paillier_decrypt(T, share, dk, mb->c);
    mpz_mod(alpha, share, T->q);
    EC_POINT *ga = pt_new(T), *ba = pt_new(T);
    pt_mul(T, ga, NULL, alpha);
    pt_mul(T, ba, mb->b.pk, a); pt_add(T, ba, ba, mb->bt.pk);
    int ok = dlog_verify(T, &mb->b) && dlog_verify(T, &mb->bt) && pt_eq(T, ba, ga);
    EC_POINT_free(ga); EC_POINT_free(ba); mpz_clear(share);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ PDL with slack */
typedef struct { mpz_t z, u2, u3, s1, s2, s3; EC_POINT* u1; } pdl_pf;
static void pdl_init(tctx* T, pdl_pf* p) { mpz_init(p->z); mpz_init(p->u2); mpz_init(p->u3); mpz_init(p->s1); mpz_init(p->s2); mpz_init(p->s3); p->u1 = pt_new(T); }
/* commitment_unknown_order (zk_pdl_with_slack/mod.rs:182-199); returns 0 where the reference unwrap()-panics */
static int commit_uo(mpz_t out, mpz_srcptr h1, mpz_srcptr h2, mpz_srcptr nt, mpz_srcptr x, mpz_srcptr r, int r_negative, mpz_t s0, mpz_t s1) {
    mpz_powm(s0, h1, x, nt);
    if (r_negative) {
        if (!mpz_invert(s1, h2, nt)) return 0;
        mpz_powm(s1, s1, r, nt);                      /* r holds |r| */
    } else mpz_powm(s1, h2, r, nt);
    mpz_mul(out, s0, s1); mpz_mod(out, out, nt);
    return 1;
}
static void pdl_hash(tctx* T, mpz_t e, const EC_POINT* Gp, const EC_POINT* Qp, mpz_srcptr c, mpz_srcptr z, const EC_POINT* u1, mpz_srcptr u2, mpz_srcptr u3) {
    hasher h; h_init(&h);
    h_point_compressed(T, &h, Gp); h_point_compressed(T, &h, Qp);
    h_bigint(&h, c); h_bigint(&h, z);
    h_point_compressed(T, &h, u1);
    h_bigint(&h, u2); h_bigint(&h, u3);
    h_result_bigint(&h, e);
}
/* PDLwSlackProof::prove (mod.rs:68-125) */
static void pdl_prove(tctx* T, pdl_pf* pf, mpz_srcptr x, mpz_srcptr r, mpz_srcptr cipher, const key_row* ek, const EC_POINT* Qp, const EC_POINT* Gp,
                      const key_row* st, mpz_srcptr alpha, mpz_srcptr beta, mpz_srcptr rho, mpz_srcptr gamma) {
    mpz_t e, n1, one; mpz_init(e); mpz_init(n1); mpz_init(one);
    mpz_set_ui(one, 1);
    commit_uo(pf->z, st->h1, st->h2, st->nt, x, rho, 0, T->t0, T->t1);                               /* :78-84 */
    pt_mul(T, pf->u1, Gp, alpha);                                                                    /* :85 */
    mpz_add_ui(n1, ek->n, 1);
    commit_uo(pf->u2, n1, beta, ek->nn, alpha, ek->n, 0, T->t0, T->t1);                              /* :86-92 */
    commit_uo(pf->u3, st->h1, st->h2, st->nt, alpha, gamma, 0, T->t0, T->t1);                        /* :93-99 */
    pdl_hash(T, e, Gp, Qp, cipher, pf->z, pf->u1, pf->u2, pf->u3);                                   /* :101-109 */
    mpz_mul(pf->s1, e, x); mpz_add(pf->s1, pf->s1, alpha);                                           /* :112 */
    commit_uo(pf->s2, r, beta, ek->n, e, one, 0, T->t0, T->t1);                                      /* :113 */
    mpz_mul(pf->s3, e, rho); mpz_add(pf->s3, pf->s3, gamma);                                         /* :114 */
    mpz_clear(e); mpz_clear(n1); mpz_clear(one);
}
/* PDLwSlackProof::verify (mod.rs:127-179) */
static int pdl_verify(tctx* T, const pdl_pf* pf, mpz_srcptr cipher, const key_row* ek, const EC_POINT* Qp, const EC_POINT* Gp, const key_row* st) {
    int ok = 0;
    mpz_t e, n1, one, tmp, u2t, u3t; mpz_init(e); mpz_init(n1); mpz_init(one); mpz_init(tmp); mpz_init(u2t); mpz_init(u3t);
    EC_POINT *a = pt_new(T), *b = pt_new(T);
    mpz_set_ui(one, 1);
    pdl_hash(T, e, Gp, Qp, cipher, pf->z, pf->u1, pf->u2, pf->u3);
    pt_mul(T, a, Gp, pf->s1);
    mpz_mod(T->t0, e, T->q); mpz_sub(T->t0, T->q, T->t0);
    pt_mul(T, b, Qp, T->t0); pt_add(T, a, a, b);                                                     /* :138-142 */
    mpz_add_ui(n1, ek->n, 1);
    if (!commit_uo(tmp, n1, pf->s2, ek->nn, pf->s1, ek->n, 0, T->t0, T->t1)) goto done;              /* :144-150 */
    if (!commit_uo(u2t, tmp, cipher, ek->nn, one, e, 1, T->t0, T->t1)) goto done;                    /* :151-157 */
    if (!commit_uo(tmp, st->h1, st->h2, st->nt, pf->s1, pf->s3, 0, T->t0, T->t1)) goto done;         /* :158-164 */
    if (!commit_uo(u3t, tmp, pf->z, st->nt, one, e, 1, T->t0, T->t1)) goto done;                     /* :166-172 */
    ok = pt_eq(T, a, pf->u1) && mpz_cmp(u2t, pf->u2) == 0 && mpz_cmp(u3t, pf->u3) == 0;
done:
    EC_POINT_free(a); EC_POINT_free(b);
    mpz_clear(e); mpz_clear(n1); mpz_clear(one); mpz_clear(tmp); mpz_clear(u2t); mpz_clear(u3t);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ one session */
typedef struct {
    mpz_t w, gamma, k, blind, com, delta, sigma, l, beta_v, ni_v, alpha_v, mu_v;
    EC_POINT *gg, *Tp, *R, *Rd, *S;
    msg_a ma; msg_b mbg, mbw; ped_pf ped; pdl_pf pdl; heg_pf heg;
    SHA256_CTX tr;
    int status, completed;
} unit;
static void unit_init(tctx* T, unit* u) {
    mpz_init(u->w); mpz_init(u->gamma); mpz_init(u->k); mpz_init(u->blind); mpz_init(u->com); mpz_init(u->delta); mpz_init(u->sigma);
    mpz_init(u->l); mpz_init(u->beta_v); mpz_init(u->ni_v); mpz_init(u->alpha_v); mpz_init(u->mu_v);
    u->gg = pt_new(T); u->Tp = pt_new(T); u->R = pt_new(T); u->Rd = pt_new(T); u->S = pt_new(T);
    mpz_init(u->ma.c); for (int x = 0; x < 3; x++) alice_init(&u->ma.pf[x]);
    mpz_init(u->mbg.c); dlog_init(T, &u->mbg.b); dlog_init(T, &u->mbg.bt);
    mpz_init(u->mbw.c); dlog_init(T, &u->mbw.b); dlog_init(T, &u->mbw.bt);
    ped_init(T, &u->ped); pdl_init(T, &u->pdl); heg_init(T, &u->heg);
}
static void tr_int(unit* u, mpz_srcptr z, size_t width) { unsigned char buf[600]; fixed_be(buf, width, z); SHA256_Update(&u->tr, buf, width); }
static void tr_pt(tctx* T, unit* u, const EC_POINT* p) { unsigned char buf[80]; size_t n = pt_bytes(T, buf, p, 1); SHA256_Update(&u->tr, buf, n); }
/* curv map_share_to_new_params [R] for two signers: lambda_own = x_peer / (x_peer - x_own), points = index + 1 */
static void lagrange2(tctx* T, mpz_t out, unsigned own, unsigned peer) {
    mpz_set_ui(T->t0, peer + 1); mpz_set_ui(T->t1, own + 1);
    mpz_sub(T->t1, T->t0, T->t1); mpz_mod(T->t1, T->t1, T->q);
    mpz_invert(T->t1, T->t1, T->q);
    mpz_mul(out, T->t0, T->t1); mpz_mod(out, out, T->q);
}
typedef struct { mpz_t v; } mz;
#define RV(name, off, limbs) mpz_t name; mpz_init(name); imp(name, r + (off), (limbs))

static void session(tctx* T, const keyring* K, unit* U2, const uint32_t* sess, const uint32_t* rnd2, uint8_t* status2, uint32_t* R2, uint32_t* sigma2,
                    uint32_t* k2, uint32_t* tvec, uint32_t* digest2) {
    const unsigned kset = sess[0];
    const unsigned party[2] = {sess[1], sess[2]};
    const key_row* row[2] = {&K->row[kset * 3 + party[0]], &K->row[kset * 3 + party[1]]};
    const key_row* stmts[3] = {&K->row[kset * 3], &K->row[kset * 3 + 1], &K->row[kset * 3 + 2]};
    mpz_t tmp, tmp2; mpz_init(tmp); mpz_init(tmp2);
    EC_POINT *pa = pt_new(T), *pb = pt_new(T);
    int fail = 0;
    for (int p = 0; p < 2; p++) { U2[p].status = ST_OK; U2[p].completed = 0; SHA256_Init(&U2[p].tr); }
    /* ---- Round 0 (rounds.rs:68-104) */
    for (int p = 0; p < 2; p++) {
        unit* u = &U2[p];
        const uint32_t* r = rnd2 + (size_t)p * RND_LIMBS;
        lagrange2(T, tmp, party[p], party[1 - p]);
        mpz_mul(u->w, tmp, row[p]->x_i); mpz_mod(u->w, u->w, T->q);                                  /* party_i.rs:546-571 */
        imp(u->gamma, r + RND_GAMMA, 8); mpz_mod(u->gamma, u->gamma, T->q);
        imp(u->k, r + RND_K, 8); mpz_mod(u->k, u->k, T->q);
        imp(u->blind, r + RND_BLIND, 8);
        pt_mul(T, u->gg, NULL, u->gamma);
        hash_commitment(T, u->com, u->gg, u->blind);                                                 /* :573-589 */
        RV(rk, RND_RK, 64);
        paillier_encrypt(T, u->ma.c, row[p], u->k, rk);                                              /* mta/mod.rs:68 */
        for (int x = 0; x < 3; x++) {
            const uint32_t* al = r + RND_AL + x * RND_AL_STRIDE;
            mpz_t a_, b_, g_, r_; mpz_init(a_); mpz_init(b_); mpz_init(g_); mpz_init(r_);
            imp(a_, al + RND_AL_ALPHA, 24); imp(b_, al + RND_AL_BETA, 64); imp(g_, al + RND_AL_GAMMA, 88); imp(r_, al + RND_AL_RHO, 72);
            alice_generate(T, &u->ma.pf[x], u->k, u->ma.c, row[p], stmts[x], rk, a_, b_, g_, r_);
            mpz_clear(a_); mpz_clear(b_); mpz_clear(g_); mpz_clear(r_);
        }
        tr_int(u, u->ma.c, 512);
        for (int x = 0; x < 3; x++) { const alice_pf* f = &u->ma.pf[x]; tr_int(u, f->z, 256); tr_int(u, f->e, 32); tr_int(u, f->s, 256); tr_int(u, f->s1, 128); tr_int(u, f->s2, 384); }
        tr_int(u, u->com, 32);
        mpz_clear(rk);
    }
    /* ---- Round 1 (rounds.rs:122-206): two MessageB::b per party, each re-verifying the peer's three range proofs */
    for (int p = 0; p < 2; p++) {
        unit *u = &U2[p], *o = &U2[1 - p];
        const uint32_t* r = rnd2 + (size_t)p * RND_LIMBS;
        RV(btg, RND_BT_G, 64); RV(rg, RND_R_G, 64); RV(nbg, RND_NB_G, 8); RV(nbtg, RND_NBT_G, 8);
        RV(btw, RND_BT_W, 64); RV(rw, RND_R_W, 64); RV(nbw, RND_NB_W, 8); RV(nbtw, RND_NBT_W, 8);
        mpz_mod(nbg, nbg, T->q); mpz_mod(nbtg, nbtg, T->q); mpz_mod(nbw, nbw, T->q); mpz_mod(nbtw, nbtw, T->q);
        int ok = message_b(T, &u->mbg, u->beta_v, u->gamma, row[1 - p], &o->ma, rg, btg, stmts, nbg, nbtg);
        ok = ok && message_b(T, &u->mbw, u->ni_v, u->w, row[1 - p], &o->ma, rw, btw, stmts, nbw, nbtw);
        if (!ok) { u->status = ST_INVALID_KEY; fail = 1; }
        else {
            const msg_b* mbs[2] = {&u->mbg, &u->mbw};
            for (int m = 0; m < 2; m++) {
                tr_int(u, mbs[m]->c, 512);
                const dlog_pf* ds[2] = {&mbs[m]->b, &mbs[m]->bt};
                for (int d = 0; d < 2; d++) { tr_pt(T, u, ds[d]->pk); tr_pt(T, u, ds[d]->Tc); tr_int(u, ds[d]->resp, 32); }
            }
        }
        mpz_clear(btg); mpz_clear(rg); mpz_clear(nbg); mpz_clear(nbtg); mpz_clear(btw); mpz_clear(rw); mpz_clear(nbw); mpz_clear(nbtw);
    }
    if (fail) goto finish;
    /* ---- Round 2 (rounds.rs:234-317) */
    for (int p = 0; p < 2; p++) {
        unit *u = &U2[p], *o = &U2[1 - p];
        const uint32_t* r = rnd2 + (size_t)p * RND_LIMBS;
        lagrange2(T, tmp, party[1 - p], party[p]);
        pt_mul(T, pa, row[1 - p]->pk, tmp);                                                          /* g_w_vec[peer] (party_i.rs:527-544) */
        int ok = get_alpha(T, u->alpha_v, &o->mbg, row[p], u->k);
        ok = ok && get_alpha(T, u->mu_v, &o->mbw, row[p], u->k);
        ok = ok && pt_eq(T, o->mbw.b.pk, pa);                                                        /* rounds.rs:281 assert_eq! */
        if (!ok) { u->status = ST_INVALID_KEY; fail = 1; continue; }
        mpz_mul(u->delta, u->k, u->gamma); mpz_add(u->delta, u->delta, u->alpha_v); mpz_add(u->delta, u->delta, u->beta_v); mpz_mod(u->delta, u->delta, T->q);
        mpz_mul(u->sigma, u->k, u->w); mpz_add(u->sigma, u->sigma, u->mu_v); mpz_add(u->sigma, u->sigma, u->ni_v); mpz_mod(u->sigma, u->sigma, T->q);
        imp(u->l, r + RND_L, 8); mpz_mod(u->l, u->l, T->q);
        pt_mul(T, u->Tp, NULL, u->sigma); pt_mul(T, pb, T->H2, u->l); pt_add(T, u->Tp, u->Tp, pb);   /* party_i.rs:620-634 */
        RV(s1, RND_PED_S1, 8); RV(s2, RND_PED_S2, 8); mpz_mod(s1, s1, T->q); mpz_mod(s2, s2, T->q);
        pedersen_prove(T, &u->ped, u->sigma, u->l, s1, s2);
        tr_int(u, u->delta, 32); tr_pt(T, u, u->Tp); tr_int(u, u->ped.e, 32); tr_pt(T, u, u->ped.a1); tr_pt(T, u, u->ped.a2); tr_pt(T, u, u->ped.com);
        tr_int(u, u->ped.z1, 32); tr_int(u, u->ped.z2, 32);
        mpz_clear(s1); mpz_clear(s2);
    }
    if (fail) goto finish;
    /* ---- Round 3 (rounds.rs:347-402) */
    mpz_t dinv; mpz_init(dinv);
    mpz_add(dinv, U2[0].delta, U2[1].delta); mpz_mod(dinv, dinv, T->q);
    if (!mpz_invert(dinv, dinv, T->q)) { U2[0].status = U2[1].status = ST_PROOF; fail = 1; }       /* reference: unwrap() panic */
    for (int p = 0; p < 2 && !fail; p++) {
        unit* u = &U2[p];
        int ok = 1;
        for (int x = 0; x < 2; x++) ok = ok && pt_eq(T, U2[x].Tp, U2[x].ped.com);
        for (int x = 0; x < 2; x++) ok = ok && pedersen_verify(T, &U2[x].ped);
        if (!ok) { u->status = ST_PROOF; continue; }
        tr_int(u, u->blind, 32); tr_pt(T, u, u->gg);
    }
    if (U2[0].status || U2[1].status) { mpz_clear(dinv); goto finish; }
    /* ---- Round 4 (rounds.rs:431-498) */
    for (int p = 0; p < 2; p++) {
        unit *u = &U2[p], *o = &U2[1 - p];
        const uint32_t* r = rnd2 + (size_t)p * RND_LIMBS;
        hash_commitment(T, tmp, o->gg, o->blind);
        if (!(pt_eq(T, o->mbg.b.pk, o->gg) && mpz_cmp(tmp, o->com) == 0)) { u->status = ST_COMMIT; fail = 1; continue; }   /* party_i.rs:650-674 */
        pt_add(T, pa, U2[0].gg, U2[1].gg); pt_mul(T, u->R, pa, dinv);                                 /* :684-686 */
        pt_mul(T, u->Rd, u->R, u->k);                                                                 /* rounds.rs:452 */
        RV(rk, RND_RK, 64); RV(al, RND_PDL_ALPHA, 24); RV(be, RND_PDL_BETA, 64); RV(rh, RND_PDL_RHO, 72); RV(ga, RND_PDL_GAMMA, 88);
        pdl_prove(T, &u->pdl, u->k, rk, u->ma.c, row[p], u->Rd, u->R, row[1 - p], al, be, rh, ga);
        tr_pt(T, u, u->Rd); tr_int(u, u->pdl.z, 256); tr_pt(T, u, u->pdl.u1); tr_int(u, u->pdl.u2, 512); tr_int(u, u->pdl.u3, 256);
        tr_int(u, u->pdl.s1, 128); tr_int(u, u->pdl.s2, 256); tr_int(u, u->pdl.s3, 384);
        mpz_clear(rk); mpz_clear(al); mpz_clear(be); mpz_clear(rh); mpz_clear(ga);
    }
    mpz_clear(dinv);
    if (fail) goto finish;
    /* ---- Round 5 (rounds.rs:525-592): every signer's proof, own included */
    for (int p = 0; p < 2; p++) {
        unit* u = &U2[p];
        const uint32_t* r = rnd2 + (size_t)p * RND_LIMBS;
        int ok = 1;
        for (int x = 0; x < 2; x++) ok = ok && pdl_verify(T, &U2[x].pdl, U2[x].ma.c, row[x], U2[x].Rd, u->R, row[1 - x]);
        if (!ok) { u->status = ST_PDL; fail = 1; continue; }
        pt_add(T, pa, EC_GROUP_get0_generator(T->grp), U2[0].Rd); pt_add(T, pa, pa, U2[1].Rd);         /* party_i.rs:768-776 */
        EC_POINT_copy(pb, EC_GROUP_get0_generator(T->grp)); EC_POINT_invert(T->grp, pb, T->bctx); pt_add(T, pa, pa, pb);
        if (!pt_eq(T, pa, EC_GROUP_get0_generator(T->grp))) { u->status = ST_PHASE5; fail = 1; continue; }
        pt_mul(T, u->S, u->R, u->sigma);                                                              /* :784 */
        RV(s1, RND_HEG_S1, 8); RV(s2, RND_HEG_S2, 8); mpz_mod(s1, s1, T->q); mpz_mod(s2, s2, T->q);
        heg_prove(T, &u->heg, u->l, u->sigma, u->R, u->Tp, u->S, s1, s2);
        tr_pt(T, u, u->S); tr_pt(T, u, u->heg.Tt); tr_pt(T, u, u->heg.A3); tr_int(u, u->heg.z1, 32); tr_int(u, u->heg.z2, 32);
        mpz_clear(s1); mpz_clear(s2);
    }
    if (fail) goto finish;
    /* ---- Round 6 (rounds.rs:612-636) */
    for (int p = 0; p < 2; p++) {
        unit* u = &U2[p];
        int ok = 1;
        for (int x = 0; x < 2; x++) ok = ok && heg_verify(T, &U2[x].heg, u->R, U2[x].Tp, U2[x].S);   /* party_i.rs:801-833 */
        if (!ok) { u->status = ST_PHASE6; continue; }
        pt_add(T, pa, EC_GROUP_get0_generator(T->grp), U2[0].S); pt_add(T, pa, pa, U2[1].S);
        EC_POINT_copy(pb, EC_GROUP_get0_generator(T->grp)); EC_POINT_invert(T->grp, pb, T->bctx); pt_add(T, pa, pa, pb);
        if (!pt_eq(T, pa, K->y[kset])) { u->status = ST_PHASE6; continue; }                           /* :835-848 */
        u->completed = 1;
    }
finish:
    for (int p = 0; p < 2; p++) {
        unit* u = &U2[p];
        unsigned char d[32];
        SHA256_Final(d, &u->tr);
        status2[p] = (uint8_t)u->status;
        if (digest2) for (int i = 0; i < 8; i++) digest2[p * 8 + i] = ((uint32_t)d[31 - 4 * i]) | ((uint32_t)d[30 - 4 * i] << 8) | ((uint32_t)d[29 - 4 * i] << 16) | ((uint32_t)d[28 - 4 * i] << 24);
        const int done = u->completed;
        if (R2) { if (done) pt_to_limbs(T, R2 + p * 16, u->R); else memset(R2 + p * 16, 0, 64); }
        if (sigma2) { if (done) expo(sigma2 + p * 8, 8, u->sigma); else memset(sigma2 + p * 8, 0, 32); }
        if (k2) { if (done) expo(k2 + p * 8, 8, u->k); else memset(k2 + p * 8, 0, 32); }
    }
    if (tvec) {
        if (U2[0].completed && U2[1].completed) { pt_to_limbs(T, tvec, U2[0].Tp); pt_to_limbs(T, tvec + 16, U2[1].Tp); pt_to_limbs(T, tvec + 32, U2[0].Tp); pt_to_limbs(T, tvec + 48, U2[1].Tp); }
        else memset(tvec, 0, 256);
    }
    EC_POINT_free(pa); EC_POINT_free(pb); mpz_clear(tmp); mpz_clear(tmp2);
}

/* ------------------------------------------------------------------------------------------------ batch driver + persistent pool */
typedef struct {
    const keyring* K; const uint32_t *sessions, *rnd; size_t n_sessions;
    uint8_t* status; uint32_t *R, *sigma, *k, *tvec, *digest;
    volatile size_t next;
} batch_job;

static void tctx_init(tctx* T) {
    T->grp = EC_GROUP_new_by_curve_name(NID_secp256k1);
    T->bctx = BN_CTX_new();
    T->bn_tmp = BN_new();
    mpz_init(T->q); mpz_init(T->q3); mpz_init(T->t0); mpz_init(T->t1); mpz_init(T->t2); mpz_init(T->t3); mpz_init(T->t4);
    BIGNUM* order = BN_new();
    EC_GROUP_get_order(T->grp, order, T->bctx);
    bn_to_mpz(T->q, order);
    BN_free(order);
    mpz_mul(T->q3, T->q, T->q); mpz_mul(T->q3, T->q3, T->q);
    /* curv base_point2: x = SHA256^3(compressed G) (checked in tests/test_oracle.py) */
    static const uint32_t H2[16] = {0x0378B795u, 0xA8DC7BFAu, 0x5FF3CE66u, 0xDD142E4Bu, 0x4BA80116u, 0x34DD4521u, 0xE3A7326Au, 0x08D13221u,
                                    0xF7C2BE88u, 0x8217E9F7u, 0xDF0DF07Au, 0x807BCBA1u, 0xBD565EA2u, 0x0848D50Du, 0x77614B5Cu, 0x5D41AC14u};
    T->H2 = pt_new(T);
    pt_from_limbs(T, T->H2, H2);
}
static void batch_worker(batch_job* j, tctx* T, unit* U2) {
    for (;;) {
        size_t s = __sync_fetch_and_add(&j->next, 1);
        if (s >= j->n_sessions) break;
        session(T, j->K, U2, j->sessions + 3 * s, j->rnd + 2 * s * RND_LIMBS, j->status + 2 * s, j->R ? j->R + 32 * s : NULL,
                j->sigma ? j->sigma + 16 * s : NULL, j->k ? j->k + 16 * s : NULL, j->tvec ? j->tvec + 64 * s : NULL, j->digest ? j->digest + 16 * s : NULL);
    }
}

/* persistent pool: workers keep their thread context (EC group, scratch, unit state) across calls */
typedef struct { pthread_t th; int id; } pool_thread;
static struct {
    pthread_mutex_t mu; pthread_cond_t go, done;
    pool_thread* t; int n; unsigned long gen; int pending; batch_job* job; int active;
} POOL = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, NULL, 0};

static void* pool_main(void* arg) {
    pool_thread* me = (pool_thread*)arg;
    tctx T; unit U2[2];
    tctx_init(&T); unit_init(&T, &U2[0]); unit_init(&T, &U2[1]);
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&POOL.mu);
        while (POOL.gen == seen) pthread_cond_wait(&POOL.go, &POOL.mu);
        seen = POOL.gen;
        batch_job* j = POOL.job;
        const int mine = me->id < POOL.active;
        pthread_mutex_unlock(&POOL.mu);
        if (mine) batch_worker(j, &T, U2);
        pthread_mutex_lock(&POOL.mu);
        if (--POOL.pending == 0) pthread_cond_signal(&POOL.done);
        pthread_mutex_unlock(&POOL.mu);
    }
    return NULL;
}
static void pool_run(batch_job* j, int nthreads) {
    pthread_mutex_lock(&POOL.mu);
    if (POOL.n < nthreads) {
        POOL.t = (pool_thread*)realloc(POOL.t, sizeof(pool_thread) * (size_t)nthreads);
        /* realloc may move the array: threads hold their id by value through a heap copy */
        for (int i = POOL.n; i < nthreads; i++) {
            pool_thread* slot = (pool_thread*)malloc(sizeof(pool_thread));
            slot->id = i;
            pthread_create(&slot->th, NULL, pool_main, slot);
            POOL.t[i] = *slot;
        }
        POOL.n = nthreads;
    }
    POOL.job = j; POOL.active = nthreads; POOL.pending = POOL.n; POOL.gen++;
    pthread_cond_broadcast(&POOL.go);
    while (POOL.pending) pthread_cond_wait(&POOL.done, &POOL.mu);
    pthread_mutex_unlock(&POOL.mu);
}

/* Inputs exactly as tecdsa_keys / tecdsa_gg20_offline_batch take them (include/tecdsa_b200.h); k_out = the k_i of each unit
 * (8 limbs, what CompletedOfflineStage.sign_keys holds).  Any output except status may be NULL. */
int oracle_gg20_offline_batch(size_t n_keysets, const uint32_t* p32, const uint32_t* q32, const uint32_t* nt64, const uint32_t* h1_64,
                              const uint32_t* h2_64, const uint32_t* x8, const uint32_t* pk16, const uint32_t* y16,
                              const uint32_t* sessions, size_t n_sessions, const uint32_t* rnd, uint8_t* status, uint32_t* R, uint32_t* sigma,
                              uint32_t* k_out, uint32_t* tvec, uint32_t* digest, int nthreads) {
    static __thread tctx T0; static __thread int t0_ready = 0; static __thread unit U1[2];
    if (!t0_ready) { tctx_init(&T0); unit_init(&T0, &U1[0]); unit_init(&T0, &U1[1]); t0_ready = 1; }
    keyring K;
    K.rows = n_keysets * 3;
    K.row = (key_row*)calloc(K.rows, sizeof(key_row));
    K.y = (EC_POINT**)calloc(n_keysets, sizeof(EC_POINT*));
    for (size_t r = 0; r < K.rows; r++) {
        key_row* k = &K.row[r];
        mpz_init(k->p); mpz_init(k->q); mpz_init(k->n); mpz_init(k->nn); mpz_init(k->nt); mpz_init(k->h1); mpz_init(k->h2); mpz_init(k->x_i);
        imp(k->p, p32 + r * 32, 32); imp(k->q, q32 + r * 32, 32); imp(k->nt, nt64 + r * 64, 64); imp(k->h1, h1_64 + r * 64, 64); imp(k->h2, h2_64 + r * 64, 64);
        imp(k->x_i, x8 + r * 8, 8);
        mpz_mul(k->n, k->p, k->q); mpz_mul(k->nn, k->n, k->n);
        k->pk = pt_new(&T0); pt_from_limbs(&T0, k->pk, pk16 + r * 16);
    }
    for (size_t s = 0; s < n_keysets; s++) { K.y[s] = pt_new(&T0); pt_from_limbs(&T0, K.y[s], y16 + s * 16); }
    for (size_t s = 0; s < n_sessions; s++)
        if (sessions[3 * s] >= n_keysets || sessions[3 * s + 1] > 2 || sessions[3 * s + 2] > 2 || sessions[3 * s + 1] == sessions[3 * s + 2]) return -1;
    batch_job j = {&K, sessions, rnd, n_sessions, status, R, sigma, k_out, tvec, digest, 0};
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n_sessions) nthreads = (int)(n_sessions ? n_sessions : 1);
    if (nthreads == 1) batch_worker(&j, &T0, U1);
    else pool_run(&j, nthreads);
    for (size_t r = 0; r < K.rows; r++) {
        key_row* k = &K.row[r];
        mpz_clear(k->p); mpz_clear(k->q); mpz_clear(k->n); mpz_clear(k->nn); mpz_clear(k->nt); mpz_clear(k->h1); mpz_clear(k->h2); mpz_clear(k->x_i);
        EC_POINT_free(k->pk);
    }
    for (size_t s = 0; s < n_keysets; s++) EC_POINT_free(K.y[s]);
    free(K.row); free(K.y);
    return 0;
}
