/* C twin of oracle/gg20_oracle.py over GMP — TEST INFRASTRUCTURE ONLY.
 *
 * It exists for two things: (1) pinning the Python restatement's big-integer arithmetic
 * against the very library the reference's default backend calls (GMP mpz_powm /
 * mpz_invert, curv-kzen `rust-gmp-kzen`, /root/reference/Cargo.toml:29), and (2) being the
 * timed CPU baseline ("port") of bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  "parity unpinned" applies
 * exactly as stated in gg20_oracle.py.
 *
 * Layout conventions match include/tecdsa_b200.h: little-endian uint32 limbs, operand-major.
 */
#include "gmp_decl.h"
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

const char* oracle_gmp_version(void) { return __gmp_version; }

static void imp(mpz_t z, const uint32_t* limbs, int k) { mpz_import(z, (size_t)k, -1, 4, 0, 0, limbs); }
static void expo(uint32_t* limbs, int k, mpz_t z) {
    size_t cnt = 0;
    memset(limbs, 0, (size_t)k * 4);
    mpz_export(limbs, &cnt, -1, 4, 0, 0, z);
}

/* BigInt::mod_pow == mpz_powm (call sites: src/utilities/mta/range_proofs.rs:52-57,86,122-141) */
typedef struct {
    const uint32_t *base, *exp, *mod, *mod_idx; uint32_t* out;
    size_t lo, hi; int k, el;
} modexp_job;

static void* modexp_worker(void* arg) {
    modexp_job* j = (modexp_job*)arg;
    mpz_t b, e, m, r;
    mpz_init(b); mpz_init(e); mpz_init(m); mpz_init(r);
    for (size_t i = j->lo; i < j->hi; i++) {
        size_t mi = j->mod_idx ? j->mod_idx[i] : i;
        imp(b, j->base + i * j->k, j->k);
        imp(e, j->exp + i * j->el, j->el);
        imp(m, j->mod + mi * j->k, j->k);
        mpz_powm(r, b, e, m);
        expo(j->out + i * j->k, j->k, r);
    }
    mpz_clear(b); mpz_clear(e); mpz_clear(m); mpz_clear(r);
    return NULL;
}

int oracle_modexp_batch(const uint32_t* base, const uint32_t* exp, const uint32_t* mod, const uint32_t* mod_idx,
                        uint32_t* out, size_t count, int k, int el, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > count) nthreads = (int)(count ? count : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    modexp_job* jobs = (modexp_job*)malloc(sizeof(modexp_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (modexp_job){base, exp, mod, mod_idx, out, count * t / nthreads, count * (t + 1) / nthreads, k, el};
        if (nthreads == 1) modexp_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, modexp_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* kzen-paillier `encrypt_with_chosen_randomness` then CRT `decrypt` (BASELINE.json configs[0];
 * call sites src/utilities/mta/mod.rs:68,165).  n = p*q (k limbs), m (k limbs), r (k limbs);
 * writes c (2k limbs) and the decrypted plaintext (k limbs). */
int oracle_paillier_roundtrip(const uint32_t* p_l, const uint32_t* q_l, const uint32_t* m_l, const uint32_t* r_l,
                              uint32_t* c_out, uint32_t* m_out, int k) {
    mpz_t p, q, n, nn, m, r, c, t, pp, qq, hp, hq, mp, mq, pinv, one;
    mpz_init(p); mpz_init(q); mpz_init(n); mpz_init(nn); mpz_init(m); mpz_init(r); mpz_init(c); mpz_init(t);
    mpz_init(pp); mpz_init(qq); mpz_init(hp); mpz_init(hq); mpz_init(mp); mpz_init(mq); mpz_init(pinv); mpz_init(one);
    imp(p, p_l, k / 2); imp(q, q_l, k / 2); imp(m, m_l, k); imp(r, r_l, k);
    mpz_set_ui(one, 1);
    mpz_mul(n, p, q); mpz_mul(nn, n, n);
    /* c = (1 + m n) r^n mod n^2 */
    mpz_powm(c, r, n, nn);
    mpz_mul(t, m, n); mpz_add_ui(t, t, 1); mpz_mod(t, t, nn);
    mpz_mul(c, c, t); mpz_mod(c, c, nn);
    expo(c_out, 2 * k, c);
    /* decrypt: recomputes its CRT constants on every call, as the reference does */
    mpz_mul(pp, p, p); mpz_mul(qq, q, q);
    mpz_sub(t, one, n); mpz_mod(t, t, pp); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, p); mpz_invert(hp, t, p);
    mpz_sub(t, one, n); mpz_mod(t, t, qq); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, q); mpz_invert(hq, t, q);
    mpz_invert(pinv, p, q);
    mpz_mod(t, c, pp); mpz_sub_ui(mp, p, 1); mpz_powm(t, t, mp, pp); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, p);
    mpz_mul(t, t, hp); mpz_mod(mp, t, p);
    mpz_mod(t, c, qq); mpz_sub_ui(mq, q, 1); mpz_powm(t, t, mq, qq); mpz_sub_ui(t, t, 1); mpz_tdiv_q(t, t, q);
    mpz_mul(t, t, hq); mpz_mod(mq, t, q);
    mpz_sub(t, mq, mp); mpz_mod(t, t, q); mpz_mul(t, t, pinv); mpz_mod(t, t, q);
    mpz_mul(t, t, p); mpz_add(t, t, mp);
    expo(m_out, k, t);
    mpz_clear(p); mpz_clear(q); mpz_clear(n); mpz_clear(nn); mpz_clear(m); mpz_clear(r); mpz_clear(c); mpz_clear(t);
    mpz_clear(pp); mpz_clear(qq); mpz_clear(hp); mpz_clear(hq); mpz_clear(mp); mpz_clear(mq); mpz_clear(pinv); mpz_clear(one);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CPU cost model of ONE offline-signing unit: the reference's big-integer operation list
 * (SURVEY.md section 8a "Per-unit totals", counted from src/utilities/mta/{mod,range_proofs}.rs,
 * src/utilities/zk_pdl_with_slack/mod.rs and sign/rounds.rs INCLUDING the reference's redundant
 * work: each MessageB::b re-verifies the three AliceProofs, Round5 re-verifies the party's own PDL
 * proof, Paillier decrypt recomputes its CRT constants per call) executed with GMP mpz_powm /
 * mpz_invert on random operands of the reference's sizes.  secp256k1 and SHA-256 work (<1 % of
 * the unit) is not included.  Used only as bench.py's CPU baseline for phases/s. */
typedef struct { size_t lo, hi; uint64_t seed; } oplist_job;

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return *s; }
static void rnd_mpz(mpz_t z, int bits, uint64_t* s, int odd_top) {
    uint32_t buf[160];
    int limbs = (bits + 31) / 32;
    for (int i = 0; i < limbs; i++) buf[i] = (uint32_t)(lcg(s) >> 32);
    if (bits & 31) buf[limbs - 1] &= (1u << (bits & 31)) - 1;
    if (odd_top) { buf[0] |= 1; buf[limbs - 1] |= 1u << ((bits - 1) & 31); }
    imp(z, buf, limbs);
}
static void powm_n(mpz_t r, mpz_t b, mpz_t e, mpz_t m, int mod_bits, int exp_bits, uint64_t* s, int times) {
    for (int i = 0; i < times; i++) {
        rnd_mpz(b, mod_bits - 1, s, 0);
        rnd_mpz(e, exp_bits, s, 0);
        mpz_powm(r, b, e, m);
    }
}
static void inv_n(mpz_t r, mpz_t b, mpz_t m, int mod_bits, uint64_t* s, int times) {
    for (int i = 0; i < times; i++) { rnd_mpz(b, mod_bits - 1, s, 0); mpz_invert(r, b, m); }
}
static void* oplist_worker(void* arg) {
    oplist_job* j = (oplist_job*)arg;
    mpz_t nt, n, nn, pp, p, b, e, r;
    mpz_init(nt); mpz_init(n); mpz_init(nn); mpz_init(pp); mpz_init(p); mpz_init(b); mpz_init(e); mpz_init(r);
    uint64_t s = j->seed + 0x9e3779b97f4a7c15ULL * (j->lo + 1);
    rnd_mpz(nt, 2048, &s, 1); rnd_mpz(n, 2048, &s, 1); mpz_mul(nn, n, n);
    rnd_mpz(p, 1024, &s, 1); mpz_mul(pp, p, p);
    for (size_t u = j->lo; u < j->hi; u++) {
        /* class A: N_tilde */
        powm_n(r, b, e, nt, 2048, 256, &s, 3 + 6 + 1 + 2);   /* a2 h1^a, a3 z^e, a7, a8 z^-e */
        powm_n(r, b, e, nt, 2048, 2304, &s, 3 + 1);          /* h2^ro */
        powm_n(r, b, e, nt, 2048, 768, &s, 3 + 6 + 1 + 2);   /* h1^alpha, h1^s1 */
        powm_n(r, b, e, nt, 2048, 2816, &s, 3 + 6 + 1 + 2);  /* h2^gamma, h2^s2/s3 */
        inv_n(r, b, nt, 2048, &s, 6 + 2);
        /* class B: N */
        powm_n(r, b, e, n, 2048, 256, &s, 3 + 1);
        /* class C: N^2 */
        powm_n(r, b, e, nn, 4096, 2048, &s, 1 + 3 + 6 + 2 + 1 + 2);  /* r^N, beta^N, s^N, r'^N, beta^N, s2^N */
        powm_n(r, b, e, nn, 4096, 256, &s, 6 + 2 + 2);              /* c^e, c_a^b, c^-e */
        powm_n(r, b, e, nn, 4096, 768, &s, 1 + 2);                  /* (N+1)^alpha, (N+1)^s1 */
        inv_n(r, b, nn, 4096, &s, 6 + 2);
        /* class D: Paillier decrypt x2 (CRT, constants recomputed per call) */
        powm_n(r, b, e, pp, 2048, 1024, &s, 4);
        inv_n(r, b, p, 1024, &s, 6);
    }
    mpz_clear(nt); mpz_clear(n); mpz_clear(nn); mpz_clear(pp); mpz_clear(p); mpz_clear(b); mpz_clear(e); mpz_clear(r);
    return NULL;
}
int oracle_unit_oplist(size_t units, int nthreads, uint64_t seed) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > units) nthreads = (int)(units ? units : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    oplist_job* jobs = (oplist_job*)malloc(sizeof(oplist_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (oplist_job){units * t / nthreads, units * (t + 1) / nthreads, seed};
        if (nthreads == 1) oplist_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, oplist_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
