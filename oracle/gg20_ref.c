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
