/* C99 consumer of include/tecdsa_b200.h: proves that the header is plain C, that the shared library links without any
 * Python / torch in the process, and (on a GPU box) that a call through the C ABI computes the right thing.
 *   gcc -std=c99 -Wall -Wextra -Werror -I include tests/c/abi_smoke.c -L multi-party-ecdsa_b200 -ltecdsa_b200 -o abi_smoke
 * Exit codes: 0 = all checks passed on a GPU; 77 = the library loaded and refused to work without a CUDA device (the product
 * has no CPU fallback); anything else = failure. */
#include "tecdsa_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define K 64

static int check(int rc, const char* what) {
    if (rc != 0) { fprintf(stderr, "%s failed: rc=%d (%s)\n", what, rc, tecdsa_last_error()); return 1; }
    return 0;
}

int main(void) {
    tecdsa_ctx* ctx = NULL;
    int rc = tecdsa_ctx_create(&ctx, 0, NULL);
    if (rc != 0) {
        printf("no CUDA device: tecdsa_ctx_create -> %d (%s)\n", rc, tecdsa_last_error());
        return rc == TECDSA_E_CUDA || rc == TECDSA_E_ARG ? 77 : 1;
    }
    /* BigInt::mod_pow known answers: 4^13 mod 497 = 445; 2^(2^11) mod (2^2047 + 5) checked through x^2 chains below */
    enum { N = 3 };
    uint32_t* base = calloc(N * K, 4), *exp = calloc(N * K, 4), *mod = calloc(N * K, 4), *out = calloc(N * K, 4);
    uint8_t status[N];
    if (!base || !exp || !mod || !out) return 1;
    base[0] = 4; exp[0] = 13; mod[0] = 497;
    base[K] = 3; exp[K] = 0; mod[K] = 1000003;                        /* x^0 = 1 */
    base[2 * K] = 7; exp[2 * K] = 5; mod[2 * K] = 4;                  /* even modulus -> TECDSA_ST_EVEN_MODULUS, zero output */
    if (check(tecdsa_modexp_batch(ctx, 2048, K, base, exp, mod, NULL, 0, out, status, N, TECDSA_HOST), "tecdsa_modexp_batch")) return 1;
    if (out[0] != 445 || out[K] != 1 || status[0] != TECDSA_ST_OK || status[1] != TECDSA_ST_OK || status[2] != TECDSA_ST_EVEN_MODULUS || out[2 * K] != 0) {
        fprintf(stderr, "modexp known answers wrong: %u %u status %d %d %d\n", out[0], out[K], status[0], status[1], status[2]);
        return 1;
    }
    for (int i = 1; i < K; i++) if (out[i] != 0 || out[K + i] != 0) { fprintf(stderr, "high limbs not zero\n"); return 1; }
    /* Scalar<Secp256k1>: (q - 1) * (q - 1) = 1 mod q;  Point: G + (-G) = identity, 1*G = G */
    static const uint32_t q_minus_1[8] = {0xD0364140u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    uint32_t prod[8];
    if (check(tecdsa_secp_scalar_mul_batch(ctx, q_minus_1, q_minus_1, prod, 1, TECDSA_HOST), "tecdsa_secp_scalar_mul_batch")) return 1;
    if (prod[0] != 1) { fprintf(stderr, "(q-1)^2 mod q != 1\n"); return 1; }
    for (int i = 1; i < 8; i++) if (prod[i]) { fprintf(stderr, "(q-1)^2 mod q != 1\n"); return 1; }
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, G[16], mG[16], sum[16];
    if (check(tecdsa_secp_mul_batch(ctx, NULL, one, G, 1, TECDSA_HOST), "tecdsa_secp_mul_batch")) return 1;
    if (check(tecdsa_secp_mul_batch(ctx, NULL, q_minus_1, mG, 1, TECDSA_HOST), "tecdsa_secp_mul_batch")) return 1;
    if (G[0] != 0x16F81798u || G[8] != 0xFB10D4B8u) { fprintf(stderr, "1*G is not the generator\n"); return 1; }
    if (check(tecdsa_secp_add_batch(ctx, G, mG, sum, 1, TECDSA_HOST), "tecdsa_secp_add_batch")) return 1;
    for (int i = 0; i < 16; i++) if (sum[i]) { fprintf(stderr, "G + (q-1)G is not the identity\n"); return 1; }
    /* SHA-256("abc") */
    static const uint8_t abc[3] = {'a', 'b', 'c'};
    uint64_t offs[2] = {0, 3};
    uint8_t dg[32];
    if (check(tecdsa_sha256_batch(ctx, abc, offs, dg, 1, TECDSA_HOST), "tecdsa_sha256_batch")) return 1;
    if (dg[0] != 0xba || dg[1] != 0x78 || dg[31] != 0xad) { fprintf(stderr, "SHA-256(abc) wrong\n"); return 1; }
    uint64_t macs = 0;
    if (check(tecdsa_ctx_work(ctx, &macs, 0), "tecdsa_ctx_work")) return 1;
    printf("abi_smoke ok: modexp, scalar, point and hash known answers through the C ABI; %llu MAC32 executed, %llu kernels launched\n",
           (unsigned long long)macs, (unsigned long long)tecdsa_ctx_launch_count(ctx));
    free(base); free(exp); free(mod); free(out);
    return tecdsa_ctx_destroy(ctx);
}
