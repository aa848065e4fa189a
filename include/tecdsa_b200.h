/* tecdsa_b200 — C ABI of the B200 batched threshold-ECDSA arithmetic engine.
 *
 * The reference (ZenGo-X/multi-party-ecdsa @ 7d8bd41) has no FFI of its own: its seam is
 * the Rust trait surface of curv-kzen `BigInt`, `Scalar/Point<Secp256k1>`, kzen-paillier
 * `Paillier::*` and the in-tree proof structs.  Every entry point below is the BATCHED form
 * of one of those scalar calls and cites the reference call site it replaces (paths are
 * relative to /root/reference).  INTEGRATION.md shows the Rust `extern "C"` shim a maintainer
 * would add so that src/protocols/* links against this library.
 *
 * Conventions
 *   - Big integers are little-endian arrays of uint32_t limbs, operand-major:
 *     x[i*K .. i*K+K) is operand i (K = bits/32).  Rows must be 16-byte aligned.
 *     Byte strings (hash inputs, compressed points) are big-endian exactly as
 *     `BigInt::to_bytes()` / `Point::to_bytes(true)` produce them.
 *   - `mem` says where caller buffers live: TECDSA_HOST (library stages H2D/D2H on the
 *     context stream) or TECDSA_DEVICE (pointers are device pointers, nothing is copied).
 *   - Every call returns 0 on success, <0 on API/CUDA failure (tecdsa_last_error()).
 *     Per-element outcomes go to `status[count]` (TECDSA_ST_*); a failed element never
 *     aborts the batch and never panics, unlike the reference's assert!/unwrap() sites.
 *   - Calls are asynchronous on the context's stream for TECDSA_DEVICE, synchronous for
 *     TECDSA_HOST.  A context is not thread-safe; distinct contexts are independent.
 *   - No randomness is drawn inside the library: every sampled value of the reference
 *     (`BigInt::sample_below`, `Scalar::random`) is an explicit input.
 */
#ifndef TECDSA_B200_H
#define TECDSA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tecdsa_ctx tecdsa_ctx;

enum { TECDSA_HOST = 0, TECDSA_DEVICE = 1 };

/* return codes */
enum { TECDSA_OK = 0, TECDSA_E_ARG = -1, TECDSA_E_CUDA = -2, TECDSA_E_NOMEM = -3, TECDSA_E_UNSUPPORTED = -4 };

/* per-element status bytes; numbering follows the reference's failure points */
enum {
    TECDSA_ST_OK = 0,
    TECDSA_ST_EVEN_MODULUS = 1,   /* mod_pow on an even modulus: outside Montgomery domain  */
    TECDSA_ST_INVALID_KEY = 2,    /* Error::InvalidKey           src/utilities/mta/mod.rs:120,130,177 */
    TECDSA_ST_RANGE = 3,          /* s1 > q^3                    src/utilities/mta/range_proofs.rs:118 */
    TECDSA_ST_NOT_INVERTIBLE = 4, /* mod_inv -> None             range_proofs.rs:123-127,136-139; zk_pdl_with_slack/mod.rs:192 */
    TECDSA_ST_HASH_MISMATCH = 5,  /* e != self.e                 range_proofs.rs:151 */
    TECDSA_ST_PDL_VERIFY = 6,     /* ZkPdlWithSlackError::Verify zk_pdl_with_slack/mod.rs:177 */
    TECDSA_ST_PHASE5_BAD_SUM = 7, /* Error::Phase5BadSum         gg_2020/party_i.rs:774 */
    TECDSA_ST_PHASE6 = 8,         /* Error::Phase6Error          gg_2020/party_i.rs:846 */
    TECDSA_ST_INVALID_SIG = 9,    /* Error::InvalidSig           gg_2020/party_i.rs:908,934 */
    TECDSA_ST_PROOF = 10,         /* a curv sigma proof (DLog/Pedersen/HomoElGamal) failed to verify */
    TECDSA_ST_COMMITMENT = 11     /* "bad gamma_i decommit"      gg_2020/party_i.rs:650-674 */
};

/* ---- context ------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t to launch on, or NULL for a private non-blocking stream.      */
int tecdsa_ctx_create(tecdsa_ctx** ctx, int device, void* stream);
int tecdsa_ctx_destroy(tecdsa_ctx* ctx);
int tecdsa_ctx_sync(tecdsa_ctx* ctx);
const char* tecdsa_last_error(void);
/* Lane-group width (4, 8, 16 or 32 lanes per operand; 32 = one warp per operand) used for
 * `mod_bits`-wide moduli.  0 restores the tuned default. */
int tecdsa_ctx_set_tpi(tecdsa_ctx* ctx, int mod_bits, int tpi);
/* Device time (ms, CUDA events on the context stream) of the kernels of the last call and
 * how many kernels that call launched. */
int tecdsa_ctx_last_kernel_ms(tecdsa_ctx* ctx, float* ms, int* launches);
/* Total kernels launched through this context since creation. */
uint64_t tecdsa_ctx_launch_count(tecdsa_ctx* ctx);

/* ---- L0: big-integer arithmetic --------------------------------------------------------
 * out[i] = base[i] ^ exp[i] mod modulus[i]            (BigInt::mod_pow -> GMP mpz_powm;
 *   call sites src/utilities/mta/range_proofs.rs:52,54,57,86,122,129,130,135,141;
 *   src/utilities/zk_pdl_with_slack/mod.rs:189,193,196)
 * mod_bits in {1024, 2048, 4096}; operands are K = mod_bits/32 limbs; the modulus must be
 * odd (else TECDSA_ST_EVEN_MODULUS and a zero output); base may be any K-limb value (it is
 * reduced); exp is `exp_limbs` limbs (the batch's public maximum width, zero-padded).
 * If mod_idx != NULL, operand i uses modulus[mod_idx[i]] out of `n_mod` rows, else row i.  */
int tecdsa_modexp_batch(tecdsa_ctx* ctx, int mod_bits, int exp_limbs, const uint32_t* base, const uint32_t* exp,
                        const uint32_t* modulus, const uint32_t* mod_idx, size_t n_mod, uint32_t* out,
                        uint8_t* status, size_t count, int mem);

/* Saturation micro-benchmark of the integer multiply-add pipe (IMAD.WIDE.U32 with carry
 * chains shaped like the Montgomery rows): 32x32+64 MACs per second on this device.  This is
 * the roofline denominator for every kernel of this library. */
int tecdsa_imad_peak(tecdsa_ctx* ctx, double* mac32_per_s, float* ms);

#ifdef __cplusplus
}
#endif
#endif
