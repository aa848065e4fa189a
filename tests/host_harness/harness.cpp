// TEST INFRASTRUCTURE: compiles the product's single-thread "glue" device code (secp256k1, SHA-256,
// plain-integer helpers, sigma proofs, Paillier CRT tail, per-key setup) for the HOST with shims
// for the CUDA qualifiers, so that the CPU test-suite can check it against the oracle without a
// GPU.  Nothing here is reachable from the product library.
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __global__
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, int n) { n &= 31; return n ? (lo >> n) | (hi << (32 - n)) : lo; }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
struct Dim3 { unsigned x, y, z; };
static Dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0};
#include "gg20_rounds.cuh"
#include "lindell17_kernels.cuh"
#include "gg18_kernels.cuh"
using namespace tecdsa;

// run a one-thread-per-element kernel for elements 0..count-1
template <class F> static void each_thread(int count, F f) {
    blockIdx.x = 0; blockDim.x = (unsigned)count;
    for (int t = 0; t < count; t++) { threadIdx.x = (unsigned)t; f(); }
    threadIdx.x = 0; blockDim.x = 1;
}

static uint32_t g_host_fb[2 * secp::FBP_WINDOWS * secp::FBP_DIGITS * 16];
extern "C" {
void h_init() {
    blockDim.x = 2 * secp::FBP_WINDOWS; blockIdx.x = 0;
    for (int t = 0; t < 2 * secp::FBP_WINDOWS; t++) { threadIdx.x = t; secp::fb_points_build(g_host_fb); }
    secp::g_fb_points = g_host_fb;
    threadIdx.x = 0; blockDim.x = 1;
}
void h_mul_fixed(uint32_t* out16, int which, const uint32_t* k8) { affine_store(out16, jac_to_affine(secp::jac_mul_fixed(which, u256_load(k8)))); }
void h_key_setup(uint32_t** tables, int rows) {
    blockDim.x = rows;
    for (int r = 0; r < rows; r++) { threadIdx.x = r; gg20_key_setup(tables, rows); }
}
void h_decrypt_finish(uint32_t* m64, uint32_t** tables, uint32_t row, const uint32_t* dp, const uint32_t* dq) {
    Arena A; memset(&A, 0, sizeof(A));
    for (int t = 0; t < KT_COUNT; t++) A.key[t] = tables[t];
    decrypt_finish(m64, A, row, dp, dq);
}
void h_crt_combine(uint32_t* x128, uint32_t** tables, uint32_t row, const uint32_t* yp, const uint32_t* yq) {
    Arena A; memset(&A, 0, sizeof(A));
    for (int t = 0; t < KT_COUNT; t++) A.key[t] = tables[t];
    crt_combine(x128, A, row, yp, yq);
}
void h_sc_from_limbs(uint32_t* out8, const uint32_t* x, int n) { U256 r = sc_from_limbs(x, n); memcpy(out8, r.v, 32); }
void h_sc_mul(uint32_t* out8, const uint32_t* a, const uint32_t* b) { U256 r = sc_mul(u256_load(a), u256_load(b)); memcpy(out8, r.v, 32); }
void h_sc_inv(uint32_t* out8, const uint32_t* a) { U256 r = sc_inv(u256_load(a)); memcpy(out8, r.v, 32); }
void h_fe_mul(uint32_t* out8, const uint32_t* a, const uint32_t* b) { U256 r = fe_mul(u256_load(a), u256_load(b)); memcpy(out8, r.v, 32); }
void h_pt_mul(uint32_t* out16, const uint32_t* p16, const uint32_t* k8) { affine_store(out16, pt_mul(affine_load(p16), u256_load(k8))); }
void h_pt_add(uint32_t* out16, const uint32_t* a16, const uint32_t* b16) { affine_store(out16, pt_add_aff(affine_load(a16), affine_load(b16))); }
void h_lagrange2(uint32_t* out8, uint32_t own, uint32_t peer) { U256 r = lagrange2(own, peer); memcpy(out8, r.v, 32); }
void h_alice_hash(uint32_t* e8, const uint32_t* N, const uint32_t* c, const uint32_t* z, const uint32_t* u, const uint32_t* w) { alice_hash(e8, N, c, z, u, w); }
void h_pdl_hash(uint32_t* e8, const uint32_t* G16, const uint32_t* Q16, const uint32_t* c, const uint32_t* z, const uint32_t* u1_16,
                const uint32_t* u2, const uint32_t* u3) {
    pdl_hash(e8, affine_load(G16), affine_load(Q16), c, z, affine_load(u1_16), u2, u3);
}
void h_hash_commit(uint32_t* out8, const uint32_t* p16, const uint32_t* blind8) { hash_commit_point(out8, affine_load(p16), blind8); }
void h_dlog_prove(uint32_t* out40, const uint32_t* sk8, const uint32_t* nonce8) { dlog_prove(out40, u256_load(sk8), u256_load(nonce8)); }
int h_dlog_verify(const uint32_t* in40) { return dlog_verify(in40) ? 1 : 0; }
int h_pedersen_verify(const uint32_t* ped64, const uint32_t* com16) { return pedersen_verify(ped64, affine_load(com16)) ? 1 : 0; }
int h_heg_verify(const uint32_t* heg48, const uint32_t* R16, const uint32_t* D16, const uint32_t* E16) {
    return heg_verify(heg48, affine_load(R16), affine_load(D16), affine_load(E16)) ? 1 : 0;
}
void h_fe_inv(uint32_t* out8, const uint32_t* a) { U256 r = secp::fe_inv(u256_load(a)); memcpy(out8, r.v, 32); }
// GLV split: m1, m2 (8 limbs each, < 2^128) and the two sign flags
void h_glv_split(uint32_t* m1, uint32_t* m2, int* neg, const uint32_t* k8) {
    U256 a, b; bool n1, n2;
    secp::glv_split(u256_load(k8), a, n1, b, n2);
    memcpy(m1, a.v, 32); memcpy(m2, b.v, 32); neg[0] = n1; neg[1] = n2;
}
void h_signed_windows5(signed char* digits, const uint32_t* m8) { secp::signed_windows5((int8_t*)digits, u256_load(m8)); }
// projective comparison against an affine point after scaling the Jacobian representation by z (a random non-trivial z)
int h_jac_eq_affine(const uint32_t* p16, const uint32_t* z8, const uint32_t* q16) {
    Affine P = affine_load(p16), Qp = affine_load(q16);
    Jac J = jac_from_affine(P);
    if (!P.inf) { U256 z = u256_load(z8), zz = fe_sqr(z); J.x = fe_mul(J.x, zz); J.y = fe_mul(J.y, fe_mul(zz, z)); J.z = z; }
    Jac K = jac_from_affine(Qp);
    return (secp::jac_eq_affine(J, Qp) ? 1 : 0) | (secp::jac_eq(J, K) ? 2 : 0);
}
void h_to_affine3(uint32_t* out48, const uint32_t* p16, const uint32_t* q16, const uint32_t* r16, const uint32_t* k8) {
    // three multiples k*P, k*Q, k*R normalised with one inversion
    Affine a, b, c;
    U256 k = u256_load(k8);
    secp::jac_to_affine3(a, b, c, jac_mul(jac_from_affine(affine_load(p16)), k), jac_mul(jac_from_affine(affine_load(q16)), k), jac_mul(jac_from_affine(affine_load(r16)), k));
    affine_store(out48, a); affine_store(out48 + 16, b); affine_store(out48 + 32, c);
}
void h_mul_add(uint32_t* d, int nd, const uint32_t* a, int na, const uint32_t* b, int nb, const uint32_t* c, int nc) { st::mul_add(d, nd, a, na, b, nb, c, nc); }

// ---- Lindell-2017 / zk_pdl / GG18 per-element kernels (csrc/lindell17_kernels.cuh, gg18_kernels.cuh) ----------------------------
typedef const uint32_t* CU;
typedef uint32_t* MU;
void h_l17_p2_pre(CU n_tab, CU key_idx, CU x2, CU k2, CU R1, CU msg, CU rho16, MU v8, MU lin128, uint8_t* st, int count) {
    each_thread(count, [&] { l17::k_l17_p2_pre(n_tab, key_idx, x2, k2, R1, msg, rho16, v8, lin128, st, count); });
}
void h_l17_p1_post(CU s_tag64, CU k1, CU R2, MU r, MU s, uint8_t* recid, uint8_t* st, int count) {
    each_thread(count, [&] { l17::k_l17_p1_post(s_tag64, k1, R2, r, s, recid, st, count); });
}
void h_l17_verify(CU r, CU s, CU pub, CU msg, uint8_t* st, int count) { each_thread(count, [&] { l17::k_l17_verify(r, s, pub, msg, st, count); }); }
void h_l17_eph_create(CU k, CU nonce, CU b1, CU b2, MU pub, MU c, MU proof, MU com1, MU com2, int count) {
    each_thread(count, [&] { l17::k_l17_eph_create(k, nonce, b1, b2, pub, c, proof, com1, com2, count); });
}
void h_l17_eph_verify(CU pub, CU c, CU proof, CU b1, CU b2, CU com1, CU com2, uint8_t* st, int count) {
    each_thread(count, [&] { l17::k_l17_eph_verify(pub, c, proof, b1, b2, com1, com2, st, count); });
}
void h_zkpdl_v1_pre(CU n_tab, CU key_idx, CU Q, CU a, CU b, CU blind, MU lin, MU ctt, MU qtag, uint8_t* st, int count) {
    each_thread(count, [&] { l17::k_zkpdl_v1_pre(n_tab, key_idx, Q, a, b, blind, lin, ctt, qtag, st, count); });
}
void h_zkpdl_p1_post(CU alpha, CU blind, MU chat, MU qhat, uint8_t* st, int count) { each_thread(count, [&] { l17::k_zkpdl_p1_post(alpha, blind, chat, qhat, st, count); }); }
void h_zkpdl_p2(CU x1, CU alpha, CU ctt, CU a, CU b, CU blind, uint8_t* st, int count) { each_thread(count, [&] { l17::k_zkpdl_p2(x1, alpha, ctt, a, b, blind, st, count); }); }
void h_zkpdl_finalize(CU chat, CU qhat, CU blind, CU qtag, uint8_t* st, int count) { each_thread(count, [&] { l17::k_zkpdl_finalize(chat, qhat, blind, qtag, st, count); }); }
void h_gg18_phase4(int parties, CU dinv, CU pk, CU gg, CU blind, CU com, MU R, uint8_t* st, int count) {
    each_thread(count, [&] { gg18::k_gg18_phase4(parties, dinv, pk, gg, blind, com, R, st, count); });
}
void h_gg18_phase5a(CU R, CU s, CU l, CU rho, CU blind, CU hs1, CU hs2, CU dn, MU com, MU vab, MU heg, MU dlog, uint8_t* st, int count) {
    each_thread(count, [&] { gg18::k_gg18_phase5a(R, s, l, rho, blind, hs1, hs2, dn, com, vab, heg, dlog, st, count); });
}
void h_gg18_phase5c(int parties, CU R, CU y, CU msg, CU rho, CU l, CU blind2, CU com, CU vab, CU blind1, CU heg, CU dlog, MU com2, MU ut, uint8_t* st, int count) {
    each_thread(count, [&] { gg18::k_gg18_phase5c(parties, R, y, msg, rho, l, blind2, com, vab, blind1, heg, dlog, com2, ut, st, count); });
}
void h_gg18_phase5d(int parties, CU ut, CU blind2, CU com2, CU vab, uint8_t* st, int count) {
    each_thread(count, [&] { gg18::k_gg18_phase5d(parties, ut, blind2, com2, vab, st, count); });
}
void h_gg18_local_sig(CU msg, CU R, CU k, CU sigma, MU s, int count) { each_thread(count, [&] { gg18::k_gg18_local_sig(msg, R, k, sigma, s, count); }); }
void h_gg18_output(int parties, CU R, CU y, CU msg, CU s, MU sr, MU ss, uint8_t* recid, uint8_t* st, int count) {
    each_thread(count, [&] { gg18::k_gg18_output(parties, R, y, msg, s, sr, ss, recid, st, count); });
}
}
