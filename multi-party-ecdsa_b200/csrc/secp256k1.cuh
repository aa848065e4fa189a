// secp256k1 field / scalar / point arithmetic for sm_100a, one thread per operation.
//
// Structure (round 2): field operations, point doubling and addition are __forceinline__ and work on registers; the units of
// non-inlined code are whole scalar multiplications (jac_mul: GLV split + signed fixed 5-bit windows over a table of 16 multiples;
// jac_mul_fixed: 8-bit windows over precomputed affine tables of G and base_point2) and the field inversion (fixed addition
// chain).  Verifiers compare a computed Jacobian point with a received affine point WITHOUT inverting (jac_eq_affine), provers
// normalise all their output points with one shared inversion (jac_to_affine2/3).
//
// Replaces curv-kzen `Scalar<Secp256k1>` / `Point<Secp256k1>` (libsecp256k1 underneath) at the
// call sites of the GG20 offline stage: /root/reference/src/protocols/multi_party_ecdsa/gg_2020/
// party_i.rs:559-563,627-630,682-686,784, src/utilities/mta/mod.rs:168-169 and
// src/utilities/zk_pdl_with_slack/mod.rs:85,138-142.  256-bit values are 8 little-endian
// uint32 limbs.  Field elements are kept fully reduced in [0,p); results are canonical, so
// any correct implementation is bit-exact with the reference's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#ifdef __CUDACC__
#include "bigint.cuh"      // IMAD.WIDE carry-chain helpers (mad_even / madc_odd_rshift)
#endif

namespace tecdsa {
namespace secp {

struct U256 { uint32_t v[8]; };

__device__ __constant__ const uint32_t P_LIMBS[8] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu,
                                                     0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__device__ __constant__ const uint32_t Q_LIMBS[8] = {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u,
                                                     0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
// 2^256 - q
__device__ __constant__ const uint32_t QC_LIMBS[5] = {0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 0x1u};
__device__ __constant__ const uint32_t GX_LIMBS[8] = {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu,
                                                      0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu};
__device__ __constant__ const uint32_t GY_LIMBS[8] = {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u,
                                                      0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u};
// curv base_point2 (party_i.rs:629,787,812): x = SHA256^3(compressed G)
__device__ __constant__ const uint32_t HX_LIMBS[8] = {0x0378B795u, 0xA8DC7BFAu, 0x5FF3CE66u, 0xDD142E4Bu,
                                                      0x4BA80116u, 0x34DD4521u, 0xE3A7326Au, 0x08D13221u};
__device__ __constant__ const uint32_t HY_LIMBS[8] = {0xF7C2BE88u, 0x8217E9F7u, 0xDF0DF07Au, 0x807BCBA1u,
                                                      0xBD565EA2u, 0x0848D50Du, 0x77614B5Cu, 0x5D41AC14u};

// GLV endomorphism (x, y) -> (BETA x, y) = LAMBDA * (x, y) and the lattice constants of the scalar split k = k1 + k2 LAMBDA
// with |k1|, |k2| < 2^128 (derivation and exhaustive bound check: tests/test_glue_host.py::test_glv_split):
//   c1 = round(k G1 / 2^384), c2 = round(k G2 / 2^384), k2 = c1 (-b1) + c2 (-b2), k1 = k - k2 LAMBDA   (all mod q)
__device__ __constant__ const uint32_t BETA_LIMBS[8] = {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu};
__device__ __constant__ const uint32_t MINUS_LAMBDA_LIMBS[8] = {0xB51283CFu, 0xE0CFC810u, 0x8EC739C2u, 0xA880B9FCu, 0x77ED9BA4u, 0x5AD9E3FDu, 0x3FA3CF1Fu, 0xAC9C52B3u};
__device__ __constant__ const uint32_t GLV_G1_LIMBS[8] = {0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
__device__ __constant__ const uint32_t GLV_G2_LIMBS[8] = {0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
__device__ __constant__ const uint32_t GLV_MINUS_B1_LIMBS[8] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u, 0u, 0u, 0u, 0u};
__device__ __constant__ const uint32_t GLV_MINUS_B2_LIMBS[8] = {0x3DB1562Cu, 0xD765CDA8u, 0x0774346Du, 0x8A280AC5u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
// curv `VerifiableSS::map_share_to_new_params` [R] for two signers: lambda_own = x_peer / (x_peer - x_own) mod q with x = index + 1;
// the six (own, peer) pairs of n = 3 (row own * 3 + peer; the diagonal is unused)
__device__ __constant__ const uint32_t LAGRANGE2_LIMBS[9][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0},
    {2, 0, 0, 0, 0, 0, 0, 0},
    {0x681B20A2u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu},
    {0xD0364140u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu},
    {0, 0, 0, 0, 0, 0, 0, 0},
    {3, 0, 0, 0, 0, 0, 0, 0},
    {0x681B20A0u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu},
    {0xD036413Fu, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu},
    {0, 0, 0, 0, 0, 0, 0, 0}};

__device__ __forceinline__ U256 u256_zero() { U256 r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
__device__ __forceinline__ U256 u256_one() { U256 r = u256_zero(); r.v[0] = 1; return r; }
__device__ __forceinline__ U256 u256_load(const uint32_t* p) { U256 r; for (int i = 0; i < 8; i++) r.v[i] = p[i]; return r; }
__device__ __forceinline__ void u256_store(uint32_t* p, const U256& a) { for (int i = 0; i < 8; i++) p[i] = a.v[i]; }
__device__ __forceinline__ bool u256_is_zero(const U256& a) { uint32_t x = 0; for (int i = 0; i < 8; i++) x |= a.v[i]; return x == 0; }
__device__ __forceinline__ bool u256_eq(const U256& a, const U256& b) { uint32_t x = 0; for (int i = 0; i < 8; i++) x |= a.v[i] ^ b.v[i]; return x == 0; }
// a >= b
__device__ __forceinline__ bool u256_ge(const U256& a, const uint32_t* b) {
    for (int i = 7; i >= 0; i--) { if (a.v[i] > b[i]) return true; if (a.v[i] < b[i]) return false; }
    return true;
}
// r = a + b, returns carry
__device__ __forceinline__ uint32_t u256_add(U256& r, const U256& a, const U256& b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
// r = a - b, returns borrow
__device__ __forceinline__ uint32_t u256_sub(U256& r, const U256& a, const uint32_t* b) {
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - b[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
}
// 8x8 -> 16 limbs.  Device: operand scanning on the even/odd accumulator sets of bigint.cuh, one
// IMAD.WIDE.U32(.X) per 32x32 product (64 per call).  Host (tests/host_harness only): portable
// product scanning with a 96-bit column accumulator.
__device__ __forceinline__ void mul_256(uint32_t (&t)[16], const U256& a, const U256& b) {
#ifdef __CUDA_ARCH__
    uint32_t E[10], O[10];
#pragma unroll
    for (int j = 0; j < 10; j++) { E[j] = 0; O[j] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        // row i: E is the even set, O the previous even set turning into the odd set
        E[0] = tecdsa::add_cc(E[0], O[1]);
        tecdsa::madc_odd_rshift<8>(O, a.v, b.v[i]);
        tecdsa::mad_even<8>(E, a.v, b.v[i]);
        t[i] = E[0];
        // row i+1: roles swapped
        O[0] = tecdsa::add_cc(O[0], E[1]);
        tecdsa::madc_odd_rshift<8>(E, a.v, b.v[i + 1]);
        tecdsa::mad_even<8>(O, a.v, b.v[i + 1]);
        t[i + 1] = O[0];
    }
    // O: even set with column 0 consumed; E: odd set
    t[8] = tecdsa::add_cc(O[1], E[0]);
#pragma unroll
    for (int j = 1; j < 8; j++) t[8 + j] = tecdsa::addc_cc(O[j + 1], E[j]);
    (void)tecdsa::addc(0, 0);
#else
    uint64_t acc = 0; uint32_t hi = 0;
    for (int k = 0; k < 15; k++) {
        for (int i = 0; i < 8; i++) {
            int j = k - i;
            if (j < 0 || j > 7) continue;
            uint64_t p = (uint64_t)a.v[i] * b.v[j];
            acc += p;
            hi += (acc < p);
        }
        t[k] = (uint32_t)acc;
        acc = (acc >> 32) | ((uint64_t)hi << 32);
        hi = 0;
    }
    t[15] = (uint32_t)acc;
#endif
}

// ------------------------------------------------------------------------------- field Fp
// fold a 512-bit value modulo p = 2^256 - 0x1000003D1
__device__ __forceinline__ U256 fe_reduce512(const uint32_t (&t)[16]) {
    // lo + hi * (2^32 + 977)
    uint32_t r[10];
    uint64_t c = 0;
    // r = lo + hi*977
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + (uint64_t)t[8 + i] * 977u; r[i] = (uint32_t)c; c >>= 32; }
    r[8] = (uint32_t)c; r[9] = 0;
    // r += hi << 32
    c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)r[i + 1] + t[8 + i]; r[i + 1] = (uint32_t)c; c >>= 32; }
    r[9] = (uint32_t)c;
    // fold r[8], r[9] (at most ~34 bits above 2^256) once more
    uint64_t top = ((uint64_t)r[9] << 32) | r[8];
    uint64_t m = top * 977ull;                  // top < 2^34 so no overflow
    U256 o;
    c = (uint64_t)r[0] + (uint32_t)m;
    o.v[0] = (uint32_t)c; c >>= 32;
    c += (uint64_t)r[1] + (m >> 32) + (uint32_t)top;
    o.v[1] = (uint32_t)c; c >>= 32;
    c += (uint64_t)r[2] + (top >> 32);
    o.v[2] = (uint32_t)c; c >>= 32;
#pragma unroll
    for (int i = 3; i < 8; i++) { c += r[i]; o.v[i] = (uint32_t)c; c >>= 32; }
    // c is 0/1: one more tiny fold, then a final conditional subtract
    if (c) {
        uint64_t d = (uint64_t)o.v[0] + 977u; o.v[0] = (uint32_t)d; d >>= 32;
        d += (uint64_t)o.v[1] + 1u; o.v[1] = (uint32_t)d; d >>= 32;
        for (int i = 2; i < 8 && d; i++) { d += o.v[i]; o.v[i] = (uint32_t)d; d >>= 32; }
    }
    if (u256_ge(o, P_LIMBS)) { U256 s; u256_sub(s, o, P_LIMBS); o = s; }
    return o;
}
__device__ __forceinline__ U256 fe_mul(const U256& a, const U256& b) { uint32_t t[16]; mul_256(t, a, b); return fe_reduce512(t); }
__device__ __forceinline__ U256 fe_sqr(const U256& a) { return fe_mul(a, a); }
// branch-free: the subtraction is always computed and selected by mask (data-dependent branches diverge inside a warp)
__device__ __forceinline__ U256 fe_add(const U256& a, const U256& b) {
    U256 r, d;
    const uint32_t c = u256_add(r, a, b);
    const uint32_t bw = u256_sub(d, r, P_LIMBS);
    const uint32_t take = (c | (bw ^ 1u)) ? 0xFFFFFFFFu : 0u;           // carry out, or r >= p
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (d.v[i] & take) | (r.v[i] & ~take);
    return r;
}
__device__ __forceinline__ U256 fe_sub(const U256& a, const U256& b) {
    U256 r;
    const uint32_t bw = u256_sub(r, a, b.v);
    const uint32_t m = bw ? 0xFFFFFFFFu : 0u;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)r.v[i] + (P_LIMBS[i] & m); r.v[i] = (uint32_t)c; c >>= 32; }
    return r;
}
__device__ __forceinline__ U256 fe_neg(const U256& a) { return fe_sub(u256_zero(), a); }      // 0 - 0 = 0: no special case
__device__ __forceinline__ U256 fe_dbl(const U256& a) { return fe_add(a, a); }
__device__ __forceinline__ U256 fe_sqr_n(U256 x, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) x = fe_sqr(x);
    return x;
}
// a^(p-2) through a fixed addition chain: p - 2 has runs of 223, 22, 1, 2 and 1 ones — 255 squarings and 15 multiplications
// (plain square-and-multiply costs ~250 multiplications because almost every bit of p - 2 is set)
static __device__ __noinline__ U256 fe_inv(const U256& a) {
    const U256 x2 = fe_mul(fe_sqr(a), a);
    const U256 x3 = fe_mul(fe_sqr(x2), a);
    const U256 x6 = fe_mul(fe_sqr_n(x3, 3), x3);
    const U256 x9 = fe_mul(fe_sqr_n(x6, 3), x3);
    const U256 x11 = fe_mul(fe_sqr_n(x9, 2), x2);
    const U256 x22 = fe_mul(fe_sqr_n(x11, 11), x11);
    const U256 x44 = fe_mul(fe_sqr_n(x22, 22), x22);
    const U256 x88 = fe_mul(fe_sqr_n(x44, 44), x44);
    const U256 x176 = fe_mul(fe_sqr_n(x88, 88), x88);
    const U256 x220 = fe_mul(fe_sqr_n(x176, 44), x44);
    const U256 x223 = fe_mul(fe_sqr_n(x220, 3), x3);
    U256 t = fe_mul(fe_sqr_n(x223, 23), x22);
    t = fe_mul(fe_sqr_n(t, 5), a);
    t = fe_mul(fe_sqr_n(t, 3), x2);
    return fe_mul(fe_sqr_n(t, 2), a);
}

// ------------------------------------------------------------------------------- scalars mod q
__device__ __forceinline__ U256 sc_reduce_once(const U256& a, uint32_t carry) {
    U256 r = a;
    if (carry || u256_ge(r, Q_LIMBS)) { U256 s; u256_sub(s, r, Q_LIMBS); r = s; }
    return r;
}
// x (16 limbs) mod q via hi * (2^256 - q) folding
__device__ __forceinline__ U256 sc_reduce512(const uint32_t (&t)[16]) {
    // first fold: lo + hi*c  (c = 129 bits) -> up to 8+5+1 = 14 limbs
    uint32_t r[14];
#pragma unroll
    for (int i = 0; i < 14; i++) r[i] = (i < 8) ? t[i] : 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 5; j++) { c += (uint64_t)t[8 + i] * QC_LIMBS[j] + r[i + j]; r[i + j] = (uint32_t)c; c >>= 32; }
        for (int k = i + 5; c && k < 14; k++) { c += r[k]; r[k] = (uint32_t)c; c >>= 32; }
    }
    // second fold: limbs 8..13 (6 limbs) * c -> at most 11 limbs
    uint32_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = (i < 8) ? r[i] : 0;
    for (int i = 0; i < 6; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 5; j++) { c += (uint64_t)r[8 + i] * QC_LIMBS[j] + s[i + j]; s[i + j] = (uint32_t)c; c >>= 32; }
        for (int k = i + 5; c && k < 12; k++) { c += s[k]; s[k] = (uint32_t)c; c >>= 32; }
    }
    // third fold: limbs 8..11 (tiny) * c
    uint32_t u[10];
#pragma unroll
    for (int i = 0; i < 10; i++) u[i] = (i < 8) ? s[i] : 0;
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 5; j++) { if (i + j < 10) { c += (uint64_t)s[8 + i] * QC_LIMBS[j] + u[i + j]; u[i + j] = (uint32_t)c; c >>= 32; } }
        for (int k = i + 5; c && k < 10; k++) { c += u[k]; u[k] = (uint32_t)c; c >>= 32; }
    }
    U256 o; for (int i = 0; i < 8; i++) o.v[i] = u[i];
    // u[8] is 0 or 1 here
    uint32_t carry = u[8];
    o = sc_reduce_once(o, carry);
    return sc_reduce_once(o, 0);
}
__device__ __forceinline__ U256 sc_mul(const U256& a, const U256& b) { uint32_t t[16]; mul_256(t, a, b); return sc_reduce512(t); }
__device__ __forceinline__ U256 sc_add(const U256& a, const U256& b) { U256 r; uint32_t c = u256_add(r, a, b); return sc_reduce_once(r, c); }
__device__ __forceinline__ U256 sc_neg(const U256& a) {
    if (u256_is_zero(a)) return a;
    U256 q = u256_load(Q_LIMBS), r; u256_sub(r, q, a.v); return r;
}
__device__ __forceinline__ U256 sc_sub(const U256& a, const U256& b) { return sc_add(a, sc_neg(b)); }
// `Scalar::from(&BigInt)`: any non-negative integer (n little-endian limbs) reduced mod q
__device__ __forceinline__ U256 sc_from_limbs(const uint32_t* x, int n) {
    U256 r = u256_zero();
    // Horner over 256-bit chunks from the top: r = r * 2^256 + chunk (mod q)
    int chunks = (n + 7) / 8;
    for (int ch = chunks - 1; ch >= 0; ch--) {
        uint32_t t[16];
        // r * 2^256 mod q == r * c mod q
        U256 cc = u256_zero();
        for (int j = 0; j < 5; j++) cc.v[j] = QC_LIMBS[j];
        mul_256(t, r, cc);
        U256 hi = sc_reduce512(t);
        U256 lo;
        for (int i = 0; i < 8; i++) { int k = ch * 8 + i; lo.v[i] = (k < n) ? x[k] : 0; }
        lo = sc_reduce_once(lo, 0);
        r = sc_add(hi, lo);
    }
    return r;
}
// a^(q-2)  (`Scalar::invert`, party_i.rs:639)
static __device__ __noinline__ U256 sc_inv(const U256& a) {
    U256 r = u256_one();
    for (int i = 255; i >= 0; i--) {
        r = sc_mul(r, r);
        uint32_t limb = Q_LIMBS[i >> 5];
        if (i < 32) limb = 0xD036413Fu;           // low limb of q-2
        if ((limb >> (i & 31)) & 1u) r = sc_mul(r, a);
    }
    return r;
}

// ------------------------------------------------------------------------------- points
struct Affine { U256 x, y; bool inf; };
struct Jac { U256 x, y, z; };     // z == 0 <=> identity

__device__ __forceinline__ Jac jac_identity() { Jac r; r.x = u256_one(); r.y = u256_one(); r.z = u256_zero(); return r; }
__device__ __forceinline__ bool jac_is_inf(const Jac& p) { return u256_is_zero(p.z); }
__device__ __forceinline__ Jac jac_from_affine(const Affine& a) {
    if (a.inf) return jac_identity();
    Jac r; r.x = a.x; r.y = a.y; r.z = u256_one(); return r;
}
__device__ __forceinline__ Affine affine_G() { Affine a; a.x = u256_load(GX_LIMBS); a.y = u256_load(GY_LIMBS); a.inf = false; return a; }
__device__ __forceinline__ Affine affine_H() { Affine a; a.x = u256_load(HX_LIMBS); a.y = u256_load(HY_LIMBS); a.inf = false; return a; }

// 2P, a = 0 (2M + 5S); the identity and points of order two (none on this curve) come out with z == 0
__device__ __forceinline__ Jac jac_dbl(const Jac& p) {
    U256 A = fe_sqr(p.x), B = fe_sqr(p.y), C = fe_sqr(B);
    U256 t = fe_add(p.x, B);
    U256 D = fe_dbl(fe_sub(fe_sub(fe_sqr(t), A), C));
    U256 E = fe_add(fe_dbl(A), A);
    U256 F = fe_sqr(E);
    Jac r;
    r.x = fe_sub(F, fe_dbl(D));
    U256 c8 = fe_dbl(fe_dbl(fe_dbl(C)));
    r.y = fe_sub(fe_mul(E, fe_sub(D, r.x)), c8);
    r.z = fe_dbl(fe_mul(p.y, p.z));                 // z == 0 stays 0
    return r;
}
// P + P reached through the addition formulas (equal inputs): rare, so it goes through a non-inlined copy of the doubling
static __device__ __noinline__ Jac jac_dbl_rare(const Jac& p) { return jac_dbl(p); }
// general Jacobian + Jacobian (12M + 4S), all special cases handled
__device__ __forceinline__ Jac jac_add(const Jac& p, const Jac& q) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    U256 z1z1 = fe_sqr(p.z), z2z2 = fe_sqr(q.z);
    U256 u1 = fe_mul(p.x, z2z2), u2 = fe_mul(q.x, z1z1);
    U256 s1 = fe_mul(fe_mul(p.y, q.z), z2z2), s2 = fe_mul(fe_mul(q.y, p.z), z1z1);
    U256 h = fe_sub(u2, u1), r = fe_sub(s2, s1);
    if (u256_is_zero(h)) {
        if (u256_is_zero(r)) return jac_dbl_rare(p);
        return jac_identity();
    }
    U256 hh = fe_sqr(h), hhh = fe_mul(hh, h), v = fe_mul(u1, hh);
    Jac o;
    o.x = fe_sub(fe_sub(fe_sqr(r), hhh), fe_dbl(v));
    o.y = fe_sub(fe_mul(r, fe_sub(v, o.x)), fe_mul(s1, hhh));
    o.z = fe_mul(fe_mul(p.z, q.z), h);
    return o;
}
__device__ __forceinline__ Jac jac_neg(const Jac& p) { Jac r = p; r.y = fe_neg(p.y); return r; }
// Jacobian + affine (8M + 3S), all special cases handled
__device__ __forceinline__ Jac jac_madd(const Jac& p, const Affine& a) {
    if (a.inf) return p;
    if (jac_is_inf(p)) return jac_from_affine(a);
    U256 z1z1 = fe_sqr(p.z);
    U256 u2 = fe_mul(a.x, z1z1), s2 = fe_mul(fe_mul(a.y, p.z), z1z1);
    U256 h = fe_sub(u2, p.x), r = fe_sub(s2, p.y);
    if (u256_is_zero(h)) {
        if (u256_is_zero(r)) return jac_dbl_rare(p);
        return jac_identity();
    }
    U256 hh = fe_sqr(h), hhh = fe_mul(hh, h), v = fe_mul(p.x, hh);
    Jac o;
    o.x = fe_sub(fe_sub(fe_sqr(r), hhh), fe_dbl(v));
    o.y = fe_sub(fe_mul(r, fe_sub(v, o.x)), fe_mul(p.y, hhh));
    o.z = fe_mul(p.z, h);
    return o;
}
__device__ __forceinline__ Jac jac_add_affine(const Jac& p, const Affine& a) { return jac_madd(p, a); }
static __device__ __noinline__ Jac jac_add_noinline(const Jac& p, const Jac& q) { return jac_add(p, q); }

// (X : Y : Z) == (x, y) without an inversion: X == x Z^2 and Y == y Z^3 (4M + 1S).  Verifiers compare the point they computed
// with the point they were sent this way.
__device__ __forceinline__ bool jac_eq_affine(const Jac& p, const Affine& a) {
    if (jac_is_inf(p) || a.inf) return jac_is_inf(p) && a.inf;
    const U256 zz = fe_sqr(p.z);
    return u256_eq(p.x, fe_mul(a.x, zz)) && u256_eq(p.y, fe_mul(a.y, fe_mul(zz, p.z)));
}
// two Jacobian points equal (cross-multiplied)
__device__ __forceinline__ bool jac_eq(const Jac& p, const Jac& q) {
    if (jac_is_inf(p) || jac_is_inf(q)) return jac_is_inf(p) && jac_is_inf(q);
    const U256 z1 = fe_sqr(p.z), z2 = fe_sqr(q.z);
    return u256_eq(fe_mul(p.x, z2), fe_mul(q.x, z1)) && u256_eq(fe_mul(p.y, fe_mul(z2, q.z)), fe_mul(q.y, fe_mul(z1, p.z)));
}
__device__ __forceinline__ Affine jac_scale(const Jac& p, const U256& zi) {
    Affine a;
    const U256 zi2 = fe_sqr(zi);
    a.x = fe_mul(p.x, zi2); a.y = fe_mul(p.y, fe_mul(zi2, zi)); a.inf = false;
    return a;
}
__device__ __forceinline__ Affine affine_inf() { Affine a; a.inf = true; a.x = u256_zero(); a.y = u256_zero(); return a; }
static __device__ __noinline__ Affine jac_to_affine(const Jac& p) {
    if (jac_is_inf(p)) return affine_inf();
    return jac_scale(p, fe_inv(p.z));
}
// two / three points with ONE shared inversion (Montgomery's trick); an identity among them takes the place of z = 1
static __device__ __noinline__ void jac_to_affine2(Affine& a, Affine& b, const Jac& p, const Jac& q) {
    const bool ip = jac_is_inf(p), iq = jac_is_inf(q);
    const U256 zp = ip ? u256_one() : p.z, zq = iq ? u256_one() : q.z;
    const U256 inv = fe_inv(fe_mul(zp, zq));
    a = ip ? affine_inf() : jac_scale(p, fe_mul(inv, zq));
    b = iq ? affine_inf() : jac_scale(q, fe_mul(inv, zp));
}
static __device__ __noinline__ void jac_to_affine3(Affine& a, Affine& b, Affine& c, const Jac& p, const Jac& q, const Jac& r) {
    const bool ip = jac_is_inf(p), iq = jac_is_inf(q), ir = jac_is_inf(r);
    const U256 zp = ip ? u256_one() : p.z, zq = iq ? u256_one() : q.z, zr = ir ? u256_one() : r.z;
    const U256 pq = fe_mul(zp, zq);
    const U256 inv = fe_inv(fe_mul(pq, zr));
    const U256 inv_pq = fe_mul(inv, zr);                    // (zp zq)^-1
    a = ip ? affine_inf() : jac_scale(p, fe_mul(inv_pq, zq));
    b = iq ? affine_inf() : jac_scale(q, fe_mul(inv_pq, zp));
    c = ir ? affine_inf() : jac_scale(r, fe_mul(inv, pq));
}

// ---- `Point * Scalar` for an arbitrary point: GLV split + signed fixed windows --------------------------------------------------
// high 128 bits (limbs 12..15) of a 512-bit product, rounded at bit 383: round(k g / 2^384)
__device__ __forceinline__ U256 mul_shift384(const U256& k, const uint32_t* g) {
    uint32_t t[16];
    mul_256(t, k, u256_load(g));
    U256 r = u256_zero();
    uint64_t c = (t[11] >> 31) & 1u;
    for (int i = 0; i < 4; i++) { c += t[12 + i]; r.v[i] = (uint32_t)c; c >>= 32; }
    r.v[4] = (uint32_t)c;
    return r;
}
// k = k1 + k2 LAMBDA (mod q) with k1 = s1 * m1, k2 = s2 * m2, m1, m2 < 2^128, s = +-1 (neg1/neg2 set for -1)
__device__ __forceinline__ void glv_split(const U256& k, U256& m1, bool& neg1, U256& m2, bool& neg2) {
    const U256 c1 = mul_shift384(k, GLV_G1_LIMBS), c2 = mul_shift384(k, GLV_G2_LIMBS);
    U256 k2 = sc_add(sc_mul(c1, u256_load(GLV_MINUS_B1_LIMBS)), sc_mul(c2, u256_load(GLV_MINUS_B2_LIMBS)));
    U256 k1 = sc_add(sc_mul(k2, u256_load(MINUS_LAMBDA_LIMBS)), k);
    // a value above 2^128 is the negative of a small one
    neg1 = (k1.v[4] | k1.v[5] | k1.v[6] | k1.v[7]) != 0;
    neg2 = (k2.v[4] | k2.v[5] | k2.v[6] | k2.v[7]) != 0;
    m1 = neg1 ? sc_neg(k1) : k1;
    m2 = neg2 ? sc_neg(k2) : k2;
}
// Signed fixed 5-bit windows of a value below 2^129: m = sum d[i] 32^i with d[i] in [-15, 16], i < 27.  FIXED windows, not a
// sparse (NAF) form: the 32 lanes of a warp run in lock step, so an addition costs the warp its full time whenever ANY lane has a
// non-zero digit — sparse digits buy nothing under SIMT, regular windows keep every lane doing useful work at every step.
static constexpr int SW_DIGITS = 27;
__device__ __forceinline__ void signed_windows5(int8_t* d, const U256& m) {
    uint32_t carry = 0;
#pragma unroll 1
    for (int i = 0; i < SW_DIGITS; i++) {
        const int bit = 5 * i, limb = bit >> 5, off = bit & 31;
        uint32_t v = limb < 5 ? m.v[limb] >> off : 0u;
        if (off > 27 && limb + 1 < 5) v |= m.v[limb + 1] << (32 - off);
        v = (v & 31u) + carry;
        carry = v > 16u ? 1u : 0u;
        d[i] = (int8_t)((int)v - (int)(carry << 5));
    }
}
// k * base.  One non-inlined unit; doubling and the table addition are inlined once each in the main loop.  GLV: k = k1 + k2 LAMBDA
// with 128-bit halves, so 130 doublings instead of 256; both halves share the table of 1..16 times the base (the second half
// through the endomorphism x -> BETA x).
static __device__ __noinline__ Jac jac_mul(const Jac& base, const U256& k) {
    const U256 kr = sc_reduce_once(k, 0);
    if (jac_is_inf(base) || u256_is_zero(kr)) return jac_identity();
    Jac tbl[16];
    tbl[0] = base;
    tbl[1] = jac_dbl_rare(base);
#pragma unroll 1
    for (int i = 2; i < 16; i++) tbl[i] = jac_add_noinline(tbl[i - 1], base);
    U256 m1, m2;
    bool neg1, neg2;
    glv_split(kr, m1, neg1, m2, neg2);
    int8_t d1[SW_DIGITS], d2[SW_DIGITS];
    signed_windows5(d1, m1);
    signed_windows5(d2, m2);
    const U256 beta = u256_load(BETA_LIMBS);
    Jac r = jac_identity();
#pragma unroll 1
    for (int i = SW_DIGITS - 1; i >= 0; i--) {
        if (i != SW_DIGITS - 1) {
#pragma unroll 1
            for (int s = 0; s < 5; s++) r = jac_dbl(r);
        }
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int dg = h ? d2[i] : d1[i];
            if (dg == 0) continue;
            Jac t = tbl[(dg > 0 ? dg : -dg) - 1];
            if (h) t.x = fe_mul(t.x, beta);
            if ((dg < 0) != (h ? neg2 : neg1)) t.y = fe_neg(t.y);
            r = jac_add(r, t);
        }
    }
    return r;
}
__device__ __forceinline__ Affine pt_mul(const Affine& p, const U256& k) { return jac_to_affine(jac_mul(jac_from_affine(p), k)); }

// ---- fixed-base tables for the two constant points of the protocol (the generator and curv's base_point2) ---------------------
// T[b][w][d-1] = d * 256^w * B_b, affine, w < 32, d = 1..255 (2 x 32 x 255 x 64 B = 1.04 MB, built once per device by
// fb_points_build; L2-resident).  k * B_b is then at most 32 mixed additions and no doubling.
static constexpr int FBP_WINDOWS = 32, FBP_DIGITS = 255;
static __device__ const uint32_t* g_fb_points = nullptr;      // set per translation unit, see tecdsa_internal_fb_points_set_*

static __device__ __noinline__ Jac jac_mul_fixed(int which, const U256& k) {
    const uint32_t* t = g_fb_points + (size_t)which * FBP_WINDOWS * FBP_DIGITS * 16;
    Jac r = jac_identity();
#pragma unroll 1
    for (int w = 0; w < FBP_WINDOWS; w++) {
        const uint32_t d = (k.v[w >> 2] >> ((w & 3) * 8)) & 255u;
        if (d) {
            Affine a;
            const uint32_t* e = t + ((size_t)w * FBP_DIGITS + (d - 1)) * 16;
            a.x = u256_load(e); a.y = u256_load(e + 8); a.inf = false;
            r = jac_madd(r, a);
        }
    }
    return r;
}
// one thread per (base, window): the 255 multiples of 256^w * B, normalised in groups with a shared inversion
static __global__ void fb_points_build(uint32_t* tables) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * FBP_WINDOWS) return;
    const int which = id / FBP_WINDOWS, w = id % FBP_WINDOWS;
    Jac b = jac_from_affine(which ? affine_H() : affine_G());
    for (int i = 0; i < 8 * w; i++) b = jac_dbl(b);
    uint32_t* out = tables + ((size_t)which * FBP_WINDOWS + w) * FBP_DIGITS * 16;
    Jac acc = b;
    for (int d = 1; d <= FBP_DIGITS; d += 3) {
        Jac p1 = acc; acc = jac_add_noinline(acc, b);
        Jac p2 = acc; acc = jac_add_noinline(acc, b);
        Jac p3 = acc; acc = jac_add_noinline(acc, b);
        Affine a1, a2, a3;
        jac_to_affine3(a1, a2, a3, p1, p2, p3);
        u256_store(out + (size_t)(d - 1) * 16, a1.x); u256_store(out + (size_t)(d - 1) * 16 + 8, a1.y);
        u256_store(out + (size_t)d * 16, a2.x); u256_store(out + (size_t)d * 16 + 8, a2.y);
        u256_store(out + (size_t)(d + 1) * 16, a3.x); u256_store(out + (size_t)(d + 1) * 16 + 8, a3.y);
    }
}
__device__ __forceinline__ bool affine_eq(const Affine& a, const Affine& b) {
    if (a.inf || b.inf) return a.inf && b.inf;
    return u256_eq(a.x, b.x) && u256_eq(a.y, b.y);
}
// y^2 == x^3 + 7 with canonical coordinates (x, y < p): curv's Point deserialisation rejects anything else, and a non-canonical
// coordinate would also enter transcript hashes as different bytes
__device__ __forceinline__ bool on_curve(const Affine& a) {
    if (a.inf) return true;
    if (u256_ge(a.x, P_LIMBS) || u256_ge(a.y, P_LIMBS)) return false;
    U256 seven = u256_zero(); seven.v[0] = 7;
    return u256_eq(fe_sqr(a.y), fe_add(fe_mul(fe_sqr(a.x), a.x), seven));
}
// points travel through the ABI as 64 bytes: x||y little-endian limbs (16 uint32); all-zero = identity
__device__ __forceinline__ Affine affine_load(const uint32_t* p) {
    Affine a; a.x = u256_load(p); a.y = u256_load(p + 8); a.inf = u256_is_zero(a.x) && u256_is_zero(a.y); return a;
}
__device__ __forceinline__ void affine_store(uint32_t* p, const Affine& a) {
    if (a.inf) { for (int i = 0; i < 16; i++) p[i] = 0; return; }
    u256_store(p, a.x); u256_store(p + 8, a.y);
}
// `Point::to_bytes(true)`: 33-byte SEC1 compressed, big-endian
__device__ __forceinline__ void affine_compress(uint8_t* out, const Affine& a) {
    out[0] = 2 + (a.y.v[0] & 1);
    for (int i = 0; i < 32; i++) out[1 + i] = (uint8_t)(a.x.v[7 - (i >> 2)] >> (8 * (3 - (i & 3))));
}

}  // namespace secp
}  // namespace tecdsa
