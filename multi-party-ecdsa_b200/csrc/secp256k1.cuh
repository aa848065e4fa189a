// secp256k1 field / scalar / point arithmetic for sm_100a, one thread per operation.
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
__device__ __forceinline__ U256 fe_add(const U256& a, const U256& b) {
    U256 r; uint32_t c = u256_add(r, a, b);
    if (c || u256_ge(r, P_LIMBS)) { U256 s; u256_sub(s, r, P_LIMBS); r = s; }
    return r;
}
__device__ __forceinline__ U256 fe_sub(const U256& a, const U256& b) {
    U256 r; uint32_t bw = u256_sub(r, a, b.v);
    if (bw) { U256 p = u256_load(P_LIMBS); U256 s; u256_add(s, r, p); r = s; }
    return r;
}
__device__ __forceinline__ U256 fe_neg(const U256& a) { return u256_is_zero(a) ? a : fe_sub(u256_zero(), a); }
__device__ __forceinline__ U256 fe_dbl(const U256& a) { return fe_add(a, a); }
// a^(p-2)
static __device__ __noinline__ U256 fe_inv(const U256& a) {
    // p - 2 = 2^256 - 2^32 - 979: plain square-and-multiply over its bits (MSB first)
    U256 r = u256_one();
    for (int i = 255; i >= 0; i--) {
        r = fe_sqr(r);
        uint32_t limb = P_LIMBS[i >> 5];
        if (i < 32) limb = 0xFFFFFC2Du;           // low limb of p-2
        if ((limb >> (i & 31)) & 1u) r = fe_mul(r, a);
    }
    return r;
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

static __device__ __noinline__ Jac jac_dbl(const Jac& p) {
    if (jac_is_inf(p) || u256_is_zero(p.y)) return jac_identity();
    U256 A = fe_sqr(p.x), B = fe_sqr(p.y), C = fe_sqr(B);
    U256 t = fe_add(p.x, B);
    U256 D = fe_dbl(fe_sub(fe_sub(fe_sqr(t), A), C));
    U256 E = fe_add(fe_dbl(A), A);
    U256 F = fe_sqr(E);
    Jac r;
    r.x = fe_sub(F, fe_dbl(D));
    U256 c8 = fe_dbl(fe_dbl(fe_dbl(C)));
    r.y = fe_sub(fe_mul(E, fe_sub(D, r.x)), c8);
    r.z = fe_dbl(fe_mul(p.y, p.z));
    return r;
}
// general Jacobian + Jacobian, all special cases handled
static __device__ __noinline__ Jac jac_add(const Jac& p, const Jac& q) {
    if (jac_is_inf(p)) return q;
    if (jac_is_inf(q)) return p;
    U256 z1z1 = fe_sqr(p.z), z2z2 = fe_sqr(q.z);
    U256 u1 = fe_mul(p.x, z2z2), u2 = fe_mul(q.x, z1z1);
    U256 s1 = fe_mul(fe_mul(p.y, q.z), z2z2), s2 = fe_mul(fe_mul(q.y, p.z), z1z1);
    U256 h = fe_sub(u2, u1), r = fe_sub(s2, s1);
    if (u256_is_zero(h)) {
        if (u256_is_zero(r)) return jac_dbl(p);
        return jac_identity();
    }
    U256 hh = fe_sqr(h), hhh = fe_mul(hh, h), v = fe_mul(u1, hh);
    Jac o;
    o.x = fe_sub(fe_sub(fe_sqr(r), hhh), fe_dbl(v));
    o.y = fe_sub(fe_mul(r, fe_sub(v, o.x)), fe_mul(s1, hhh));
    o.z = fe_mul(fe_mul(p.z, q.z), h);
    return o;
}
__device__ __forceinline__ Jac jac_add_affine(const Jac& p, const Affine& a) { return jac_add(p, jac_from_affine(a)); }
__device__ __forceinline__ Jac jac_neg(const Jac& p) { Jac r = p; r.y = fe_neg(p.y); return r; }
static __device__ __noinline__ Affine jac_to_affine(const Jac& p) {
    Affine a;
    if (jac_is_inf(p)) { a.inf = true; a.x = u256_zero(); a.y = u256_zero(); return a; }
    U256 zi = fe_inv(p.z), zi2 = fe_sqr(zi);
    a.x = fe_mul(p.x, zi2); a.y = fe_mul(p.y, fe_mul(zi2, zi)); a.inf = false;
    return a;
}
// `Point * Scalar`: 4-bit fixed-window, table of 16 Jacobian multiples in local memory
static __device__ __noinline__ Jac jac_mul(const Jac& base, const U256& k) {
    Jac tbl[16];
    tbl[0] = jac_identity();
    tbl[1] = base;
    for (int i = 2; i < 16; i++) tbl[i] = (i & 1) ? jac_add(tbl[i - 1], base) : jac_dbl(tbl[i >> 1]);
    Jac r = jac_identity();
    for (int w = 63; w >= 0; w--) {
        if (w != 63) { r = jac_dbl(r); r = jac_dbl(r); r = jac_dbl(r); r = jac_dbl(r); }
        uint32_t d = (k.v[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) r = jac_add(r, tbl[d]);
    }
    return r;
}
__device__ __forceinline__ Affine pt_mul(const Affine& p, const U256& k) { return jac_to_affine(jac_mul(jac_from_affine(p), k)); }

// Jacobian + affine (8M + 3S), all special cases handled
static __device__ __noinline__ Jac jac_madd(const Jac& p, const Affine& a) {
    if (a.inf) return p;
    if (jac_is_inf(p)) return jac_from_affine(a);
    U256 z1z1 = fe_sqr(p.z);
    U256 u2 = fe_mul(a.x, z1z1), s2 = fe_mul(fe_mul(a.y, p.z), z1z1);
    U256 h = fe_sub(u2, p.x), r = fe_sub(s2, p.y);
    if (u256_is_zero(h)) {
        if (u256_is_zero(r)) return jac_dbl(p);
        return jac_identity();
    }
    U256 hh = fe_sqr(h), hhh = fe_mul(hh, h), v = fe_mul(p.x, hh);
    Jac o;
    o.x = fe_sub(fe_sub(fe_sqr(r), hhh), fe_dbl(v));
    o.y = fe_sub(fe_mul(r, fe_sub(v, o.x)), fe_mul(p.y, hhh));
    o.z = fe_mul(p.z, h);
    return o;
}

// Fixed-base tables for the two constant points of the protocol (the generator and curv's
// base_point2): T[b][w][d-1] = d * 16^w * B_b, affine, w < 64, d = 1..15 (123 KB, built once per
// context by fb_points_build).  k * B_b is then at most 64 mixed additions and no doubling.
static constexpr int FBP_WINDOWS = 64, FBP_DIGITS = 15;
static __device__ const uint32_t* g_fb_points = nullptr;      // set per translation unit, see set_fb_points()

__device__ __forceinline__ Jac jac_mul_fixed(int which, const U256& k) {
    const uint32_t* t = g_fb_points + (size_t)which * FBP_WINDOWS * FBP_DIGITS * 16;
    Jac r = jac_identity();
    for (int w = 0; w < FBP_WINDOWS; w++) {
        uint32_t d = (k.v[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) {
            Affine a;
            const uint32_t* e = t + ((size_t)w * FBP_DIGITS + (d - 1)) * 16;
            a.x = u256_load(e); a.y = u256_load(e + 8); a.inf = false;
            r = jac_madd(r, a);
        }
    }
    return r;
}
// one thread per (base, window)
static __global__ void fb_points_build(uint32_t* tables) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * FBP_WINDOWS) return;
    const int which = id / FBP_WINDOWS, w = id % FBP_WINDOWS;
    Jac b = jac_from_affine(which ? affine_H() : affine_G());
    for (int i = 0; i < 4 * w; i++) b = jac_dbl(b);
    Jac acc = b;
    uint32_t* out = tables + ((size_t)which * FBP_WINDOWS + w) * FBP_DIGITS * 16;
    for (int d = 1; d <= FBP_DIGITS; d++) {
        Affine a = jac_to_affine(acc);
        u256_store(out + (size_t)(d - 1) * 16, a.x); u256_store(out + (size_t)(d - 1) * 16 + 8, a.y);
        acc = jac_add(acc, b);
    }
}
__device__ __forceinline__ bool affine_eq(const Affine& a, const Affine& b) {
    if (a.inf || b.inf) return a.inf && b.inf;
    return u256_eq(a.x, b.x) && u256_eq(a.y, b.y);
}
__device__ __forceinline__ bool on_curve(const Affine& a) {
    if (a.inf) return true;
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
