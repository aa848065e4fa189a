// Single-thread multi-precision helpers for the "glue" steps between the job-list kernels:
// the reference's plain-integer arithmetic that is NOT a modular exponentiation —
// `e * a + alpha` (/root/reference/src/utilities/mta/range_proofs.rs:87-88), `(alpha * N + 1)`
// (:53), `s1 > q^3` (:118), Paillier's L-function / CRT recombination (kzen-paillier decrypt,
// called at src/utilities/mta/mod.rs:165), and the one-time per-key constants.  Operands live
// in global or local memory as little-endian uint32 limbs; everything is a plain loop.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tecdsa {
namespace st {

__device__ __forceinline__ void copy(uint32_t* d, const uint32_t* s, int n) { for (int i = 0; i < n; i++) d[i] = s[i]; }
__device__ __forceinline__ void zero(uint32_t* d, int n) { for (int i = 0; i < n; i++) d[i] = 0; }
__device__ __forceinline__ bool is_zero(const uint32_t* a, int n) { uint32_t x = 0; for (int i = 0; i < n; i++) x |= a[i]; return x == 0; }
// -1, 0, 1
__device__ __forceinline__ int cmp(const uint32_t* a, const uint32_t* b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1; }
    return 0;
}
// compare a (na limbs) with b (nb limbs)
__device__ __forceinline__ int cmp2(const uint32_t* a, int na, const uint32_t* b, int nb) {
    int n = na > nb ? na : nb;
    for (int i = n - 1; i >= 0; i--) {
        uint32_t x = i < na ? a[i] : 0, y = i < nb ? b[i] : 0;
        if (x != y) return x > y ? 1 : -1;
    }
    return 0;
}
// d = a + b (n limbs), returns carry
__device__ __forceinline__ uint32_t add(uint32_t* d, const uint32_t* a, const uint32_t* b, int n) {
    uint64_t c = 0;
    for (int i = 0; i < n; i++) { c += (uint64_t)a[i] + b[i]; d[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
// d = a - b (n limbs), returns borrow
__device__ __forceinline__ uint32_t sub(uint32_t* d, const uint32_t* a, const uint32_t* b, int n) {
    int64_t c = 0;
    for (int i = 0; i < n; i++) { c += (int64_t)a[i] - b[i]; d[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
}
// d (na+nb limbs) = a * b
__device__ __forceinline__ void mul(uint32_t* d, const uint32_t* a, int na, const uint32_t* b, int nb) {
    zero(d, na + nb);
    for (int i = 0; i < na; i++) {
        uint64_t c = 0;
        uint32_t ai = a[i];
        for (int j = 0; j < nb; j++) { c += (uint64_t)ai * b[j] + d[i + j]; d[i + j] = (uint32_t)c; c >>= 32; }
        d[i + nb] = (uint32_t)c;
    }
}
// d (nd limbs, nd >= na+nb) = a*b + c  (c has nc <= nd limbs); exact, no modulus.  The running sum lives in a thread-private
// buffer (L1-resident local memory) and is written to `d` once: `d` is usually a global arena row, and a read-modify-write of
// global memory in the inner loop is what made the plain-integer glue kernels slow.
__device__ __forceinline__ void mul_add(uint32_t* d, int nd, const uint32_t* a, int na, const uint32_t* b, int nb,
                                        const uint32_t* c, int nc) {
    constexpr int CAP = 132;
    uint32_t t[CAP];
    uint32_t* w = nd <= CAP ? t : d;
    for (int i = 0; i < nd; i++) w[i] = i < nc ? c[i] : 0;
    for (int i = 0; i < na; i++) {
        uint64_t cy = 0;
        const uint32_t ai = a[i];
        for (int j = 0; j < nb; j++) { cy += (uint64_t)ai * b[j] + w[i + j]; w[i + j] = (uint32_t)cy; cy >>= 32; }
        for (int k = i + nb; cy && k < nd; k++) { cy += w[k]; w[k] = (uint32_t)cy; cy >>= 32; }
    }
    if (w != d) for (int i = 0; i < nd; i++) d[i] = w[i];
}
// low n limbs of a*b (both n limbs)
__device__ __forceinline__ void mul_low(uint32_t* d, const uint32_t* a, const uint32_t* b, int n) {
    zero(d, n);
    for (int i = 0; i < n; i++) {
        uint64_t c = 0;
        uint32_t ai = a[i];
        for (int j = 0; i + j < n; j++) { c += (uint64_t)ai * b[j] + d[i + j]; d[i + j] = (uint32_t)c; c >>= 32; }
    }
}
// -m^-1 mod 2^32
__device__ __forceinline__ uint32_t neg_inv32_st(uint32_t m0) {
    uint32_t x = m0;
    for (int i = 0; i < 5; i++) x *= 2u - m0 * x;
    return 0u - x;
}
// Montgomery product d = a*b*R^-1 mod m, R = 2^(32n), a<m or b<m; t is scratch of 2n+1 limbs.  d may alias a or b.
__device__ __forceinline__ void mont_mul(uint32_t* d, const uint32_t* a, const uint32_t* b, const uint32_t* m, uint32_t m0inv,
                                         int n, uint32_t* t) {
    zero(t, 2 * n + 1);
    for (int i = 0; i < n; i++) {
        uint64_t c = 0;
        uint32_t ai = a[i];
        for (int j = 0; j < n; j++) { c += (uint64_t)ai * b[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        for (int k = i + n; c && k <= 2 * n; k++) { c += t[k]; t[k] = (uint32_t)c; c >>= 32; }
        uint32_t q = t[i] * m0inv;
        c = 0;
        for (int j = 0; j < n; j++) { c += (uint64_t)q * m[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        for (int k = i + n; c && k <= 2 * n; k++) { c += t[k]; t[k] = (uint32_t)c; c >>= 32; }
    }
    if (t[2 * n] || cmp(t + n, m, n) >= 0) sub(d, t + n, m, n);
    else copy(d, t + n, n);
}
// r = 2^(32n) mod m for m with its top bit possibly clear (bit-serial; one-time setup only)
__device__ __forceinline__ void r_mod_m(uint32_t* r, const uint32_t* m, int n) {
    // r = 1; 32n doublings mod m
    zero(r, n); r[0] = 1;
    if (cmp(r, m, n) >= 0) { zero(r, n); return; }          // m == 1
    for (int i = 0; i < 32 * n; i++) {
        uint32_t top = r[n - 1] >> 31;
        for (int j = n - 1; j > 0; j--) r[j] = (r[j] << 1) | (r[j - 1] >> 31);
        r[0] <<= 1;
        if (top || cmp(r, m, n) >= 0) sub(r, r, m, n);
    }
}
// d = a mod m for a of na limbs (any size), bit-serial reduction via doubling (setup / rare paths only)
__device__ __forceinline__ void mod_slow(uint32_t* d, const uint32_t* a, int na, const uint32_t* m, int n) {
    zero(d, n);
    for (int i = 32 * na - 1; i >= 0; i--) {
        uint32_t top = d[n - 1] >> 31;
        for (int j = n - 1; j > 0; j--) d[j] = (d[j] << 1) | (d[j - 1] >> 31);
        d[0] = (d[0] << 1) | ((a[i >> 5] >> (i & 31)) & 1u);
        if (top || cmp(d, m, n) >= 0) sub(d, d, m, n);
    }
}

}  // namespace st
}  // namespace tecdsa
