// Group big-integer arithmetic for sm_100a.
//
// A K-limb (32-bit limbs, little-endian) integer is spread over a group of TPI
// consecutive lanes of one warp; every lane keeps L = K/TPI contiguous limbs in
// registers (lane g of the group owns limbs [g*L, (g+1)*L)).  TPI = 32 is the
// "one warp per operand" shape; smaller TPI packs several operands per warp.
// All multiply-accumulate work is written as mad.lo.cc / madc.hi.cc pairs that
// ptxas fuses into one IMAD.WIDE.U32(.X) each (32x32+64 with carry in/out), on
// two interleaved accumulator sets (even / odd columns) so no carry ever has to
// be re-aligned.  Cross-lane traffic is three warp shuffles per Montgomery row
// (multiplier limb, quotient limb, limb shifted down to the neighbour) plus a
// ballot-based carry resolve once per product.
//
// Reference arithmetic this replaces: curv-kzen `BigInt::mod_pow / mod_mul`
// (GMP mpz_powm) as called from /root/reference/src/utilities/mta/range_proofs.rs:52-57,
// 86, 122-141 and src/utilities/zk_pdl_with_slack/mod.rs:182-199.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tecdsa {

static constexpr uint32_t FULL = 0xffffffffu;

// ---------------------------------------------------------------- PTX carry helpers
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}

// Two accumulator sets per lane, L+2 registers each.  As the EVEN set X[j] is column j;
// as the ODD set X[j] is column j+1.  A 32x32 product at column c always lands on an
// aligned register pair of one of the two sets, so every MAC is a single IMAD.WIDE.U32.X.

// even set: pairs (X[j],X[j+1]) += a[j]*b for even j; the carry ripples pair to pair into X[L].
template <int L>
__device__ __forceinline__ void mad_even(uint32_t (&X)[L + 2], const uint32_t (&a)[L], uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(X[0]), "+r"(X[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int j = 2; j < L; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(X[j]), "+r"(X[j + 1]) : "r"(a[j]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(X[L]));
}
// odd set: pairs (X[j-1],X[j]) += a[j]*b for odd j; carry into X[L].
template <int L>
__device__ __forceinline__ void mad_odd(uint32_t (&X)[L + 2], const uint32_t (&a)[L], uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(X[0]), "+r"(X[1]) : "r"(a[1]), "r"(b));
#pragma unroll
    for (int j = 3; j < L; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(X[j - 1]), "+r"(X[j]) : "r"(a[j]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(X[L]));
}
// The previous row's even set (its column 0 already consumed) turns into this row's odd
// set: X[j] <- X[j+2] fused into the MAC, with the carry flag on entry taken as carry-in
// (it comes from folding the stray column-1 limb into the other set).
template <int L>
__device__ __forceinline__ void madc_odd_rshift(uint32_t (&X)[L + 2], const uint32_t (&a)[L], uint32_t b) {
#pragma unroll
    for (int j = 1; j < L; j += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                     : "=&r"(X[j - 1]), "=r"(X[j]) : "r"(a[j]), "r"(b), "r"(X[j + 1]), "r"(X[j + 2]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(X[L]));
    X[L + 1] = 0;
}

// ---------------------------------------------------------------- group helpers
template <int TPI> __device__ __forceinline__ int group_lane() { return threadIdx.x & (TPI - 1); }

// Carry-lookahead across the lanes of a group.  `cout` is this lane's carry out
// (0/1), `ones` says all its limbs are 0xffffffff (it would propagate).  Returns the
// carry INTO this lane; *top gets the carry out of the group's last lane.
template <int TPI>
__device__ __forceinline__ uint32_t group_carry(uint32_t cout, bool ones, uint32_t* top) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = lane & ~(uint32_t)(TPI - 1);
    uint64_t gmask = (TPI == 32) ? 0xffffffffull : ((1ull << TPI) - 1);
    uint64_t g = (uint64_t)(__ballot_sync(FULL, cout != 0) >> base) & gmask;
    uint64_t p = (uint64_t)(__ballot_sync(FULL, ones) >> base) & gmask;
    uint64_t gs = g << 1;
    uint64_t cin = ((p + gs) ^ p ^ gs) | gs;
    *top = (uint32_t)(cin >> TPI) & 1u;
    return (uint32_t)(cin >> (lane - base)) & 1u;
}

template <int L> __device__ __forceinline__ bool all_ones(const uint32_t (&t)[L]) {
    uint32_t x = t[0];
#pragma unroll
    for (int j = 1; j < L; j++) x &= t[j];
    return x == 0xffffffffu;
}

// t += c (c is 0/1) rippling through this lane's limbs only (the lookahead above has
// already accounted for the carry this may push out).
template <int L> __device__ __forceinline__ void add_small(uint32_t (&t)[L], uint32_t c) {
    t[0] = add_cc(t[0], c);
#pragma unroll
    for (int j = 1; j < L; j++) t[j] = addc_cc(t[j], 0);
    (void)addc(0, 0);
}

// Group-wide t = t + (m & ~n) + (ov at limb 0)  == t - n (mod 2^(32K)) when m = ~0, ov = 1.
// Returns the carry out of the whole K-limb addition (1 iff t >= n for the subtract case).
template <int TPI, int L>
__device__ __forceinline__ uint32_t group_sub_masked(uint32_t (&t)[L], const uint32_t (&n)[L], uint32_t m, uint32_t ov) {
    const int gl = group_lane<TPI>();
    uint32_t cin0 = (gl == 0) ? ov : 0u;
    (void)add_cc(cin0, 0xffffffffu);                 // CC.CF = cin0
    t[0] = addc_cc(t[0], (~n[0]) & m);
#pragma unroll
    for (int j = 1; j < L; j++) t[j] = addc_cc(t[j], (~n[j]) & m);
    uint32_t c = addc(0, 0);
    uint32_t top;
    uint32_t ci = group_carry<TPI>(c, all_ones<L>(t), &top);
    add_small<L>(t, ci);
    return top;
}

// t += (m & y), group-wide; returns carry out
template <int TPI, int L>
__device__ __forceinline__ uint32_t group_add_masked(uint32_t (&t)[L], const uint32_t (&y)[L], uint32_t m) {
    t[0] = add_cc(t[0], y[0] & m);
#pragma unroll
    for (int j = 1; j < L; j++) t[j] = addc_cc(t[j], y[j] & m);
    uint32_t c = addc(0, 0);
    uint32_t top;
    uint32_t ci = group_carry<TPI>(c, all_ones<L>(t), &top);
    add_small<L>(t, ci);
    return top;
}

// One Montgomery row.  A is this row's even set; B was the previous row's even set (still
// unshifted, column 0 gone) and becomes this row's odd set; `inc` is the limb the right
// neighbour handed down after the previous row.  Returns the limb to hand to the left.
template <int TPI, int L>
__device__ __forceinline__ uint32_t mont_row(uint32_t (&A)[L + 2], uint32_t (&B)[L + 2], const uint32_t (&a)[L],
                                             const uint32_t (&n)[L], uint32_t bi, uint32_t n0inv, uint32_t inc) {
    B[L] = add_cc(B[L], inc);
    B[L + 1] = addc(0, 0);
    A[0] = add_cc(A[0], B[1]);          // stray column; its carry enters the chain below
    madc_odd_rshift<L>(B, a, bi);
    mad_even<L>(A, a, bi);
    uint32_t q = __shfl_sync(FULL, A[0] * n0inv, 0, TPI);
    mad_odd<L>(B, n, q);
    mad_even<L>(A, n, q);
    uint32_t dn = __shfl_down_sync(FULL, A[0], 1, TPI);
    return (group_lane<TPI>() == TPI - 1) ? 0u : dn;
}

// Montgomery product r = a*b*R^-1 mod n, R = 2^(32K), n odd.  With one operand < n and the
// other < R the value before the final step is < 2n, so the single conditional subtract
// returns the canonical residue in [0, n): every Montgomery-domain value in the engine is
// canonical.
template <int TPI, int L>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L],
                                         const uint32_t (&n)[L], uint32_t n0inv) {
    static_assert(L >= 2 && (L % 2) == 0, "L must be even");
    const int gl = group_lane<TPI>();
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; j++) { E[j] = 0; O[j] = 0; }
    uint32_t inc = 0;
#pragma unroll 1
    for (int gi = 0; gi < TPI; gi++) {
#pragma unroll
        for (int li = 0; li < L; li += 2) {
            uint32_t b0 = __shfl_sync(FULL, b[li], gi, TPI);
            uint32_t b1 = __shfl_sync(FULL, b[li + 1], gi, TPI);
            inc = mont_row<TPI, L>(E, O, a, n, b0, n0inv, inc);
            inc = mont_row<TPI, L>(O, E, a, n, b1, n0inv, inc);
        }
    }
    // After the last row: O is the even set with column 0 consumed (value O >> 32, plus
    // `inc` at column L-1), E is the odd set (E[j] at column j after the shift).
    uint32_t T[L], h0, h1;
    O[L] = add_cc(O[L], inc);
    uint32_t x = addc(0, 0);
    T[0] = add_cc(O[1], E[0]);
#pragma unroll
    for (int j = 1; j < L; j++) T[j] = addc_cc(O[j + 1], E[j]);
    h0 = addc_cc(E[L], x);
    h1 = addc(0, 0);
    uint32_t i0 = __shfl_up_sync(FULL, h0, 1, TPI);
    uint32_t i1 = __shfl_up_sync(FULL, h1, 1, TPI);
    if (gl == 0) { i0 = 0; i1 = 0; }
    T[0] = add_cc(T[0], i0);
    T[1] = addc_cc(T[1], i1);
#pragma unroll
    for (int j = 2; j < L; j++) T[j] = addc_cc(T[j], 0);
    uint32_t c = addc(0, 0);
    uint32_t top;
    uint32_t ci = group_carry<TPI>(c, all_ones<L>(T), &top);
    add_small<L>(T, ci);
    uint32_t ov = __shfl_sync(FULL, h0, TPI - 1, TPI) + top;   // 0 or 1: value >= R
    // a, b < n  =>  value < 2n: one conditional subtract gives the canonical residue
    uint32_t D[L];
#pragma unroll
    for (int j = 0; j < L; j++) D[j] = T[j];
    uint32_t ge = group_sub_masked<TPI, L>(D, n, 0xffffffffu, 1u);
    const bool take = (ov | ge) != 0;
#pragma unroll
    for (int j = 0; j < L; j++) r[j] = take ? D[j] : T[j];
}

// ---------------------------------------------------------------------------------------------
// Pieces of the row machinery shared with the N-adic arithmetic modulo a perfect square (nadic.cuh):
//   rows_finish / reduce_once : the tail of a product (set merge, carry resolve, conditional subtract)
//   group_mul_wide : the plain 2K-limb product a*b (no reduction), low half / high half per lane

// merge the two accumulator sets after the last row into K limbs + the overflow word
template <int TPI, int L>
__device__ __forceinline__ uint32_t rows_finish(uint32_t (&T)[L], uint32_t (&E)[L + 2], uint32_t (&O)[L + 2], uint32_t inc) {
    const int gl = group_lane<TPI>();
    uint32_t h0, h1;
    O[L] = add_cc(O[L], inc);
    uint32_t x = addc(0, 0);
    T[0] = add_cc(O[1], E[0]);
#pragma unroll
    for (int j = 1; j < L; j++) T[j] = addc_cc(O[j + 1], E[j]);
    h0 = addc_cc(E[L], x);
    h1 = addc(0, 0);
    uint32_t i0 = __shfl_up_sync(FULL, h0, 1, TPI);
    uint32_t i1 = __shfl_up_sync(FULL, h1, 1, TPI);
    if (gl == 0) { i0 = 0; i1 = 0; }
    T[0] = add_cc(T[0], i0);
    T[1] = addc_cc(T[1], i1);
#pragma unroll
    for (int j = 2; j < L; j++) T[j] = addc_cc(T[j], 0);
    uint32_t c = addc(0, 0);
    uint32_t top;
    uint32_t ci = group_carry<TPI>(c, all_ones<L>(T), &top);
    add_small<L>(T, ci);
    return __shfl_sync(FULL, h0, TPI - 1, TPI) + top;           // words above the K limbs (0 or 1)
}
// value < 2n (ov = its bit above the K limbs) -> canonical residue
template <int TPI, int L>
__device__ __forceinline__ void reduce_once(uint32_t (&r)[L], const uint32_t (&T)[L], uint32_t ov, const uint32_t (&n)[L]) {
    uint32_t D[L];
#pragma unroll
    for (int j = 0; j < L; j++) D[j] = T[j];
    uint32_t ge = group_sub_masked<TPI, L>(D, n, 0xffffffffu, 1u);
    const bool take = (ov | ge) != 0;
#pragma unroll
    for (int j = 0; j < L; j++) r[j] = take ? D[j] : T[j];
}

// product-only row; lane 0's column 0 is a finished limb of the product
template <int TPI, int L>
__device__ __forceinline__ uint32_t mul_row(uint32_t (&A)[L + 2], uint32_t (&B)[L + 2], const uint32_t (&a)[L], uint32_t bi, uint32_t inc, uint32_t& low) {
    B[L] = add_cc(B[L], inc);
    B[L + 1] = addc(0, 0);
    A[0] = add_cc(A[0], B[1]);
    madc_odd_rshift<L>(B, a, bi);
    mad_even<L>(A, a, bi);
    low = __shfl_sync(FULL, A[0], 0, TPI);
    uint32_t dn = __shfl_down_sync(FULL, A[0], 1, TPI);
    return (group_lane<TPI>() == TPI - 1) ? 0u : dn;
}
template <int TPI, int L>
__device__ __forceinline__ void group_mul_wide(uint32_t (&lo)[L], uint32_t (&hi)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    const int gl = group_lane<TPI>();
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; j++) { E[j] = 0; O[j] = 0; }
    uint32_t inc = 0;
#pragma unroll 1
    for (int gi = 0; gi < TPI; gi++) {
        const bool mine = gi == gl;
#pragma unroll
        for (int li = 0; li < L; li += 2) {
            uint32_t b0 = __shfl_sync(FULL, b[li], gi, TPI);
            uint32_t b1 = __shfl_sync(FULL, b[li + 1], gi, TPI);
            uint32_t l0, l1;
            inc = mul_row<TPI, L>(E, O, a, b0, inc, l0);
            inc = mul_row<TPI, L>(O, E, a, b1, inc, l1);
            if (mine) { lo[li] = l0; lo[li + 1] = l1; }
        }
    }
    (void)rows_finish<TPI, L>(hi, E, O, inc);
}

// t >= n ? t - n : t   (one conditional subtract; canonical output for t < 2n)
template <int TPI, int L>
__device__ __forceinline__ void cond_sub(uint32_t (&t)[L], const uint32_t (&n)[L]) {
    uint32_t d[L];
#pragma unroll
    for (int j = 0; j < L; j++) d[j] = t[j];
    uint32_t ge = group_sub_masked<TPI, L>(d, n, 0xffffffffu, 1u);
    if (ge) {
#pragma unroll
        for (int j = 0; j < L; j++) t[j] = d[j];
    }
}

// t = 2t mod n for canonical t < n (any n < R).
template <int TPI, int L>
__device__ __forceinline__ void mod_double(uint32_t (&t)[L], const uint32_t (&n)[L]) {
    const int gl = group_lane<TPI>();
    uint32_t msb = t[L - 1] >> 31;
    uint32_t inb = __shfl_up_sync(FULL, msb, 1, TPI);
    if (gl == 0) inb = 0;
    uint32_t ov = __shfl_sync(FULL, msb, TPI - 1, TPI);
#pragma unroll
    for (int j = L - 1; j > 0; j--) t[j] = (t[j] << 1) | (t[j - 1] >> 31);
    t[0] = (t[0] << 1) | inb;
    uint32_t D[L];
#pragma unroll
    for (int j = 0; j < L; j++) D[j] = t[j];
    uint32_t ge = group_sub_masked<TPI, L>(D, n, 0xffffffffu, 1u);
    if (ov | ge) {
#pragma unroll
        for (int j = 0; j < L; j++) t[j] = D[j];
    }
}

// group-wide logical shifts by one bit
template <int TPI, int L>
__device__ __forceinline__ void group_shl1(uint32_t (&t)[L]) {
    uint32_t inb = __shfl_up_sync(FULL, t[L - 1] >> 31, 1, TPI);
    if (group_lane<TPI>() == 0) inb = 0;
#pragma unroll
    for (int j = L - 1; j > 0; j--) t[j] = (t[j] << 1) | (t[j - 1] >> 31);
    t[0] = (t[0] << 1) | inb;
}
template <int TPI, int L>
__device__ __forceinline__ void group_shr1(uint32_t (&t)[L]) {
    uint32_t inb = __shfl_down_sync(FULL, t[0] & 1u, 1, TPI);
    if (group_lane<TPI>() == TPI - 1) inb = 0;
#pragma unroll
    for (int j = 0; j < L - 1; j++) t[j] = (t[j] >> 1) | (t[j + 1] << 31);
    t[L - 1] = (t[L - 1] >> 1) | (inb << 31);
}

// x = 2^(32K) mod n, canonical, for ANY n >= 1 (n need not have its top bit set): shift n up
// until its top bit is set, take R - n', then walk n' back down with one compare-subtract per
// bit.  For the 2047..2048-bit moduli of the protocol this is 0..2 iterations.
template <int TPI, int L>
__device__ __forceinline__ void r_mod_n(uint32_t (&x)[L], const uint32_t (&n)[L]) {
    uint32_t m[L];
#pragma unroll
    for (int j = 0; j < L; j++) m[j] = n[j];
    int s = 0;
    while (true) {                                       // warp-uniform trip count
        uint32_t top = __shfl_sync(FULL, m[L - 1] >> 31, TPI - 1, TPI);
        bool need = (top == 0) && (s < 32 * TPI * L);
        if (!__any_sync(FULL, need)) break;
        if (need) { s++; }
        uint32_t sh[L];
#pragma unroll
        for (int j = 0; j < L; j++) sh[j] = m[j];
        group_shl1<TPI, L>(sh);
        if (need) {
#pragma unroll
            for (int j = 0; j < L; j++) m[j] = sh[j];
        }
    }
#pragma unroll
    for (int j = 0; j < L; j++) x[j] = 0;
    (void)group_sub_masked<TPI, L>(x, m, 0xffffffffu, 1u);    // R - n'
    cond_sub<TPI, L>(x, m);
    int smax = s;
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) smax = max(smax, __shfl_xor_sync(FULL, smax, off));
#pragma unroll 1
    for (int i = 0; i < smax; i++) {
        const bool act = i < s;
        uint32_t sh[L], y[L];
#pragma unroll
        for (int j = 0; j < L; j++) { sh[j] = m[j]; y[j] = x[j]; }
        group_shr1<TPI, L>(sh);
        cond_sub<TPI, L>(y, sh);
        if (act) {
#pragma unroll
            for (int j = 0; j < L; j++) { m[j] = sh[j]; x[j] = y[j]; }
        }
    }
}

// -n^-1 mod 2^32 (n odd)
__device__ __forceinline__ uint32_t neg_inv32(uint32_t n0) {
    uint32_t x = n0;                 // 3 correct bits
#pragma unroll
    for (int i = 0; i < 5; i++) x *= 2u - n0 * x;
    return 0u - x;
}

// Vectorised load/store of this lane's L limbs of operand `idx` (operand-major layout,
// K = TPI*L limbs each, 16-byte aligned rows).
template <int TPI, int L>
__device__ __forceinline__ void load_limbs(uint32_t (&x)[L], const uint32_t* __restrict__ p) {
    const int gl = group_lane<TPI>();
    if constexpr (L % 4 == 0) {
        const uint4* q = reinterpret_cast<const uint4*>(p) + gl * (L / 4);
#pragma unroll
        for (int j = 0; j < L / 4; j++) { uint4 v = q[j]; x[4*j] = v.x; x[4*j+1] = v.y; x[4*j+2] = v.z; x[4*j+3] = v.w; }
    } else {
        const uint2* q = reinterpret_cast<const uint2*>(p) + gl * (L / 2);
#pragma unroll
        for (int j = 0; j < L / 2; j++) { uint2 v = q[j]; x[2*j] = v.x; x[2*j+1] = v.y; }
    }
}
template <int TPI, int L>
__device__ __forceinline__ void store_limbs(uint32_t* __restrict__ p, const uint32_t (&x)[L]) {
    const int gl = group_lane<TPI>();
    if constexpr (L % 4 == 0) {
        uint4* q = reinterpret_cast<uint4*>(p) + gl * (L / 4);
#pragma unroll
        for (int j = 0; j < L / 4; j++) q[j] = make_uint4(x[4*j], x[4*j+1], x[4*j+2], x[4*j+3]);
    } else {
        uint2* q = reinterpret_cast<uint2*>(p) + gl * (L / 2);
#pragma unroll
        for (int j = 0; j < L / 2; j++) q[j] = make_uint2(x[2*j], x[2*j+1]);
    }
}

}  // namespace tecdsa
