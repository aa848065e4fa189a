// Montgomery SQUARING for lane-distributed operands: r = a*a*R^-1 mod n with fewer multiply-accumulates than mont_mul(a, a).
//
// mont_mul interleaves K product rows with K reduction rows: 2*K*L wide MACs per lane.  A square has symmetric partial
// products, but under SIMT the usual "skip the lower triangle" buys nothing: in every row the lanes that could skip their
// products wait for the lanes that cannot.  Here the work is re-partitioned in BLOCKS of L x L limbs (block (g, h) = limbs of
// lane g times limbs of lane h, at limb offset (g+h)*L):
//   * lane t computes every unordered block {g <= h} with g + h = t (mod TPI): TPI/2 + 1 blocks at most, all lanes busy in
//     every step (both operands of a block come through shuffles, the 2L-limb block product is computed in-lane);
//   * off-diagonal blocks are added twice (a one-bit shift), diagonal blocks once, into two in-lane accumulators — the blocks
//     that land in the low half of the 2K-limb square (g + h = t) and those in the high half (g + h = t + TPI);
//   * the accumulators of neighbouring lanes overlap by L limbs: three group-wide additions assemble T = a^2 as (Tlo, Thi);
//   * the Montgomery quotient only depends on Tlo: K reduction-only rows (L MACs each) give U = (Tlo + m*n)/R, and
//     r = U + Thi (< 2n), one conditional subtraction.
// MACs per lane: (TPI/2 + 1)*L^2 + K*L instead of 2*K*L — 12.5 % fewer for 4 lanes, 18.75 % fewer for 8 lanes.
// The value is the one mont_mul(a, a) returns (canonical residue): tests/test_modexp_gpu.py compares the two paths and GMP.
#pragma once
#include "bigint.cuh"

namespace tecdsa {

// One row of an in-lane product x * y (y arrives limb by limb): returns the finished limb of the product.
template <int L>
__device__ __forceinline__ uint32_t lane_mul_row(uint32_t (&A)[L + 2], uint32_t (&B)[L + 2], const uint32_t (&x)[L], uint32_t b) {
    B[L + 1] = 0;
    A[0] = add_cc(A[0], B[1]);
    madc_odd_rshift<L>(B, x, b);
    mad_even<L>(A, x, b);
    return A[0];
}
// (lo, hi) = x * (limbs y[0..L) of lane `src`), entirely inside the lane
template <int TPI, int L>
__device__ __forceinline__ void lane_mul(uint32_t (&lo)[L], uint32_t (&hi)[L], const uint32_t (&x)[L], const uint32_t (&y)[L], int src) {
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L + 2; j++) { E[j] = 0; O[j] = 0; }
#pragma unroll
    for (int li = 0; li < L; li += 2) {
        const uint32_t b0 = __shfl_sync(FULL, y[li], src, TPI);
        const uint32_t b1 = __shfl_sync(FULL, y[li + 1], src, TPI);
        lo[li] = lane_mul_row<L>(E, O, x, b0);
        lo[li + 1] = lane_mul_row<L>(O, E, x, b1);
    }
    // O: even set with column 0 consumed, E: odd set (as after the last row of mont_mul)
    hi[0] = add_cc(O[1], E[0]);
#pragma unroll
    for (int j = 1; j < L; j++) hi[j] = addc_cc(O[j + 1], E[j]);
    (void)addc(0, 0);                                   // the product has exactly 2L limbs
}

// Reduction-only Montgomery row: the accumulators take q*n for the quotient digit q of their current low limb.
template <int TPI, int L>
__device__ __forceinline__ uint32_t redc_row(uint32_t (&A)[L + 2], uint32_t (&B)[L + 2], const uint32_t (&n)[L], uint32_t n0inv, uint32_t inc) {
    B[L] = add_cc(B[L], inc);
    B[L + 1] = addc(0, 0);
    const uint32_t q = __shfl_sync(FULL, (A[0] + B[1]) * n0inv, 0, TPI);
    A[0] = add_cc(A[0], B[1]);                          // its carry enters the chain below
    madc_odd_rshift<L>(B, n, q);
    mad_even<L>(A, n, q);
    const uint32_t dn = __shfl_down_sync(FULL, A[0], 1, TPI);
    return (group_lane<TPI>() == TPI - 1) ? 0u : dn;
}

// t += y group-wide (carry lookahead across the lanes); returns the carry out of the group
template <int TPI, int L>
__device__ __forceinline__ uint32_t group_add(uint32_t (&t)[L], const uint32_t (&y)[L]) {
    return group_add_masked<TPI, L>(t, y, 0xffffffffu);
}
// t += w at limb 0 of lane `at` only (w is a small word held by every lane that should add; others pass 0)
template <int TPI, int L>
__device__ __forceinline__ uint32_t group_add_word(uint32_t (&t)[L], uint32_t w) {
    uint32_t y[L];
#pragma unroll
    for (int j = 0; j < L; j++) y[j] = 0;
    y[0] = w;
    return group_add_masked<TPI, L>(t, y, 0xffffffffu);
}

template <int TPI, int L>
__device__ __forceinline__ void mont_sqr(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&n)[L], uint32_t n0inv) {
    static_assert(TPI >= 2 && (TPI & (TPI - 1)) == 0 && L >= 2 && (L % 2) == 0, "shape");
    constexpr int STEPS = TPI / 2 + 1;
    const int gl = group_lane<TPI>();
    // accumulators: blocks with g + h == gl (low half of the square) and with g + h == gl + TPI (high half); 2L limbs + a top word
    uint32_t accl[2 * L + 1], acch[2 * L + 1];
#pragma unroll
    for (int j = 0; j < 2 * L + 1; j++) { accl[j] = 0; acch[j] = 0; }
    int g = -1;
#pragma unroll 1
    for (int s = 0; s < STEPS; s++) {
        // the next g with g <= h = (gl - g) mod TPI; lanes that have run out of blocks go through the motions with weight 0
        int h = 0;
        bool valid = false;
        for (int c = g + 1; c < TPI; c++) {
            const int hh = (gl - c) & (TPI - 1);
            if (c <= hh) { g = c; h = hh; valid = true; break; }
        }
        if (!valid) { g = TPI; h = 0; }
        const int gs = valid ? g : 0;
        uint32_t x[L], lo[L], hi[L];
#pragma unroll
        for (int j = 0; j < L; j++) x[j] = __shfl_sync(FULL, a[j], gs, TPI);
        lane_mul<TPI, L>(lo, hi, x, a, h);
        const bool twice = valid && gs != h;
        const bool high = (gs + h) >= TPI;
        const uint32_t ml = (valid && !high) ? 0xffffffffu : 0u, mh = (valid && high) ? 0xffffffffu : 0u;
        // p = (hi:lo) << (twice ? 1 : 0), 2L + 1 limbs, added to one of the two accumulators
        uint32_t prev = 0;
        uint32_t cl = 0, ch = 0;
#pragma unroll
        for (int j = 0; j < 2 * L + 1; j++) {
            const uint32_t cur = j < L ? lo[j] : (j < 2 * L ? hi[j - L] : 0u);
            const uint32_t p = twice ? ((cur << 1) | (prev >> 31)) : cur;
            prev = cur;
            // two independent carry chains kept in ordinary registers (the PTX carry flag cannot be held across both)
            const uint64_t sl = (uint64_t)accl[j] + (p & ml) + cl;
            accl[j] = (uint32_t)sl; cl = (uint32_t)(sl >> 32);
            const uint64_t sh = (uint64_t)acch[j] + (p & mh) + ch;
            acch[j] = (uint32_t)sh; ch = (uint32_t)(sh >> 32);
        }
    }
    // ---- assemble T = a^2: digit t (L limbs) of the low half = accl.low(t) + accl.high(t-1) + accl.top(t-2); the high half takes
    // acch the same way plus what spills over from the low half's last lanes
    uint32_t Tlo[L], Thi[L], y[L];
#pragma unroll
    for (int j = 0; j < L; j++) { Tlo[j] = accl[j]; Thi[j] = acch[j]; }
    const int below1 = (gl + TPI - 1) & (TPI - 1), below2 = (gl + TPI - 2) & (TPI - 1);
    // second halves, rotated one lane up
#pragma unroll
    for (int j = 0; j < L; j++) y[j] = __shfl_sync(FULL, accl[L + j], below1, TPI);
    uint32_t yl[L], yh[L];
#pragma unroll
    for (int j = 0; j < L; j++) { yl[j] = gl >= 1 ? y[j] : 0u; yh[j] = gl == 0 ? y[j] : 0u; }
#pragma unroll
    for (int j = 0; j < L; j++) { const uint32_t v = __shfl_sync(FULL, acch[L + j], below1, TPI); if (gl >= 1) yh[j] = v; }
    uint32_t c = group_add<TPI, L>(Tlo, yl);
    (void)group_add<TPI, L>(Thi, yh);
    (void)group_add_word<TPI, L>(Thi, gl == 0 ? c : 0u);
    // top words, rotated two lanes up
    {
        const uint32_t tl = __shfl_sync(FULL, accl[2 * L], below2, TPI), th = __shfl_sync(FULL, acch[2 * L], below2, TPI);
        c = group_add_word<TPI, L>(Tlo, gl >= 2 ? tl : 0u);
        (void)group_add_word<TPI, L>(Thi, gl >= 2 ? th : tl);      // lanes 0 and 1 take the low half's spill-over
        (void)group_add_word<TPI, L>(Thi, gl == 0 ? c : 0u);
    }
    // ---- U = (Tlo + m*n) / R through K reduction-only rows, then r = U + Thi
    uint32_t E[L + 2], O[L + 2];
#pragma unroll
    for (int j = 0; j < L; j++) { E[j] = Tlo[j]; O[j] = 0; }
    E[L] = 0; E[L + 1] = 0; O[L] = 0; O[L + 1] = 0;
    uint32_t inc = 0;
#pragma unroll 1
    for (int gi = 0; gi < TPI; gi++) {
#pragma unroll
        for (int li = 0; li < L; li += 2) {
            inc = redc_row<TPI, L>(E, O, n, n0inv, inc);
            inc = redc_row<TPI, L>(O, E, n, n0inv, inc);
        }
    }
    uint32_t U[L];
    uint32_t ov = rows_finish<TPI, L>(U, E, O, inc);
    ov += group_add<TPI, L>(U, Thi);
    reduce_once<TPI, L>(r, U, ov, n);
}

}  // namespace tecdsa
