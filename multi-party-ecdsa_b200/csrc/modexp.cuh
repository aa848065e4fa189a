// Batched fixed-window Montgomery exponentiation, one lane-group per operand.
//
// Replaces every `BigInt::mod_pow(base, exp, modulus)` on the GG20 offline-signing path
// (/root/reference/src/utilities/mta/range_proofs.rs:52-57,86,122,129-141;
//  src/utilities/zk_pdl_with_slack/mod.rs:189-196; kzen-paillier encrypt/mul/decrypt called
//  from src/utilities/mta/mod.rs:68,133,140,165).  Results are the canonical residue in
// [0, modulus), identical to GMP mpz_powm for odd moduli.
#pragma once
#include "bigint.cuh"
#include "sqr.cuh"

namespace tecdsa {

static constexpr int WINDOW_BITS = 5;

// Executed-work accounting (tecdsa_ctx_work): every job adds the 32x32+64 multiply-accumulates of the products it ran to a
// device counter — one atomicAdd per job, counted from the loop trip counts of the kernel itself.
__host__ __device__ constexpr unsigned long long mac_mont(int K) { return 2ull * K * K + K; }            // one Montgomery product (rows + quotient digits)
__host__ __device__ constexpr unsigned long long mac_sqr(int K, int TPI) { return (unsigned long long)(TPI / 2 + 1) * (K / TPI) * K + (unsigned long long)K * K + K; }   // mont_sqr (sqr.cuh): blocks + reduction rows
__host__ __device__ constexpr int setup_products(int K) { int t = 0; for (unsigned v = 32u * K; v > 1; v >>= 1) t++; return t; }   // squarings of mont_setup

// Everything a group needs to exponentiate modulo one modulus.
template <int L> struct MontCtx {
    uint32_t n[L];
    uint32_t rr[L];      // R^2 mod n
    uint32_t one[L];     // R mod n
    uint32_t n0inv;
};

// R mod n and R^2 mod n for this group's modulus (already in ctx.n).
template <int TPI, int L>
__device__ __forceinline__ void mont_setup(MontCtx<L>& c) {
    const int gl = group_lane<TPI>();
    c.n0inv = neg_inv32(__shfl_sync(FULL, c.n[0], 0, TPI));
    uint32_t x[L];
    r_mod_n<TPI, L>(x, c.n);
#pragma unroll
    for (int j = 0; j < L; j++) c.one[j] = x[j];
    // x = 2^e * R with e = 1, then square (e -> 2e) / double (e -> e+1) up to e = 32*K.
    mod_double<TPI, L>(x, c.n);
    const uint32_t target = 32u * TPI * L;
    int top = 31 - __clz(target);
#pragma unroll 1
    for (int bit = top - 1; bit >= 0; bit--) {
        uint32_t y[L];
#pragma unroll
        for (int j = 0; j < L; j++) y[j] = x[j];
        mont_mul<TPI, L>(x, x, y, c.n, c.n0inv);
        if ((target >> bit) & 1u) mod_double<TPI, L>(x, c.n);
    }
#pragma unroll
    for (int j = 0; j < L; j++) c.rr[j] = x[j];
    (void)gl;
}

// Montgomery form -> canonical residue in [0, n).
template <int TPI, int L>
__device__ __forceinline__ void mont_to_plain(uint32_t (&x)[L], const MontCtx<L>& c) {
    const int gl = group_lane<TPI>();
    uint32_t u[L];
#pragma unroll
    for (int j = 0; j < L; j++) u[j] = 0;
    if (gl == 0) u[0] = 1;
    mont_mul<TPI, L>(x, x, u, c.n, c.n0inv);
}

// exponent window `w` (WINDOW_BITS wide) of a little-endian limb array
__device__ __forceinline__ uint32_t exp_window(const uint32_t* __restrict__ e, int exp_limbs, int w) {
    int bit = w * WINDOW_BITS;
    int limb = bit >> 5, off = bit & 31;
    uint32_t lo = __ldg(e + limb);
    uint32_t hi = (limb + 1 < exp_limbs) ? __ldg(e + limb + 1) : 0u;
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> off) & ((1u << WINDOW_BITS) - 1);
}

// acc = base^exp in Montgomery form.  `tbl` is this group's private window table
// (2^WINDOW_BITS entries of K limbs, operand-major) in global memory.
template <int TPI, int L, bool SQR>
__device__ __forceinline__ void mont_pow(uint32_t (&acc)[L], const uint32_t (&base)[L], const uint32_t* __restrict__ e,
                                         int exp_limbs, const MontCtx<L>& c, uint32_t* __restrict__ tbl) {
    constexpr int K = TPI * L;
    constexpr int TBL = 1 << WINDOW_BITS;
    uint32_t xr[L], t[L];
    mont_mul<TPI, L>(xr, base, c.rr, c.n, c.n0inv);          // base * R
    store_limbs<TPI, L>(tbl, c.one);
    store_limbs<TPI, L>(tbl + K, xr);
#pragma unroll
    for (int j = 0; j < L; j++) t[j] = xr[j];
#pragma unroll 1
    for (int i = 2; i < TBL; i++) {
        mont_mul<TPI, L>(t, t, xr, c.n, c.n0inv);
        store_limbs<TPI, L>(tbl + i * K, t);
    }
    __syncwarp();
    const int nw = (exp_limbs * 32 + WINDOW_BITS - 1) / WINDOW_BITS;
    load_limbs<TPI, L>(acc, tbl + exp_window(e, exp_limbs, nw - 1) * K);
    const int steps = (nw - 1) * (WINDOW_BITS + 1);
    int w = nw - 2, s = 0;
    uint32_t bb[L];
#pragma unroll 1
    for (int it = 0; it < steps; it++) {
        if (s == WINDOW_BITS) {
            load_limbs<TPI, L>(bb, tbl + exp_window(e, exp_limbs, w) * K);
            w--; s = 0;
            mont_mul<TPI, L>(acc, acc, bb, c.n, c.n0inv);
        } else {
            s++;
            if (SQR) mont_sqr<TPI, L>(acc, acc, c.n, c.n0inv);       // dedicated squaring (sqr.cuh): fewer multiply-accumulates
            else {
#pragma unroll
                for (int j = 0; j < L; j++) bb[j] = acc[j];
                mont_mul<TPI, L>(acc, acc, bb, c.n, c.n0inv);
            }
        }
    }
}

// 4 blocks of 128 threads per SM: ptxas fits the 4-lane kernel into 128 registers with 8 bytes of spills (136 without the bound);
// measured 0.3 % faster at 1024 / 2048 / 4096 bits (profiles/r02_modexp_minb.md)
#ifndef MODEXP_MINB
#define MODEXP_MINB 4
#endif
template <int K, int TPI, bool SQR>
__global__ void __launch_bounds__(128, MODEXP_MINB)
modexp_kernel(const uint32_t* __restrict__ base, const uint32_t* __restrict__ exp, const uint32_t* __restrict__ mod,
              const uint32_t* __restrict__ mod_idx, uint32_t* __restrict__ out, uint8_t* __restrict__ status,
              uint32_t* __restrict__ table, int count, int exp_limbs, unsigned long long* __restrict__ work) {
    constexpr int L = K / TPI;
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) / TPI;
    const bool live = slot < count;
    const int idx = live ? slot : count - 1;
    MontCtx<L> c;
    const uint32_t mi = mod_idx ? __ldg(mod_idx + idx) : (uint32_t)idx;
    load_limbs<TPI, L>(c.n, mod + (size_t)mi * K);
    uint32_t b[L], acc[L];
    load_limbs<TPI, L>(b, base + (size_t)idx * K);
    const uint32_t n_low = __shfl_sync(FULL, c.n[0], 0, TPI);
    if ((n_low & 1u) == 0) {                 // Montgomery needs an odd modulus: flag, emit zero
        if (live) {
#pragma unroll
            for (int j = 0; j < L; j++) acc[j] = 0;
            store_limbs<TPI, L>(out + (size_t)idx * K, acc);
            if (group_lane<TPI>() == 0 && status) status[idx] = 1;
        }
        // keep the warp converged for the shuffles of the other groups
        c.n[0] |= (group_lane<TPI>() == 0) ? 1u : 0u;
    }
    mont_setup<TPI, L>(c);
    mont_pow<TPI, L, SQR>(acc, b, exp + (size_t)idx * exp_limbs, exp_limbs, c,
                     table + (size_t)slot * (K << WINDOW_BITS));
    mont_to_plain<TPI, L>(acc, c);
    if (live && (n_low & 1u)) {
        store_limbs<TPI, L>(out + (size_t)idx * K, acc);
        if (group_lane<TPI>() == 0 && status) status[idx] = 0;
        if (group_lane<TPI>() == 0 && work) {
            const int nw = (exp_limbs * 32 + WINDOW_BITS - 1) / WINDOW_BITS;
            const unsigned long long mults = setup_products(K) + 1 + ((1 << WINDOW_BITS) - 2) + (unsigned long long)(nw - 1) + 1;
            const unsigned long long sqrs = (unsigned long long)(nw - 1) * WINDOW_BITS;
            atomicAdd(work, mults * mac_mont(K) + sqrs * (SQR ? mac_sqr(K, TPI) : mac_mont(K)));
        }
    }
}

}  // namespace tecdsa
