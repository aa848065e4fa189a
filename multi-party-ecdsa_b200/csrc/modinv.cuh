// Batched modular inverse for odd moduli, one lane-group per operand.
//
// Replaces `BigInt::mod_inv(a, n) -> Option<BigInt>` (GMP mpz_invert) at
// /root/reference/src/utilities/mta/range_proofs.rs:122,135 and
// src/utilities/zk_pdl_with_slack/mod.rs:192; `ok = 0` is the reference's `None`
// (gcd(a, n) != 1), which the callers turn into "reject" (range_proofs.rs:123-127,136-139).
//
// Algorithm: Kaliski's almost-inverse (shift/subtract only, no per-step modular correction),
// then the 2^-k factor is removed with one or two Montgomery products.  Every step is
// executed branch-free by all groups of the warp; a group whose v reached 0 stops committing.
#pragma once
#include "jobs.cuh"

namespace tecdsa {

// a > b over the whole group (lexicographic from the top limb / top lane)
template <int TPI, int L>
__device__ __forceinline__ bool group_gt(const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    bool gt = false, eq = true;
#pragma unroll
    for (int j = L - 1; j >= 0; j--) {
        if (eq && a[j] != b[j]) { gt = a[j] > b[j]; eq = false; }
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = lane & ~(uint32_t)(TPI - 1);
    const uint32_t gmask = (TPI == 32) ? 0xffffffffu : (((1u << TPI) - 1u) << base);
    uint32_t ne = __ballot_sync(FULL, !eq) & gmask;
    uint32_t g = __ballot_sync(FULL, gt) & gmask;
    if (ne == 0) return false;
    uint32_t top = 31 - __clz(ne);
    return (g >> top) & 1u;
}
template <int TPI, int L>
__device__ __forceinline__ bool group_is_zero(const uint32_t (&a)[L]) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < L; j++) x |= a[j];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = lane & ~(uint32_t)(TPI - 1);
    const uint32_t gmask = (TPI == 32) ? 0xffffffffu : (((1u << TPI) - 1u) << base);
    return (__ballot_sync(FULL, x != 0) & gmask) == 0;
}
template <int TPI, int L>
__device__ __forceinline__ bool group_is_one(const uint32_t (&a)[L]) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 1; j < L; j++) x |= a[j];
    x |= (group_lane<TPI>() == 0) ? (a[0] ^ 1u) : a[0];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = lane & ~(uint32_t)(TPI - 1);
    const uint32_t gmask = (TPI == 32) ? 0xffffffffu : (((1u << TPI) - 1u) << base);
    return (__ballot_sync(FULL, x != 0) & gmask) == 0;
}
// out = a^-1 mod n (canonical); returns false when gcd(a, n) != 1.  `m` holds n and its
// Montgomery constants; a may be any value below 2^(32K).
template <int TPI, int L>
__device__ __forceinline__ bool group_modinv(uint32_t (&out)[L], const uint32_t (&a_in)[L], const MontCtx<L>& m) {
    const int gl = group_lane<TPI>();
    constexpr int MBITS = 32 * TPI * L;
    uint32_t u[L], v[L], r[L], s[L];
    // v = a mod n (canonical): a*R^2/R = aR, then /R
    {
        uint32_t one_plain[L];
#pragma unroll
        for (int j = 0; j < L; j++) one_plain[j] = 0;
        if (gl == 0) one_plain[0] = 1;
        mont_mul<TPI, L>(v, a_in, m.rr, m.n, m.n0inv);
        mont_mul<TPI, L>(v, v, one_plain, m.n, m.n0inv);
    }
#pragma unroll
    for (int j = 0; j < L; j++) { u[j] = m.n[j]; r[j] = 0; s[j] = 0; }
    if (gl == 0) s[0] = 1;
    int k = 0;
    uint32_t rh = 0, sh = 0;            // r, s < 2n may need one bit above the K limbs
#pragma unroll 1
    for (int it = 0; it < 2 * MBITS + 2; it++) {
        const bool active = !group_is_zero<TPI, L>(v);
        if (!__any_sync(FULL, active)) break;
        const uint32_t u0 = __shfl_sync(FULL, u[0], 0, TPI), v0 = __shfl_sync(FULL, v[0], 0, TPI);
        const bool cu = (u0 & 1u) == 0;
        const bool cv = !cu && (v0 & 1u) == 0;
        const bool cs = !cu && !cv;
        const bool gt = group_gt<TPI, L>(u, v);
        const bool hu = cu || (cs && gt);              // u is the one that gets halved
        uint32_t X[L], Y[L], Pp[L], Qq[L];
#pragma unroll
        for (int j = 0; j < L; j++) { X[j] = hu ? u[j] : v[j]; Y[j] = hu ? v[j] : u[j]; Pp[j] = hu ? r[j] : s[j]; Qq[j] = hu ? s[j] : r[j]; }
        const uint32_t ms = cs ? 0xffffffffu : 0u;
        (void)group_sub_masked<TPI, L>(X, Y, ms, cs ? 1u : 0u);      // X -= Y when both odd (X >= Y)
        group_shr1<TPI, L>(X);
        uint32_t Ph = hu ? rh : sh, Qh = hu ? sh : rh;
        Ph += (Qh & ms & 1u) + group_add_masked<TPI, L>(Pp, Qq, ms); // P += Q when both odd
        const uint32_t qtop = __shfl_sync(FULL, Qq[L - 1] >> 31, TPI - 1, TPI);
        group_shl1<TPI, L>(Qq);
        Qh = (Qh << 1) | qtop;
        if (active) {
#pragma unroll
            for (int j = 0; j < L; j++) {
                if (hu) { u[j] = X[j]; r[j] = Pp[j]; s[j] = Qq[j]; }
                else { v[j] = X[j]; s[j] = Pp[j]; r[j] = Qq[j]; }
            }
            if (hu) { rh = Ph; sh = Qh; } else { sh = Ph; rh = Qh; }
            k++;
        }
    }
    // both predicates contain warp ballots: evaluate them unconditionally (no short-circuit), the
    // groups of a warp disagree on the outcome
    const bool u_one = group_is_one<TPI, L>(u);
    const bool v_zero = group_is_zero<TPI, L>(v);
    const bool ok = u_one & v_zero;
    // r < 2n (rh is its bit above the K limbs): bring below n, then x = n - r == a^-1 * 2^k (mod n)
    {
        uint32_t D[L];
#pragma unroll
        for (int j = 0; j < L; j++) D[j] = r[j];
        uint32_t ge = group_sub_masked<TPI, L>(D, m.n, 0xffffffffu, 1u);
        if (rh | ge) {
#pragma unroll
            for (int j = 0; j < L; j++) r[j] = D[j];
        }
    }
    uint32_t x[L];
#pragma unroll
    for (int j = 0; j < L; j++) x[j] = m.n[j];
    (void)group_sub_masked<TPI, L>(x, r, 0xffffffffu, 1u);           // n - r  (r <= n)
    cond_sub<TPI, L>(x, m.n);                                        // r == 0 -> x == n -> 0
    // remove 2^k with two Montgomery products by powers of two: x * 2^j1 / R * 2^j2 / R with
    // j1 + j2 = 2*MBITS - k.  k differs between the groups of a warp, so this tail must be
    // branch-free (every lane runs the same shuffles): both products are always executed.
    int jj = 2 * MBITS - k;
    if (jj > 2 * MBITS - 2) jj = 2 * MBITS - 2;            // k < 2 only for a == 0 / n == 1 (no inverse, result unused)
    const int j1 = jj < MBITS ? jj : MBITS - 1;
    const int j2 = jj - j1;
    uint32_t p[L];
#pragma unroll
    for (int j = 0; j < L; j++) p[j] = ((j1 >> 5) == gl * L + j) ? (1u << (j1 & 31)) : 0u;
    mont_mul<TPI, L>(x, x, p, m.n, m.n0inv);
#pragma unroll
    for (int j = 0; j < L; j++) p[j] = ((j2 >> 5) == gl * L + j) ? (1u << (j2 & 31)) : 0u;
    mont_mul<TPI, L>(out, x, p, m.n, m.n0inv);
    return ok;
}

struct InvClass {
    Operand mod;
    Operand in;
    uint32_t* out;
    uint8_t* ok;            // 1 = inverse exists (per instance, ok_stride bytes apart)
    uint32_t ok_stride;
    uint32_t out_stride;
    Operand nadic;          // nadic_inv_kernel only (nadic_inv.cuh): inverse modulo N^2, `mod` names N and this the key's constants row
    int count;
    int item_begin;
};
struct InvLaunch {
    InvClass cls[16];
    int n_classes;
    int total_items;
};

template <int K, int TPI>
__global__ void __launch_bounds__(128)
inv_jobs_kernel(const InvLaunch* __restrict__ launch, unsigned int* __restrict__ counter) {
    constexpr int L = K / TPI;
    constexpr int GPW = 32 / TPI;
    const int lane = threadIdx.x & 31;
    const int total = launch->total_items;
    while (true) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if ((int)item >= total) break;
        int ci = 0;
        while (ci + 1 < launch->n_classes && launch->cls[ci + 1].item_begin <= (int)item) ci++;
        const InvClass& c = launch->cls[ci];
        const int g = ((int)item - c.item_begin) * GPW + lane / TPI;
        const bool live = g < c.count;
        const int i = live ? g : c.count - 1;
        MontCtx<L> m;
        load_operand<TPI, L>(m.n, c.mod, i);
        mont_setup<TPI, L>(m);
        uint32_t a[L], o[L];
        load_operand<TPI, L>(a, c.in, i);
        bool ok = group_modinv<TPI, L>(o, a, m);
        if (live) {
            if (!ok) {
#pragma unroll
                for (int j = 0; j < L; j++) o[j] = 0;
            }
            store_limbs<TPI, L>(c.out + (size_t)g * c.out_stride, o);
            if ((lane & (TPI - 1)) == 0 && c.ok) c.ok[(size_t)g * c.ok_stride] = ok ? 1 : 0;
        }
        __syncwarp();
    }
}

}  // namespace tecdsa
