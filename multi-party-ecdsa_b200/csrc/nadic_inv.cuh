// Modular inverse modulo a square N^2 from an inverse modulo N plus one Hensel (Newton) step:
//
//     y0 = (c mod N)^-1 mod N,      c^-1 mod N^2 = y0 * (2 - c*y0) mod N^2
//
// (c*y0 = 1 + kN  =>  c*y0*(2 - c*y0) = 1 - k^2 N^2).  c is invertible modulo N^2 exactly when it is modulo N, so the
// `ok` flag is the one of the K-limb inversion.  The Kaliski almost-inverse is quadratic in the operand width: the
// 2048-bit inversion plus five N-adic products costs about a third of the 4096-bit inversion it replaces
// (`BigInt::mod_inv(c, NN)` at /root/reference/src/utilities/mta/range_proofs.rs:135 and
// src/utilities/zk_pdl_with_slack/mod.rs:192).  Same job descriptors as inv_jobs_kernel, with `mod` naming N and
// `nadic` the constants row of the key (nadic.cuh).
#pragma once
#include "modinv.cuh"
#include "nadic.cuh"

namespace tecdsa {

// (A - B) mod N^2 in digits
template <int TPI, int L>
__device__ __forceinline__ void dig_sub(Dig<L>& Z, const Dig<L>& A, const Dig<L>& B, const uint32_t (&n)[L]) {
    uint32_t T[L];
#pragma unroll
    for (int j = 0; j < L; j++) T[j] = A.d0[j];
    const uint32_t ge = group_sub_masked<TPI, L>(T, B.d0, 0xffffffffu, 1u);      // carry out == (A0 >= B0)
    (void)group_add_masked<TPI, L>(T, n, ge ? 0u : 0xffffffffu);                   // borrowed: + N, and one less in the high digit
    uint32_t hi[L], bor[L];
    mod_sub<TPI, L>(hi, A.d1, B.d1, n);
#pragma unroll
    for (int j = 0; j < L; j++) bor[j] = 0;
    if (group_lane<TPI>() == 0) bor[0] = ge ? 0u : 1u;
    mod_sub<TPI, L>(hi, hi, bor, n);
#pragma unroll
    for (int j = 0; j < L; j++) { Z.d0[j] = T[j]; Z.d1[j] = hi[j]; }
}

template <int K, int TPI>
__global__ void __launch_bounds__(128)
nadic_inv_kernel(const InvLaunch* __restrict__ launch, unsigned int* __restrict__ counter) {
    constexpr int L = K / TPI;
    constexpr int GPW = 32 / TPI;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (TPI - 1);
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
        m.n0inv = neg_inv32(__shfl_sync(FULL, m.n[0], 0, TPI));
        const uint32_t* consts = operand_at(c.nadic, i);
        load_limbs<TPI, L>(m.one, consts + NADIC_ONE * K);          // R mod N  (low digit of R mod N^2)
        load_limbs<TPI, L>(m.rr, consts + NADIC_RR2 * K);           // R^2 mod N

        Dig<L> X, Y, W;
        to_nadic<TPI, L>(X, c.in, i, consts, m.n, m.n0inv);         // c * R mod N^2; its low digit is (c mod N) * R mod N
        uint32_t a[L], y0[L];
#pragma unroll
        for (int j = 0; j < L; j++) a[j] = 0;
        if (gl == 0) a[0] = 1;
        mont_mul<TPI, L>(a, X.d0, a, m.n, m.n0inv);                 // c mod N
        const bool ok = group_modinv<TPI, L>(y0, a, m);
#pragma unroll
        for (int j = 0; j < L; j++) { W.d0[j] = y0[j]; W.d1[j] = 0; }
        load_dig<TPI, L>(Y, consts + NADIC_RR2 * K);
        nadic_mul<TPI, L>(Y, W, Y, false, m.n, m.n0inv);            // y0 * R
        nadic_mul<TPI, L>(X, X, Y, true, m.n, m.n0inv);             // c * y0 * R
        load_dig<TPI, L>(W, consts + NADIC_ONE * K);
        dig_add<TPI, L>(W, W, W, m.n);                              // 2R
        dig_sub<TPI, L>(W, W, X, m.n);                              // (2 - c*y0) * R
        nadic_mul<TPI, L>(Y, Y, W, true, m.n, m.n0inv);             // y0 * (2 - c*y0) * R
#pragma unroll
        for (int j = 0; j < L; j++) { W.d0[j] = 0; W.d1[j] = 0; }
        if (gl == 0) W.d0[0] = 1;
        nadic_mul<TPI, L>(Y, Y, W, true, m.n, m.n0inv);             // leave the Montgomery domain
        uint32_t lo[L], hi[L];
        group_mul_wide<TPI, L>(lo, hi, Y.d1, m.n);
        const uint32_t cy = group_add_masked<TPI, L>(lo, Y.d0, 0xffffffffu);
#pragma unroll
        for (int j = 0; j < L; j++) a[j] = 0;
        if (gl == 0) a[0] = cy;
        (void)group_add_masked<TPI, L>(hi, a, 0xffffffffu);
        if (live) {
            if (!ok) {
#pragma unroll
                for (int j = 0; j < L; j++) { lo[j] = 0; hi[j] = 0; }
            }
            uint32_t* o = c.out + (size_t)g * c.out_stride;
            store_limbs<TPI, L>(o, lo);
            store_limbs<TPI, L>(o + K, hi);
            if (gl == 0 && c.ok) c.ok[(size_t)g * c.ok_stride] = ok ? 1 : 0;
        }
        __syncwarp();
    }
}

}  // namespace tecdsa
