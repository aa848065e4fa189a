// Inverses modulo EVEN moduli (phi(N), p - 1) from the odd-modulus inversion job list:
//     a^-1 mod m = (1 + m*t) / a,   t = -(m^-1) mod a          (a odd, gcd(a, m) = 1; the division is exact)
// so one Kaliski job modulo the odd number a serves `BigInt::mod_inv(&xhi, &phi)` (/root/reference/src/protocols/multi_party_ecdsa/
// gg_2020/party_i.rs:146), N^-1 mod phi(N) of zk-paillier's correct-key proof and of `Paillier::open` (gg_2020/blame.rs:252-256).
// Included by .cu files only (kernels in an anonymous namespace).
#pragma once
#include "st_bigint.cuh"
#include "../../include/tecdsa_b200.h"

namespace {

// N = p*q, phi = (p-1)(q-1) per instance
__global__ void k_kp_pre(const uint32_t* p32, const uint32_t* q32, uint32_t* n64, uint32_t* phi64, int count) {
    using namespace tecdsa;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t *p = p32 + (size_t)i * 32, *q = q32 + (size_t)i * 32;
    st::mul(n64 + (size_t)i * 64, p, 32, q, 32);
    uint32_t pm[32], qm[32], one[32];
    st::zero(one, 32); one[0] = 1;
    st::sub(pm, p, one, 32); st::sub(qm, q, one, 32);
    st::mul(phi64 + (size_t)i * 64, pm, 32, qm, 32);
}
// x = a^-1 mod m from minv = (m mod a)^-1 mod a (ok flag of the inversion job): x = ((1 + m*(a - minv)) / a), exact division
// by the odd a = multiplication by a^-1 mod 2^2048 (Newton lifting).  neg != 0 writes m - x instead (generate_h1_h2_N_tilde
// hands back the negated exponents, party_i.rs:152-153).  ok == 0 (or an even a): x = 0, status NOT_INVERTIBLE.
__global__ void k_inv_even_post(const uint32_t* a64, const uint32_t* m64, const uint32_t* minv64, const uint8_t* ok, uint32_t* x64, int neg,
                                uint8_t* status, int count) {
    using namespace tecdsa;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t *a = a64 + (size_t)i * 64, *m = m64 + (size_t)i * 64, *mi = minv64 + (size_t)i * 64;
    uint32_t* x = x64 + (size_t)i * 64;
    if (!ok[i] || !(a[0] & 1u)) {
        st::zero(x, 64);
        if (status && status[i] == 0) status[i] = TECDSA_ST_NOT_INVERTIBLE;
        return;
    }
    uint32_t t[64], w[128], ainv[64], t1[64], t2[64], one = 1;
    st::sub(t, a, mi, 64);                                   // t = a - minv  in (0, a)
    st::mul_add(w, 128, m, 64, t, 64, &one, 1);              // 1 + m*t
    st::zero(ainv, 64); ainv[0] = 0u - st::neg_inv32_st(a[0]);
    for (int it = 0; it < 6; it++) {                         // 32 -> 2048 correct bits
        uint32_t two[64]; st::zero(two, 64); two[0] = 2;
        st::mul_low(t1, a, ainv, 64);
        st::sub(t2, two, t1, 64);
        st::mul_low(t1, ainv, t2, 64);
        st::copy(ainv, t1, 64);
    }
    st::mul_low(t1, w, ainv, 64);                            // x < m < 2^2048: the low half decides
    if (neg) st::sub(x, m, t1, 64); else st::copy(x, t1, 64);
}

}  // namespace
