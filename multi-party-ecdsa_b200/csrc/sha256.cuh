// SHA-256 for sm_100a, one thread per digest, with the byte encodings the reference's
// Fiat-Shamir transcripts use: curv `DigestExt::chain_bigint` = big-endian magnitude of the
// integer, minimal length (zero -> one 0x00 byte) [R], `chain_point` = 65-byte uncompressed
// SEC1 [R], and `HashCommitment` (/root/reference/src/protocols/multi_party_ecdsa/gg_2020/
// party_i.rs:577-580).  Transcript orders: src/utilities/mta/range_proofs.rs:143-150,175-182;
// src/utilities/zk_pdl_with_slack/mod.rs:102-110,128-136.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tecdsa {

__device__ __constant__ const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

struct Sha256 {
    uint32_t h[8];
    uint32_t w[16];     // current block, big-endian words
    uint32_t fill;      // bytes in the current block
    uint64_t total;     // bytes absorbed

    __device__ void init() {
        h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
        h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
        for (int i = 0; i < 16; i++) w[i] = 0;
        fill = 0; total = 0;
    }
    __device__ static uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
    __device__ __noinline__ void compress() {
        uint32_t m[16];
        for (int i = 0; i < 16; i++) m[i] = w[i];
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
        for (int i = 0; i < 64; i++) {
            uint32_t wi;
            if (i < 16) wi = m[i];
            else {
                uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
                uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
                wi = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
                m[i & 15] = wi;
            }
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + SHA_K[i] + wi;
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        for (int i = 0; i < 16; i++) w[i] = 0;
        fill = 0;
    }
    __device__ void put(uint8_t byte) {
        w[fill >> 2] |= (uint32_t)byte << (24 - 8 * (fill & 3));
        fill++; total++;
        if (fill == 64) compress();
    }
    __device__ void put_bytes(const uint8_t* p, int n) { for (int i = 0; i < n; i++) put(p[i]); }
    // `chain_bigint`: minimal big-endian magnitude of an n-limb little-endian integer
    __device__ void put_bigint(const uint32_t* limbs, int n) {
        int top = n - 1;
        while (top > 0 && limbs[top] == 0) top--;
        uint32_t tl = limbs[top];
        int nb = tl >> 24 ? 4 : tl >> 16 ? 3 : tl >> 8 ? 2 : 1;      // zero -> one 0x00 byte
        for (int b = nb - 1; b >= 0; b--) put((uint8_t)(tl >> (8 * b)));
        for (int i = top - 1; i >= 0; i--) {
            uint32_t v = limbs[i];
            put((uint8_t)(v >> 24)); put((uint8_t)(v >> 16)); put((uint8_t)(v >> 8)); put((uint8_t)v);
        }
    }
    // fixed-width big-endian field (used by the engine's canonical transcript digest)
    __device__ void put_fixed(const uint32_t* limbs, int n) {
        for (int i = n - 1; i >= 0; i--) {
            uint32_t v = limbs[i];
            put((uint8_t)(v >> 24)); put((uint8_t)(v >> 16)); put((uint8_t)(v >> 8)); put((uint8_t)v);
        }
    }
    // digest as 8 little-endian limbs of the big-endian 256-bit integer (`result_bigint`)
    __device__ void finish(uint32_t* out_limbs) {
        uint64_t bits = total * 8;
        put(0x80);
        while (fill != 56) put(0);
        w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
        compress();
        for (int i = 0; i < 8; i++) out_limbs[i] = h[7 - i];
    }
};

}  // namespace tecdsa
