// Probe for the next-round design in DESIGN.md section 8 / tools/tc_montgomery_study.py: checks on the GPU that
// mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 has the fragment layout the design assumes (rows of A = jobs, i.e. the
// lane groups of 4 threads; a thread's A registers are 32-bit limbs = 4 consecutive u8 digits; a thread's D registers are
// its own job's output columns), and measures its issue rate next to IMAD.WIDE.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/probes/mma_u8_probe.cu && gpurun_out/mma_probe
// Not part of the library (build() only compiles multi-party-ecdsa_b200/csrc).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_u8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A: 16 x 32 u8 row-major, B: 32 x 8 u8 (B[k][n]), D: 16 x 8 s32.  Assumed layout (PTX ISA, m16n8k32 .u8):
//   g = lane >> 2, t = lane & 3
//   a0: A[g][4t..4t+3]      a1: A[g+8][4t..4t+3]     a2: A[g][16+4t..]     a3: A[g+8][16+4t..]
//   b0: B[4t..4t+3][g]      b1: B[16+4t..][g]
//   d0: D[g][2t]  d1: D[g][2t+1]  d2: D[g+8][2t]  d3: D[g+8][2t+1]
__global__ void probe(const uint8_t* A, const uint8_t* B, int* D) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint32_t a[4], b[2];
    auto pack_a = [&](int row, int col) {
        const uint8_t* p = A + row * 32 + col;
        return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    };
    auto pack_b = [&](int k, int n) {
        return (uint32_t)B[k * 8 + n] | ((uint32_t)B[(k + 1) * 8 + n] << 8) | ((uint32_t)B[(k + 2) * 8 + n] << 16) | ((uint32_t)B[(k + 3) * 8 + n] << 24);
    };
    a[0] = pack_a(g, 4 * t); a[1] = pack_a(g + 8, 4 * t); a[2] = pack_a(g, 16 + 4 * t); a[3] = pack_a(g + 8, 16 + 4 * t);
    b[0] = pack_b(4 * t, g); b[1] = pack_b(16 + 4 * t, g);
    int d[4] = {0, 0, 0, 0};
    mma_u8(d, a, b);
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

// issue-rate comparison: `iters` dependent-free MMAs per warp vs the same count of IMAD.WIDE
__global__ void rate_mma(int* sink, int iters) {
    uint32_t a[4] = {threadIdx.x * 2654435761u, 0x01020304u, blockIdx.x + 7u, 0x0a0b0c0du}, b[2] = {0x11223344u, threadIdx.x + 1u};
    int d0[4] = {0, 0, 0, 0}, d1[4] = {1, 1, 1, 1}, d2[4] = {2, 2, 2, 2}, d3[4] = {3, 3, 3, 3};
    for (int i = 0; i < iters; i++) { mma_u8(d0, a, b); mma_u8(d1, a, b); mma_u8(d2, a, b); mma_u8(d3, a, b); }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = d0[0] + d1[1] + d2[2] + d3[3];
}
__global__ void rate_imad(unsigned long long* sink, int iters) {
    unsigned long long acc[8];
    uint32_t x = threadIdx.x * 2654435761u + 1u, y = blockIdx.x * 40503u + 3u;
    for (int j = 0; j < 8; j++) acc[j] = j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] += (unsigned long long)(x + j) * (y + i);
    }
    unsigned long long s = 0;
    for (int j = 0; j < 8; j++) s ^= acc[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    uint8_t hA[16 * 32], hB[32 * 8];
    int hD[16 * 8], want[16 * 8];
    srand(7);
    for (auto& v : hA) v = (uint8_t)(rand() & 255);
    for (auto& v : hB) v = (uint8_t)(rand() & 255);
    for (int r = 0; r < 16; r++) for (int n = 0; n < 8; n++) { int s = 0; for (int k = 0; k < 32; k++) s += hA[r * 32 + k] * hB[k * 8 + n]; want[r * 8 + n] = s; }
    uint8_t *dA, *dB; int* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(dA, dB, dD);
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128; i++) bad += hD[i] != want[i];
    printf("layout %s (%d mismatches), %s\n", bad ? "MISMATCH" : "ok", bad, cudaGetErrorString(cudaGetLastError()));

    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    int* s1; unsigned long long* s2;
    cudaMalloc(&s1, (size_t)blocks * threads * 4); cudaMalloc(&s2, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    rate_mma<<<blocks, threads>>>(s1, 100);
    cudaEventRecord(e0); rate_mma<<<blocks, threads>>>(s1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * iters * 4;
    printf("mma.sync m16n8k32 u8: %.1f G MMA/s = %.2f P u8-MAC/s (%.1f per clk per SM at 1.9 GHz)\n", mmas / ms / 1e6, mmas * 4096 / ms / 1e12,
           mmas / (ms * 1e-3) / sms / 1.9e9);
    rate_imad<<<blocks, threads>>>(s2, 100);
    cudaEventRecord(e0); rate_imad<<<blocks, threads>>>(s2, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)blocks * threads * iters * 8;
    printf("IMAD.WIDE: %.2f T MAC32/s\n", macs / ms / 1e9);
    return bad != 0;
}
