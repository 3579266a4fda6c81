// Micro-benchmark 2: do the FMA pipe (FFMA2), the ALU pipe (FMNMX / F2FP / IADD) and the MUFU pipe overlap with EACH OTHER?
// per iteration and chain: MUFU x ex2, NF x fma.rn.f32x2, NM x 3-input max.f32, NC x cvt.rn.bf16x2.f32, NS x fma.rn.f32
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/coissue2_bench tools/coissue2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MUFU, int NF, int NM, int NC, int NS>
__global__ void k(uint64_t* out, int iters, uint32_t seed) {
    uint32_t e[8], m[8], cv[8], fs[8];
    uint64_t a[8];
    for (int i = 0; i < 8; ++i) {
        e[i] = (seed + threadIdx.x) * (2 * i + 3); a[i] = (uint64_t)e[i] * 0x100000001ull;
        m[i] = e[i] * 7; cv[i] = e[i] * 11; fs[i] = e[i] * 13;
    }
    const uint64_t c = a[0] ^ 0x3f8000003f800000ull;
    const uint32_t c32 = (uint32_t)c, d32 = c32 * 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int n = 0; n < MUFU; ++n) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(e[q]));
#pragma unroll
            for (int n = 0; n < NF; ++n) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[q]) : "l"(c));
#pragma unroll
            for (int n = 0; n < NM; ++n) asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(m[q]) : "r"(c32), "r"(d32));
#pragma unroll
            for (int n = 0; n < NC; ++n) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(cv[q]) : "r"(c32));
#pragma unroll
            for (int n = 0; n < NS; ++n) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(fs[q]) : "r"(c32), "r"(d32));
        }
    }
    uint64_t r = 0;
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ e[i] ^ m[i] ^ cv[i] ^ fs[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MUFU, int NF, int NM, int NC, int NS>
void run() {
    uint64_t* out;
    cudaMalloc(&out, 148 * 8 * 256 * 8);
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MUFU, NF, NM, NC, NS><<<148 * 8, 256>>>(out, 16, 1);
    cudaEventRecord(e0);
    k<MUFU, NF, NM, NC, NS><<<148 * 8, 256>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = 148.0 * 8 * 8 * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = (ms * 1e-3) * (clk * 1e3) * 148 * 4 / groups;
    printf("%d ex2 + %d FFMA2 + %d FMNMX3 + %d F2FP + %d FFMA : %6.2f SMSP cycles per group (XU %d, FMA pipe %d, ALU pipe %d, issue %d)\n",
           MUFU, NF, NM, NC, NS, cyc, 8 * MUFU, 2 * NF + NS, 2 * (NM + NC), MUFU + NF + NM + NC + NS);
    cudaFree(out);
}
int main() {
    run<0, 4, 4, 0, 0>(); run<0, 4, 2, 2, 0>(); run<0, 2, 2, 2, 0>(); run<0, 4, 0, 0, 4>();
    run<1, 2, 2, 0, 0>(); run<1, 3, 3, 0, 0>(); run<1, 4, 4, 0, 0>(); run<1, 2, 2, 2, 0>(); run<1, 3, 2, 2, 0>(); run<1, 4, 2, 2, 0>();
    run<1, 1, 1, 1, 0>(); run<1, 2, 1, 1, 0>(); run<1, 3, 1, 1, 0>(); run<1, 4, 1, 1, 0>();
    run<1, 2, 1, 1, 2>(); run<1, 2, 1, 1, 4>(); run<2, 4, 2, 2, 0>(); run<2, 8, 2, 2, 0>(); run<2, 8, 4, 4, 0>();
    return 0;
}
