// Micro-benchmark (round 1): throughput of the exponential variants the item-attention softmax could use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
    for (int i = 0; i < iters; ++i) {
#define OP(r)                                                                          \
        if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r));            \
        if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));              \
        if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r));         \
        if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(r));               \
        if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r));             \
        if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r));            \
        if (MODE == 6) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(r));              \
        if (MODE == 7) asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(r) : "r"(a0), "r"(a1)); \
        if (MODE == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(r) : "r"(a1)); \
        if (MODE == 9) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r) : "r"(a0), "r"(a1)); \
        if (MODE == 10) asm volatile("shf.l.wrap.b32 %0, %0, 23, %0;\n\tadd.s32 %0, %0, %1;" : "+r"(r) : "r"(a1));
        OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
// packed fp32x2 forms (64-bit register pairs), 8 independent chains
template <int MODE>
__global__ void k2(uint64_t* out, int iters, uint32_t seed) {
    uint64_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = (uint64_t)(seed + threadIdx.x) * (2 * i + 3);
    const uint64_t c = a[0] ^ 0x3f8000003f800000ull;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[q]) : "l"(c));
            if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[q]) : "l"(c));
            if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[q]) : "l"(c));
        }
    }
    uint64_t r = 0;
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run2(const char* name) {
    uint64_t* out;
    cudaMalloc(&out, 148 * 8 * 256 * 8);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k2<MODE><<<148 * 8, 256>>>(out, 16, 1);
    cudaEventRecord(e0);
    k2<MODE><<<148 * 8, 256>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_clk_sm = ops / (ms * 1e-3) / (clk * 1e3) / 148.0;
    printf("%-28s %8.3f ms  %7.2f instr-lanes/clk/SM (at %d MHz nominal)  => %7.2f results/clk/SM\n", name, ms, per_clk_sm, clk / 1000,
           per_clk_sm * 2);
    cudaFree(out);
}
template <int MODE>
void run(const char* name, int results_per_op) {
    uint32_t* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 16, 1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_clk_sm = ops / (ms * 1e-3) / (clk * 1e3) / 148.0;
    printf("%-28s %8.3f ms  %7.2f instr-lanes/clk/SM (at %d MHz nominal)  => %7.2f results/clk/SM\n", name, ms, per_clk_sm, clk / 1000,
           per_clk_sm * results_per_op);
    cudaFree(out);
}
int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("tanh.approx.f32", 1);
    run<4>("tanh.approx.f16x2", 2);
    run<5>("fma.rn.f32", 1);
    run<6>("fma.rn.f16x2", 2);
    run<7>("max.f32 (3 inputs)", 1);
    run<8>("cvt.rn.bf16x2.f32", 1);
    run<9>("fma.rn.f32 (3 registers)", 1);
    run<10>("shl 23 + add.s32", 1);
    run2<0>("fma.rn.f32x2");
    run2<1>("add.rn.f32x2");
    run2<2>("mul.rn.f32x2");
    return 0;
}
