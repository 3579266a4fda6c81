// Micro-benchmark: how many FMA-pipe / ALU-pipe instructions issue "for free" next to one MUFU.EX2 per lane?
// Each thread runs 8 independent chains; per iteration every chain does 1 ex2 + N instructions of one other kind.
// Reports SMSP cycles per (1 ex2 + N ops) warp-instruction group (MUFU alone = 8 cycles: 4 lanes / clk / SMSP).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/coissue_bench tools/coissue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND, int N, int MUFU>
__global__ void k(uint64_t* out, int iters, uint32_t seed) {
    uint32_t e[8];
    uint64_t a[8];
    for (int i = 0; i < 8; ++i) { e[i] = (seed + threadIdx.x) * (2 * i + 3); a[i] = (uint64_t)e[i] * 0x100000001ull; }
    const uint64_t c = a[0] ^ 0x3f8000003f800000ull;
    const uint32_t c32 = (uint32_t)c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(e[q]));
#pragma unroll
            for (int n = 0; n < N; ++n) {
                uint32_t lo = (uint32_t)a[q];
                if (KIND == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[q]) : "l"(c));
                if (KIND == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
                if (KIND == 2) { asm volatile("max.f32 %0, %0, %1, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
                if (KIND == 3) { asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
                if (KIND == 4) { asm volatile("add.s32 %0, %0, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
                if (KIND == 5) { asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
                if (KIND == 6) { asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(lo) : "r"(c32)); a[q] = (a[q] & 0xffffffff00000000ull) | lo; }
            }
        }
    }
    uint64_t r = 0;
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int KIND, int N, int MUFU>
void run(const char* name) {
    uint64_t* out;
    cudaMalloc(&out, 148 * 8 * 256 * 8);
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, N, MUFU><<<148 * 8, 256>>>(out, 16, 1);
    cudaEventRecord(e0);
    k<KIND, N, MUFU><<<148 * 8, 256>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = 148.0 * 8 * 8 * iters * 8;  // warp-level groups (8 warps per block)
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = (ms * 1e-3) * (clk * 1e3) * 148 * 4 / groups;
    printf("%d ex2 + %d x %-22s %7.3f ms  %6.2f SMSP cycles per group\n", MUFU, N, name, ms, cyc);
    cudaFree(out);
}
#define ROW(KIND, NAME) run<KIND, 1, 1>(NAME); run<KIND, 2, 1>(NAME); run<KIND, 3, 1>(NAME); run<KIND, 4, 1>(NAME); run<KIND, 6, 1>(NAME); run<KIND, 8, 1>(NAME); run<KIND, 4, 0>(NAME);
int main() {
    ROW(0, "fma.rn.f32x2") ROW(1, "fma.rn.f32") ROW(2, "max.f32") ROW(3, "cvt.rn.bf16x2.f32") ROW(4, "add.s32") ROW(5, "fma.rn.f16x2") ROW(6, "mad.lo.s32")
    return 0;
}
