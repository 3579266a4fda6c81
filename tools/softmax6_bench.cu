// Micro-benchmark (round 2): what the item-attention softmax inner loop could cost per warp-element if
//   ROWSUM = 0  the row sum left the softmax threads (P x ones on the tensor core),
//   HALF   = 1  the MUFU share ran as cvt.rn.f16x2.f32 + ex2.approx.f16x2.  ptxas turns that into
//               F2FP + 2 x MUFU.EX2.F16 + PRMT (profiles/r2_f16x2_sass.txt): four issue slots per pair, not two.
// Scores are re-read from shared memory every tile (16 x LDS.128, the stand-in for tcgen05.ld) so nothing is hoisted
// and no extra arithmetic is added to the loop (r1's loop paid one FADD2 per pair for that).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/softmax6_bench tools/softmax6_bench.cu
#include <cstdio>
#include "../npe_pfn_b200/csrc/attn_tc.cuh"
namespace pfn { std::string& last_error() { static std::string s; return s; } }
using namespace pfn;

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
    uint32_t h;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
    return h;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t h) {
    uint32_t p;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(p) : "r"(h));
    return p;
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// K of every 16 pairs on the FMA pipes (degree DEG); ROWSUM: packed fp32 row sum in the loop; HALF: f16 outputs
template <int K, int DEG, int ROWSUM, int HALF, int XSRC>
__global__ void __launch_bounds__(128) bench6(float* out, int tiles) {
    __shared__ uint4 src[128 * 16];  // 64 scores per thread
    __shared__ uint4 sink[128 * 2];
    for (int i = 0; i < 16; ++i) {
        uint4 v;
        v.x = __float_as_uint(-0.37f * (float)((i * 28 + threadIdx.x) % 61));
        v.y = __float_as_uint(-0.37f * (float)((i * 28 + 7 + threadIdx.x) % 61));
        v.z = __float_as_uint(-0.37f * (float)((i * 28 + 14 + threadIdx.x) % 61));
        v.w = __float_as_uint(-0.37f * (float)((i * 28 + 21 + threadIdx.x) % 61));
        src[i * 128 + threadIdx.x] = v;
    }
    __syncthreads();
    uint64_t l2 = 0;
    uint32_t ovf = 0;
    uint32_t s[64];
    const uint64_t CM = pk2(kExpMagic, kExpMagic), NEG1 = pk2(-1.f, -1.f);
    for (int j = 0; j < tiles; ++j) {
        const uint32_t sp = smem_u32(src + threadIdx.x);
        if (XSRC == 1 || j == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(s[4 * i]), "=r"(s[4 * i + 1]), "=r"(s[4 * i + 2]), "=r"(s[4 * i + 3]) : "r"(sp + i * 2048));
        }
        const float dj = 1e-7f * (float)j;
        const uint64_t D2 = pk2(dj, dj);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const bool poly = ((i % 16 + 1) * K) / 16 != ((i % 16) * K) / 16;
            float x0 = __uint_as_float(s[2 * i]), x1 = __uint_as_float(s[2 * i + 1]);
            if (XSRC == 0) upk2(fadd2(pk2(x0, x1), D2), x0, x1);  // r1's loop-variant offset: one FADD2 per pair
            float p0 = 0.f, p1 = 0.f;
            if (!poly) {
                if (HALF) {
                    pk[i] = ex2_f16x2(cvt_f16x2(x0, x1));
                } else {
                    p0 = fast_exp2(x0);
                    p1 = fast_exp2(x1);
                    pk[i] = pack_bf16x2(p0, p1);
                }
            } else {
                const uint64_t XC = pk2(fmaxf(x0, HALF ? -14.f : -125.f), fmaxf(x1, HALF ? -14.f : -125.f));
                const uint64_t t2 = fadd2(XC, CM);
                const uint64_t f2 = fadd2(XC, ffma2(t2, NEG1, CM));
                float t0, t1;
                upk2(t2, t0, t1);
                if (HALF) {
                    float f0, f1;
                    upk2(f2, f0, f1);
                    const uint32_t fh = cvt_f16x2(f0, f1);
                    uint32_t q;
                    if (DEG == 2) {
                        q = hfma2(0x33A133A1u, fh, 0x39A139A1u);  // 0.2384, 0.7034
                        q = hfma2(q, fh, 0x3C003C00u);            // 1.0004 ~ 1.0
                    } else {
                        q = hfma2(0x2B102B10u, fh, 0x33C333C3u);  // 0.05517, 0.2426
                        q = hfma2(q, fh, 0x398C398Cu);            // 0.6933
                        q = hfma2(q, fh, 0x3C003C00u);
                    }
                    uint32_t w;
                    asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(w) : "r"(__float_as_uint(t0)), "r"(__float_as_uint(t1)));
                    pk[i] = q + ((w & 0x003F003Fu) << 10);
                } else {
                    uint64_t q2;
                    if (DEG == 2) {
                        q2 = ffma2(pk2(0.23842893540859222f, 0.23842893540859222f), f2, pk2(0.7034479975700378f, 0.7034479975700378f));
                        q2 = ffma2(q2, f2, pk2(1.0004431009292603f, 1.0004431009292603f));
                    } else {
                        q2 = ffma2(pk2(0.05517163127660751f, 0.05517163127660751f), f2, pk2(0.2426111251115799f, 0.2426111251115799f));
                        q2 = ffma2(q2, f2, pk2(0.6932609677314758f, 0.6932609677314758f));
                        q2 = ffma2(q2, f2, pk2(0.9999280571937561f, 0.9999280571937561f));
                    }
                    float q0, q1;
                    upk2(q2, q0, q1);
                    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
                    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
                    pk[i] = pack_bf16x2(p0, p1);
                }
            }
            if (ROWSUM && !HALF) l2 = fadd2(l2, pk2(p0, p1));
            if (HALF) { if (i & 1) asm("max.u16x2 %0, %0, %1;" : "+r"(ovf) : "r"(pk[i] | pk[i - 1])); }  // LOP3 + VIMNMX per two pairs
            else if (poly || !ROWSUM) ovf |= pk[i];
        }
        const uint32_t dp = smem_u32(sink + threadIdx.x);
#pragma unroll
        for (int q = 0; q < 2; ++q)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dp + q * 2048),
                         "r"(pk[16 * q + 0] ^ pk[16 * q + 4] ^ pk[16 * q + 8] ^ pk[16 * q + 12]),
                         "r"(pk[16 * q + 1] ^ pk[16 * q + 5] ^ pk[16 * q + 9] ^ pk[16 * q + 13]),
                         "r"(pk[16 * q + 2] ^ pk[16 * q + 6] ^ pk[16 * q + 10] ^ pk[16 * q + 14]),
                         "r"(pk[16 * q + 3] ^ pk[16 * q + 7] ^ pk[16 * q + 11] ^ pk[16 * q + 15]) : "memory");
    }
    float l0, l1;
    upk2(l2, l0, l1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + __uint_as_float(sink[threadIdx.x].x) + (float)(ovf & 0x40004000u);
}

template <int K, int DEG, int ROWSUM, int HALF, int XSRC = 0>
void run(int ctas_per_sm) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 128 * 4);
    const int tiles = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench6<K, DEG, ROWSUM, HALF, XSRC><<<148 * ctas_per_sm, 128>>>(out, 16);
    cudaEventRecord(e0);
    bench6<K, DEG, ROWSUM, HALF, XSRC><<<148 * ctas_per_sm, 128>>>(out, tiles);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double elems = 148.0 * ctas_per_sm * 128 * (double)tiles * 64;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_clk_sm = elems / (ms * 1e-3) / (clk * 1e3) / 148.0;
    printf("%s src=%s rowsum=%d K=%2d deg=%d warps/SMSP=%d  %7.3f ms  %6.2f elems/clk/SM = %5.2f clk per warp-elem per SMSP => %6.1f TFLOP/s equivalent\n",
           HALF ? "f16x2" : "f32  ", XSRC ? "lds  " : "fadd2", ROWSUM, K, DEG, ctas_per_sm, ms, per_clk_sm, 128.0 / per_clk_sm, elems / (ms * 1e-3) * 128 / 1e12);
    cudaFree(out);
}

int main() {
    for (int w : {3, 4}) {
        run<0, 3, 1, 0>(w); run<4, 3, 1, 0>(w); run<5, 3, 1, 0>(w); run<6, 3, 1, 0>(w);            // v5 form (r1's lean loop)
        run<0, 3, 0, 0>(w); run<4, 3, 0, 0>(w); run<5, 3, 0, 0>(w); run<6, 3, 0, 0>(w); run<7, 3, 0, 0>(w); run<8, 3, 0, 0>(w);
        run<6, 2, 0, 0>(w); run<7, 2, 0, 0>(w); run<8, 2, 0, 0>(w);
        run<0, 3, 0, 1>(w); run<4, 3, 0, 1>(w); run<6, 3, 0, 1>(w); run<6, 2, 0, 1>(w); run<8, 2, 0, 1>(w);
        run<5, 3, 1, 0, 1>(w); run<5, 3, 0, 0, 1>(w); run<6, 3, 0, 0, 1>(w); run<7, 3, 0, 0, 1>(w);  // scores from shared memory
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
