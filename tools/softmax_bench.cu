// Micro-benchmark: instruction-throughput bound of the item-attention softmax inner loop (attn_tc.cuh::softmax_exp32)
// on registers only -- no TMEM, no barriers, no tensor pipe.  Per "tile" a thread does what a softmax thread does for
// 64 scores: the 3-input maximum pass, P = 2^(S*sc - m) with K of every 16 pairs on the FMA pipes, the packed
// row sum and the bf16 packing; results go to shared memory (stand-in for tcgen05.st).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/softmax_bench tools/softmax_bench.cu
#include <cstdio>
#include "../npe_pfn_b200/csrc/attn_tc.cuh"
namespace pfn { std::string& last_error() { static std::string s; return s; } }
using namespace pfn;

template <int K, int DEG>
__global__ void __launch_bounds__(128) bench(float* out, int tiles, float sc0) {
    __shared__ uint4 sink[128 * 2];
    uint32_t s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = __float_as_uint(-0.37f * (float)((i * 7 + threadIdx.x) % 61));
    uint64_t l2 = 0;
    float m_ref = 0.f, acc = 0.f;
    for (int j = 0; j < tiles; ++j) {
        const float sc = sc0 + 1e-7f * (float)j;  // loop-variant: nothing can be hoisted
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx0) : "f"(__uint_as_float(s[2 * i])), "f"(__uint_as_float(s[2 * i + 1])));
            asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(mx1) : "f"(__uint_as_float(s[32 + 2 * i])), "f"(__uint_as_float(s[32 + 2 * i + 1])));
        }
        const float mt = fmaxf(mx0, mx1) * sc;
        if (mt > m_ref + 8.0f) m_ref = rintf(mt);
        const float cm = kExpMagic - m_ref, smin = (m_ref - 125.0f) * (1.0f / sc);
        uint32_t pk[32];
        softmax_exp32<K, DEG>(reinterpret_cast<const uint32_t(&)[32]>(s[0]), pk, sc, -m_ref, cm, smin, l2);
        softmax_exp32<K, DEG>(reinterpret_cast<const uint32_t(&)[32]>(s[32]), pk + 16, sc, -m_ref, cm, smin, l2);
        volatile uint4* dst = sink + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 2; ++q) {  // 8 x 16 B per tile like a 32-register tcgen05.st; two slots keep smem small
            dst[q * 128].x = pk[16 * q + 0] ^ pk[16 * q + 4] ^ pk[16 * q + 8] ^ pk[16 * q + 12];
            dst[q * 128].y = pk[16 * q + 1] ^ pk[16 * q + 5] ^ pk[16 * q + 9] ^ pk[16 * q + 13];
            dst[q * 128].z = pk[16 * q + 2] ^ pk[16 * q + 6] ^ pk[16 * q + 10] ^ pk[16 * q + 14];
            dst[q * 128].w = pk[16 * q + 3] ^ pk[16 * q + 7] ^ pk[16 * q + 11] ^ pk[16 * q + 15];
        }
    }
    float l0, l1;
    upk2(l2, l0, l1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + acc + __uint_as_float(sink[threadIdx.x].x);
}

// software-pipelined variant: units of 32 scores; the maximum pass of the NEXT unit is interleaved with the
// exponentials of the current one (one 3-input max per exponential pair)
template <int K, int DEG>
__global__ void __launch_bounds__(128) bench_pipe(float* out, int tiles, float sc0) {
    __shared__ uint4 sink[128 * 2];
    uint32_t s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = __float_as_uint(-0.37f * (float)((i * 7 + threadIdx.x) % 61));
    uint64_t l2 = 0;
    float m_ref = 0.f, mxa = -1.0f, mxb = -2.0f;
    for (int j = 0; j < tiles; ++j) {
        const float sc = sc0 + 1e-7f * (float)j;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float mt = fmaxf(mxa, mxb) * sc;
            if (mt > m_ref + 8.0f) m_ref = rintf(mt);
            const float cm = kExpMagic - m_ref, smin = (m_ref - 125.0f) * (1.0f / sc);
            uint32_t pk[16];
            float na = -INFINITY, nb = -INFINITY;
            const uint32_t* nxt = s + 32 * (u ^ 1);
            softmax_exp32<K, DEG>(reinterpret_cast<const uint32_t(&)[32]>(s[32 * u]), pk, sc, -m_ref, cm, smin, l2, [&](int i) {
                if (i & 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(na) : "f"(__uint_as_float(nxt[2 * i])), "f"(__uint_as_float(nxt[2 * i + 1])));
                else asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(nb) : "f"(__uint_as_float(nxt[2 * i])), "f"(__uint_as_float(nxt[2 * i + 1])));
            });
            mxa = na; mxb = nb;
            volatile uint4* dst = sink + threadIdx.x;
            dst[u * 128].x = pk[0] ^ pk[4] ^ pk[8] ^ pk[12];
            dst[u * 128].y = pk[1] ^ pk[5] ^ pk[9] ^ pk[13];
            dst[u * 128].z = pk[2] ^ pk[6] ^ pk[10] ^ pk[14];
            dst[u * 128].w = pk[3] ^ pk[7] ^ pk[11] ^ pk[15];
        }
    }
    float l0, l1;
    upk2(l2, l0, l1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + __uint_as_float(sink[threadIdx.x].x);
}

// "lean" variant: what the inner loop would cost if the tensor core delivered x = S*sc - m directly (scale folded
// into the Q projection, -m as an extra K = 16 block of the QK^T MMA) and the tile maximum were replaced by an
// overflow check on the packed outputs (OR of all words, one LOP3 per pair).
template <int K, int DEG, int TRUNC>
__global__ void __launch_bounds__(128) bench_lean(float* out, int tiles, float sc0) {
    __shared__ uint4 sink[128 * 2];
    uint32_t s[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] = __float_as_uint(-0.37f * (float)((i * 7 + threadIdx.x) % 61));
    uint64_t l2 = 0;
    uint32_t ovf = 0;
    for (int j = 0; j < tiles; ++j) {
        const float d = 1e-7f * (float)j;  // loop-variant offset so nothing is hoisted (costs one FADD2 per pair here; the real
                                           // kernel would read fresh TMEM values instead)
        const uint64_t D2 = pk2(d, d), CM = pk2(kExpMagic, kExpMagic), NEG1 = pk2(-1.f, -1.f);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const bool poly = ((i % 16 + 1) * K) / 16 != ((i % 16) * K) / 16;
            float p0, p1;
            const uint64_t X2 = fadd2(pk2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), D2);
            if (!poly) {
                float x0, x1;
                upk2(X2, x0, x1);
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
            } else {
                float x0, x1;
                upk2(X2, x0, x1);
                const uint64_t XC = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
                const uint64_t t2 = fadd2(XC, CM);
                const uint64_t r2 = ffma2(t2, NEG1, CM);   // -(n)
                const uint64_t f2 = fadd2(XC, r2);
                uint64_t q2 = ffma2(pk2(0.05517163127660751f, 0.05517163127660751f), f2, pk2(0.2426111251115799f, 0.2426111251115799f));
                q2 = ffma2(q2, f2, pk2(0.6932609677314758f, 0.6932609677314758f));
                q2 = ffma2(q2, f2, pk2(0.9999280571937561f, 0.9999280571937561f));
                float q0, q1, t0, t1;
                upk2(q2, q0, q1);
                upk2(t2, t0, t1);
                p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
                p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
            }
            l2 = fadd2(l2, pk2(p0, p1));
            if (TRUNC) asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[i]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));  // truncating pack
            else pk[i] = pack_bf16x2(p0, p1);
            ovf |= pk[i];
        }
        volatile uint4* dst = sink + threadIdx.x;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            dst[q * 128].x = pk[16 * q + 0] ^ pk[16 * q + 4] ^ pk[16 * q + 8] ^ pk[16 * q + 12];
            dst[q * 128].y = pk[16 * q + 1] ^ pk[16 * q + 5] ^ pk[16 * q + 9] ^ pk[16 * q + 13];
            dst[q * 128].z = pk[16 * q + 2] ^ pk[16 * q + 6] ^ pk[16 * q + 10] ^ pk[16 * q + 14];
            dst[q * 128].w = pk[16 * q + 3] ^ pk[16 * q + 7] ^ pk[16 * q + 11] ^ pk[16 * q + 15];
        }
    }
    float l0, l1;
    upk2(l2, l0, l1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + __uint_as_float(sink[threadIdx.x].x) + (float)(ovf & 0x40004000u);
}

template <int K, int DEG, int PIPE = 0>
void run(int ctas_per_sm) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 128 * 4);
    const int tiles = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&](int n) {
        if (PIPE == 3) bench_lean<K, DEG, 1><<<148 * ctas_per_sm, 128>>>(out, n, 0.255f);
        else if (PIPE == 2) bench_lean<K, DEG, 0><<<148 * ctas_per_sm, 128>>>(out, n, 0.255f);
        else if (PIPE == 1) bench_pipe<K, DEG><<<148 * ctas_per_sm, 128>>>(out, n, 0.255f);
        else bench<K, DEG><<<148 * ctas_per_sm, 128>>>(out, n, 0.255f);
    };
    launch(16);
    cudaEventRecord(e0);
    launch(tiles);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double elems = 148.0 * ctas_per_sm * 128 * (double)tiles * 64;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_clk_sm = elems / (ms * 1e-3) / (clk * 1e3) / 148.0;
    printf("%s K=%2d deg=%d warps/SMSP=%d  %7.3f ms  %6.2f elems/clk/SM  = %5.2f clk per warp-elem per SMSP  => %6.1f TFLOP/s equivalent (128 FLOP/elem)\n",
           PIPE == 3 ? "lean+prmt" : PIPE == 2 ? "lean     " : PIPE ? "pipelined" : "two-pass ", K, DEG, ctas_per_sm, ms, per_clk_sm, 128.0 / per_clk_sm, elems / (ms * 1e-3) * 128 / 1e12);
    cudaFree(out);
}
int main(int argc, char** argv) {
    if (argc > 1) {  // lean-variant sweep only
        for (int w : {3}) {
            run<0, 3, 3>(w); run<3, 3, 3>(w); run<4, 3, 3>(w); run<5, 3, 3>(w); run<6, 3, 3>(w);
        }
        for (int w : {1, 3, 4}) {
            run<0, 3>(w); run<0, 3, 2>(w); run<2, 3, 2>(w); run<3, 3, 2>(w); run<4, 3, 2>(w); run<5, 3, 2>(w); run<6, 3, 2>(w); run<8, 3, 2>(w);
        }
        return cudaDeviceSynchronize() != cudaSuccess;
    }

    for (int w : {1, 2, 3, 4}) { run<0, 3>(w); run<0, 3, 1>(w); }
    for (int w : {1, 3}) {
        run<1, 3, 1>(w); run<2, 3, 1>(w); run<3, 3, 1>(w); run<4, 3, 1>(w); run<5, 3, 1>(w); run<6, 3, 1>(w);
        run<2, 2, 1>(w); run<3, 2, 1>(w); run<4, 2, 1>(w); run<6, 2, 1>(w);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
