// Dense projection GEMMs of the per-feature transformer (QKV / out-proj / MLP / decoder):
//   C[M, N] = A[M, K] (bf16, row stride lda) x W[N, K]^T (bf16, row-major), fp32 accumulate,
// with the epilogues the layer needs fused in (bias+GELU, residual+LayerNorm, bias+scale->fp32).
// This is the warp-level (mma.sync) implementation: 128x192 CTA tile, 64-wide K slices,
// 3-stage cp.async pipeline, XOR-swizzled shared memory read with ldmatrix.
#pragma once
#include "common.cuh"

namespace pfn {

enum GemmEpi { EPI_BF16 = 0, EPI_BIAS_GELU_BF16 = 1, EPI_RESID_LN = 2, EPI_BIAS_SCALE_F32 = 3,
               EPI_FEATURE_ATTN = 4 /* gemm_tc.cuh only: QKV projection + attention between features in the epilogue */ };

struct GemmArgs {
    const bf16* A;
    int64_t lda;
    const bf16* W;  // [N, K]
    int64_t M;
    int N;
    int K;
    bf16* Cb;  // bf16 output (EPI_BF16 / GELU / RESID_LN copy)
    int64_t ldcb;
    float* Cf;  // fp32 output (logits) or fp32 residual stream (in/out, RESID_LN)
    int64_t ldcf;
    const float* bias;
    float scale;
    float ln_eps;
};

constexpr int GM_BM = 128, GM_BN = 192, GM_BK = 64, GM_STAGES = 3, GM_THREADS = 256;
constexpr int GM_STAGE_BYTES = (GM_BM + GM_BN) * GM_BK * 2;
constexpr int GM_SMEM_BYTES = GM_STAGES * GM_STAGE_BYTES;

__device__ __forceinline__ uint32_t gm_swz(int row, int chunk) {  // byte offset inside a [rows][64] bf16 tile
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int EPI>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_mma_kernel(GemmArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * GM_BM;
    const int n0 = blockIdx.y * GM_BN;
    const uint32_t sbase = smem_u32(smem);
    const int ktiles = p.K / GM_BK;

    auto load_stage = [&](int stage, int kt) {
        const uint32_t sa = sbase + stage * GM_STAGE_BYTES;
        const uint32_t sb = sa + GM_BM * GM_BK * 2;
        const int k0 = kt * GM_BK;
#pragma unroll
        for (int i = 0; i < (GM_BM * 8) / GM_THREADS; ++i) {
            int idx = tid + i * GM_THREADS;
            int row = idx >> 3, ch = idx & 7;
            int64_t gr = m0 + row;
            bool ok = gr < p.M;
            const bf16* src = p.A + (ok ? gr : 0) * p.lda + k0 + ch * 8;
            cp_async16(sa + gm_swz(row, ch), src, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < (GM_BN * 8) / GM_THREADS; ++i) {
            int idx = tid + i * GM_THREADS;
            int row = idx >> 3, ch = idx & 7;
            int gn = n0 + row;
            bool ok = gn < p.N;
            const bf16* src = p.W + (int64_t)(ok ? gn : 0) * p.K + k0 + ch * 8;
            cp_async16(sb + gm_swz(row, ch), src, ok ? 16 : 0);
        }
    };

    float acc[24][4];
#pragma unroll
    for (int j = 0; j < 24; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < ktiles) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < ktiles; ++kt) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + GM_STAGES - 1;
            if (nk < ktiles) load_stage(nk % GM_STAGES, nk);
            cp_async_commit();
        }
        const uint32_t sa = sbase + (kt % GM_STAGES) * GM_STAGE_BYTES;
        const uint32_t sb = sa + GM_BM * GM_BK * 2;
#pragma unroll
        for (int kk = 0; kk < GM_BK / 16; ++kk) {
            uint32_t a[4];
            {
                int row = warp * 16 + (lane & 15);
                int ch = 2 * kk + (lane >> 4);
                ldmatrix_x4(a[0], a[1], a[2], a[3], sa + gm_swz(row, ch));
            }
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                uint32_t b0, b1, b2, b3;
                int row = 16 * j + (lane & 7) + ((lane >> 4) << 3);
                int ch = 2 * kk + ((lane >> 3) & 1);
                ldmatrix_x4(b0, b1, b2, b3, sb + gm_swz(row, ch));
                mma_bf16_16816(acc[2 * j], a, b0, b1);
                mma_bf16_16816(acc[2 * j + 1], a, b2, b3);
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----------------------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = m0 + warp * 16 + g, r1 = r0 + 8;
    const bool ok0 = r0 < p.M, ok1 = r1 < p.M;

    if (EPI == EPI_RESID_LN) {
        // N == 192 == GM_BN: each warp owns complete rows.  x = acc + residual; LayerNorm (no affine).
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            int col = 8 * j + 2 * t;
            float2 x0 = ok0 ? *reinterpret_cast<const float2*>(p.Cf + r0 * p.ldcf + col) : make_float2(0.f, 0.f);
            float2 x1 = ok1 ? *reinterpret_cast<const float2*>(p.Cf + r1 * p.ldcf + col) : make_float2(0.f, 0.f);
            acc[j][0] += x0.x; acc[j][1] += x0.y; acc[j][2] += x1.x; acc[j][3] += x1.y;
            s0 += acc[j][0] + acc[j][1];
            s1 += acc[j][2] + acc[j][3];
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float mu0 = s0 * (1.0f / kE), mu1 = s1 * (1.0f / kE);
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            float d;
            d = acc[j][0] - mu0; v0 += d * d;
            d = acc[j][1] - mu0; v0 += d * d;
            d = acc[j][2] - mu1; v1 += d * d;
            d = acc[j][3] - mu1; v1 += d * d;
        }
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        const float rs0 = rsqrtf(v0 * (1.0f / kE) + p.ln_eps), rs1 = rsqrtf(v1 * (1.0f / kE) + p.ln_eps);
#pragma unroll
        for (int j = 0; j < 24; ++j) {
            int col = 8 * j + 2 * t;
            float y00 = (acc[j][0] - mu0) * rs0, y01 = (acc[j][1] - mu0) * rs0;
            float y10 = (acc[j][2] - mu1) * rs1, y11 = (acc[j][3] - mu1) * rs1;
            if (ok0) {
                *reinterpret_cast<float2*>(p.Cf + r0 * p.ldcf + col) = make_float2(y00, y01);
                *reinterpret_cast<uint32_t*>(p.Cb + r0 * p.ldcb + col) = pack_bf16x2(y00, y01);
            }
            if (ok1) {
                *reinterpret_cast<float2*>(p.Cf + r1 * p.ldcf + col) = make_float2(y10, y11);
                *reinterpret_cast<uint32_t*>(p.Cb + r1 * p.ldcb + col) = pack_bf16x2(y10, y11);
            }
        }
        return;
    }

#pragma unroll
    for (int j = 0; j < 24; ++j) {
        int col = n0 + 8 * j + 2 * t;
        if (col >= p.N) continue;  // N is even everywhere we use this
        float b0 = 0.f, b1 = 0.f;
        if (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_SCALE_F32) {
            if (p.bias) { b0 = p.bias[col]; b1 = p.bias[col + 1]; }
        }
        float y00 = acc[j][0] + b0, y01 = acc[j][1] + b1, y10 = acc[j][2] + b0, y11 = acc[j][3] + b1;
        if (EPI == EPI_BIAS_GELU_BF16) {
            y00 = gelu_erf(y00); y01 = gelu_erf(y01); y10 = gelu_erf(y10); y11 = gelu_erf(y11);
        }
        if (EPI == EPI_BIAS_SCALE_F32) {
            if (ok0) *reinterpret_cast<float2*>(p.Cf + r0 * p.ldcf + col) = make_float2(y00 * p.scale, y01 * p.scale);
            if (ok1) *reinterpret_cast<float2*>(p.Cf + r1 * p.ldcf + col) = make_float2(y10 * p.scale, y11 * p.scale);
        } else {
            if (ok0) *reinterpret_cast<uint32_t*>(p.Cb + r0 * p.ldcb + col) = pack_bf16x2(y00, y01);
            if (ok1) *reinterpret_cast<uint32_t*>(p.Cb + r1 * p.ldcb + col) = pack_bf16x2(y10, y11);
        }
    }
}

template <int EPI>
static inline cudaError_t launch_gemm_mma(const GemmArgs& a, cudaStream_t st) {
    static bool configured_dev[64] = {};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_mma_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             GM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(a.M, GM_BM), (unsigned)ceil_div(a.N, GM_BN));
    gemm_mma_kernel<EPI><<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace pfn
