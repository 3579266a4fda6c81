// Dense projections of the per-feature transformer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA):
//   C[M, N] = A[M, K] (bf16, row stride lda) x W[N, K]^T (bf16), fp32 accumulation in TMEM,
// with the layer's element-wise work fused into the epilogue (same four variants as gemm_mma.cuh):
//   EPI_BF16            -> bf16                              (QKV / Q projections)
//   EPI_BIAS_GELU_BF16  -> gelu(acc + bias) as bf16          (MLP up-projection, decoder hidden)
//   EPI_RESID_LN        -> LayerNorm(acc + residual), fp32 residual stream in/out + bf16 copy (N == 192)
//   EPI_BIAS_SCALE_F32  -> (acc + bias) * scale as fp32      (decoder logits / temperature)
//   EPI_FEATURE_ATTN    -> the feature-attention QKV projection FUSED with the attention between the T <= 16 tokens of a
//                          row: W is the head-pair-major permutation of [Wq; Wk; Wv] (chunk hp = q, k, v of heads 2hp and
//                          2hp+1), an M tile holds whole rows (floor(128 / T) rows), the epilogue rounds the accumulator
//                          to bf16 into shared memory and runs the row-wise attention there with warp-level mma.sync
//                          (the arithmetic of feature_attn_mma_kernel, bit for bit); only the attention output [tokens,
//                          192] goes back to HBM - the [tokens, 576] qkv round trip (2.3 KB per token and layer) is gone.
//
// Persistent, warp-specialised, one CTA per SM (320 threads):
//   warps 0-3 / 4-7  two epilogue warpgroups, alternating over the CTA's tiles (thread = output row = TMEM lane):
//                    tcgen05.ld the 128 x 192 fp32 accumulator, apply the epilogue, store straight to global;
//   warp 8           TMA producer: A (128 x 64) and W (192 x 64) k-blocks, 128-byte swizzle, 4-stage mbarrier ring;
//   warp 9           TMEM allocation (2 x 256 columns = two accumulators) + single-thread tcgen05.mma issue
//                    (M128 N192 K16, four per k-block), tcgen05.commit frees the stage / publishes the accumulator.
// While one warpgroup drains accumulator b, the tensor pipe already fills accumulator b^1.
#pragma once
#include <cuda.h>

#include "attn_tc.cuh"
#include "common.cuh"
#include "gemm_mma.cuh"

namespace pfn {

constexpr int GT_BM = 128, GT_BN = 192, GT_BK = 64, GT_STAGES = 4, GT_THREADS = 320;
constexpr int GT_A_BYTES = GT_BM * GT_BK * 2;   // 16 KB
constexpr int GT_B_BYTES = GT_BN * GT_BK * 2;   // 24 KB
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;
constexpr int GT_EPI_WARPS = 8;
constexpr int GT_STAGING_BYTES = 8192;          // per epilogue warp: two 4 KB sub-tile buffers
constexpr int GT_SMEM_BYTES = 1024 + GT_STAGES * GT_STAGE_BYTES + GT_EPI_WARPS * GT_STAGING_BYTES + 512;
// EPI_FEATURE_ATTN: 3 ring stages; per epilogue warpgroup a [128 tokens][192 bf16] q|k|v tile with 400-byte rows (16-byte
// aligned, 100 words = 4 mod 32 banks: the 8 rows x 4 columns of an mma fragment load hit 32 different banks)
constexpr int GT_FA_STAGES = 3, GT_FA_ROW_BYTES = 400, GT_FA_WG_BYTES = GT_BM * GT_FA_ROW_BYTES;
constexpr int GT_FA_SMEM_BYTES = 1024 + GT_FA_STAGES * GT_STAGE_BYTES + 2 * GT_FA_WG_BYTES + 512;
template <int EPI> struct GtCfg {
    static constexpr int stages = EPI == EPI_FEATURE_ATTN ? GT_FA_STAGES : GT_STAGES;
    static constexpr int staging_bytes = EPI == EPI_FEATURE_ATTN ? 2 * GT_FA_WG_BYTES : GT_EPI_WARPS * GT_STAGING_BYTES;
    static constexpr int smem_bytes = EPI == EPI_FEATURE_ATTN ? GT_FA_SMEM_BYTES : GT_SMEM_BYTES;
};
constexpr int GT_TMEM_COLS = 512, GT_ACC_STRIDE = 256;

struct GemmTcArgs {
    int64_t M;
    int N, K;
    bf16* Cb;
    int64_t ldcb;
    float* Cf;
    int64_t ldcf;
    const float* bias;
    float scale, ln_eps;
    int n_chunks;
    int64_t tiles;
    // EPI_FEATURE_ATTN: tokens per row, rows per M tile, total rows, output [rows * T, 192] bf16 (row stride 192)
    int T, rows_per_tile;
    int64_t R;
    int64_t m_stride;  // first A row of M tile i = i * m_stride (GT_BM for every other epilogue)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major tile with 128-byte rows, 128-byte swizzle: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

// GELU with erf by Abramowitz & Stegun, arranged for few instructions:
//   erf(|z|) = 1 - w,  w = (a1 t + a2 t^2 + ...) exp(-z^2),  t = 1 / (1 + p |z|),  z = x / sqrt(2)
//   gelu(x)  = 0.5 x (1 + sign(x) erf|z|) = max(x, 0) - |x| (w / 2)
// (the polynomial carries the factor 1/2, p and the exponent scale are expressed in x, so neither z nor a sign transfer
// is ever formed).  PFN_GELU_TERMS = 3: formula 7.1.25, |erf error| <= 2.5e-5, i.e. |gelu error| <= 1.3e-5 |x| -- far
// below the bf16 rounding (2^-9 relative) applied to every value this produces; one rcp, one ex2, 4 FFMA, 4 FMUL, 1 FMNMX.
// PFN_GELU_TERMS = 5: formula 7.1.26, |erf error| < 7e-7 (two more FFMA).
#ifndef PFN_GELU_TERMS
#define PFN_GELU_TERMS 3
#endif
__device__ __forceinline__ float gelu_fast(float x) {
    const float ax = fabsf(x);
    float t, q;
#if PFN_GELU_TERMS == 5
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f)));
    q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    q = fmaf(q, t, 0.5f * 1.421413741f);
    q = fmaf(q, t, 0.5f * -0.284496736f);
    q = fmaf(q, t, 0.5f * 0.254829592f);
#else
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.47047f * 0.70710678118654752440f, ax, 1.0f)));
    q = fmaf(0.5f * 0.7478556f, t, 0.5f * -0.0958798f);
    q = fmaf(q, t, 0.5f * 0.3480242f);
#endif
    const float e = fast_exp2((x * x) * (-0.5f * 1.4426950408889634f));  // exp(-x^2 / 2)
    const float w = q * (t * e);
    return fmaf(-ax, w, fmaxf(x, 0.0f));
}

// ---- epilogue: every global access goes through TMA on 32 x 32 sub-tiles staged in swizzled shared memory ----------
// (direct per-thread row accesses touch 32 different lines per warp instruction and saturate the LSU: r1 profile)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// sub-tile layouts: fp32 [32 rows][128 B] with the 128-byte swizzle, bf16 [32 rows][64 B] with the 64-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int q) { return (uint32_t)(row * 128 + ((q ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint32_t sw64_off(int row, int q) { return (uint32_t)(row * 64 + ((q ^ ((row >> 1) & 3)) << 4)); }

struct EpiMaps {
    const CUtensorMap* cb;  // bf16 output, box 32 x 32, SWIZZLE_64B
    const CUtensorMap* cf;  // fp32 output / residual, box 32 x 32, SWIZZLE_128B
};

// One warp, one 128 x 192 accumulator tile: rows [row0, row0 + 32) of the tile (row0 already includes the warp's
// lane quarter).  `stage` = this warp's 8 KB staging area, `rbar` = its two residual mbarriers, `ruse` = use counters.
template <int EPI>
__device__ __forceinline__ void gemm_tc_epilogue(const GemmTcArgs& p, const EpiMaps& mp, uint32_t tacc, int row0, int n0,
                                                 uint32_t stage, uint32_t rbar, uint32_t (&ruse)[2], int lane) {
    constexpr int NCH = GT_BN / 32;
    if (EPI == EPI_RESID_LN) {
        // pass 1: x = acc + residual (fp32 residual slices arrive by TMA, double-buffered), row sums, x parked in TMEM
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            __syncwarp();
            if (lane == 0 && c + 1 < NCH) {
                const uint32_t b = (uint32_t)(c + 1) & 1u;
                mbar_expect_tx(rbar + 8 * b, 4096);
                tma_load_2d(stage + b * 4096, mp.cf, rbar + 8 * b, (c + 1) * 32, row0);
            }
            uint32_t v[32];
            tmem_ld32(tacc + c * 32, v);
            tmem_wait_ld();
            const uint32_t b = (uint32_t)c & 1u;
            mbar_wait(rbar + 8 * b, ruse[b] & 1u);
            ruse[b]++;
            const uint32_t rb = stage + b * 4096;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 r4 = lds128(rb + sw128_off(lane, q));
                const float x0 = __uint_as_float(v[4 * q + 0]) + r4.x, x1 = __uint_as_float(v[4 * q + 1]) + r4.y;
                const float x2 = __uint_as_float(v[4 * q + 2]) + r4.z, x3 = __uint_as_float(v[4 * q + 3]) + r4.w;
                s1 += (x0 + x1) + (x2 + x3);
                s2 = fmaf(x0, x0, s2); s2 = fmaf(x1, x1, s2); s2 = fmaf(x2, x2, s2); s2 = fmaf(x3, x3, s2);
                v[4 * q + 0] = __float_as_uint(x0); v[4 * q + 1] = __float_as_uint(x1);
                v[4 * q + 2] = __float_as_uint(x2); v[4 * q + 3] = __float_as_uint(x3);
            }
            tmem_st32(tacc + c * 32, v);
        }
        tmem_wait_st();
        const float mu = s1 * (1.0f / kE);
        const float var = fmaxf(s2 * (1.0f / kE) - mu * mu, 0.f);
        const float rs = rsqrtf(var + p.ln_eps);
        // pass 2: normalise; fp32 sub-tile -> buffer 0, bf16 sub-tile -> buffer 1, both stored by TMA
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(tacc + c * 32, v);
            tmem_wait_ld();
            if (lane == 0) bulk_wait_read0();  // the previous sub-tile stores have finished reading the buffers
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = (__uint_as_float(v[8 * q + i]) - mu) * rs;
                sts128(stage + sw128_off(lane, 2 * q), __float_as_uint(y[0]), __float_as_uint(y[1]), __float_as_uint(y[2]),
                       __float_as_uint(y[3]));
                sts128(stage + sw128_off(lane, 2 * q + 1), __float_as_uint(y[4]), __float_as_uint(y[5]),
                       __float_as_uint(y[6]), __float_as_uint(y[7]));
                sts128(stage + 4096 + sw64_off(lane, q), pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                       pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(mp.cf, stage, c * 32, row0);
                tma_store_2d(mp.cb, stage + 4096, c * 32, row0);
                bulk_commit();
            }
        }
        return;
    }
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tacc + c * 32, v);
        tmem_wait_ld();
        const uint32_t buf = stage + (uint32_t)(c & 1) * 4096;  // alternate buffers: only wait for the store before last
        // two staging buffers alternate: before rewriting one, all stores but the most recent have read their source
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        if (EPI == EPI_BIAS_SCALE_F32) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int col = col0 + 4 * q;
                float b[4] = {0.f, 0.f, 0.f, 0.f};
                if (col + 3 < p.N) {
                    const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col);
                    b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
                } else {
                    for (int i = 0; i < 4; ++i)
                        if (col + i < p.N) b[i] = p.bias[col + i];
                }
                sts128(buf + sw128_off(lane, q), __float_as_uint((__uint_as_float(v[4 * q + 0]) + b[0]) * p.scale),
                       __float_as_uint((__uint_as_float(v[4 * q + 1]) + b[1]) * p.scale),
                       __float_as_uint((__uint_as_float(v[4 * q + 2]) + b[2]) * p.scale),
                       __float_as_uint((__uint_as_float(v[4 * q + 3]) + b[3]) * p.scale));
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(v[8 * q + i]);
                if (EPI == EPI_BIAS_GELU_BF16) {
                    if (p.bias) {
                        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col0 + 8 * q);
                        const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col0 + 8 * q + 4);
                        y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w;
                        y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = gelu_fast(y[i]);
                }
                sts128(buf + sw64_off(lane, q), pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                       pack_bf16x2(y[6], y[7]));
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(EPI == EPI_BIAS_SCALE_F32 ? mp.cf : mp.cb, buf, col0, row0);
            bulk_commit();
        }
    }
}

// EPI_FEATURE_ATTN, one epilogue warpgroup (4 warps, 128 threads = 128 tokens of the M tile), one head pair `hp`.
// `buf` = the warpgroup's [128][GT_FA_ROW_BYTES] tile, `bar_id` = its named barrier, `acc_empty` = the accumulator's
// release barrier (arrived as soon as the accumulator has been copied out, so the next tile's MMAs overlap the attention).
__device__ __forceinline__ void gemm_tc_feature_attn(const GemmTcArgs& p, uint32_t tacc, uint8_t* buf, int bar_id, uint32_t acc_empty,
                                                     int64_t m_blk, int hp, int wq, int lane) {
    // 1. accumulator row (q | k | v of two heads, 192 fp32) -> bf16 -> this token's row of the tile
    uint8_t* myrow = buf + (size_t)(wq * 32 + lane) * GT_FA_ROW_BYTES;
#pragma unroll 1
    for (int c = 0; c < GT_BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tacc + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
            w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
            w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
            w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
            *reinterpret_cast<uint4*>(myrow + c * 64 + q * 16) = w;
        }
    }
    tc_fence_before();
    mbar_arrive(acc_empty);  // the tensor pipe may refill this accumulator now
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    // 2. attention between the T tokens of each row, one warp per (row, head): the whole T x T problem is one 16 x 16 tile
    //    (S = Q K^T: 2 n-tiles x 2 k-steps of m16n8k16; O = P V: 4 n-tiles x 1 k-step), as feature_attn_mma_kernel does
    const int T = p.T, G = p.rows_per_tile;
    const int g = lane >> 2, tq = lane & 3;
    constexpr int ts = GT_FA_ROW_BYTES / 2;  // token stride in bf16 elements
    auto ld32 = [](const bf16* ptr) { return *reinterpret_cast<const uint32_t*>(ptr); };
    for (int pair = wq; pair < 2 * G; pair += 4) {
        const int rl = pair >> 1, hh = pair & 1;
        const int64_t r = m_blk * G + rl;
        if (r >= p.R) break;  // warp-uniform (pairs grow with the row)
        const bf16* base = reinterpret_cast<const bf16*>(buf) + (size_t)(rl * T) * ts + hh * kDh;
        const int q0 = min(g, T - 1), q1 = min(g + 8, T - 1);
        float s[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            uint32_t a[4];
            a[0] = ld32(base + q0 * ts + 16 * kk + 2 * tq);
            a[1] = ld32(base + q1 * ts + 16 * kk + 2 * tq);
            a[2] = ld32(base + q0 * ts + 16 * kk + 8 + 2 * tq);
            a[3] = ld32(base + q1 * ts + 16 * kk + 8 + 2 * tq);
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int key = min(8 * n + g, T - 1);
                const uint32_t b0 = ld32(base + key * ts + 64 + 16 * kk + 2 * tq);
                const uint32_t b1 = ld32(base + key * ts + 64 + 16 * kk + 8 + 2 * tq);
                mma_bf16_16816(s[n], a, b0, b1);
            }
        }
        const float sc = 0.17677669529663687f * 1.4426950408889634f;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const int key = 8 * n + 2 * tq;
            if (key >= T) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
            if (key + 1 >= T) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
        float l[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
            mx[i] *= sc;
        }
        uint32_t pa[4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            const float p0 = exp2f(fmaf(s[n][0], sc, -mx[0])), p1 = exp2f(fmaf(s[n][1], sc, -mx[0]));
            const float p2 = exp2f(fmaf(s[n][2], sc, -mx[1])), p3 = exp2f(fmaf(s[n][3], sc, -mx[1]));
            l[0] += p0 + p1;
            l[1] += p2 + p3;
            pa[2 * n + 0] = pack_bf16x2(p0, p1);
            pa[2 * n + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
            l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
        }
        const unsigned short* vb = reinterpret_cast<const unsigned short*>(base + 128);
        const int k0 = min(2 * tq, T - 1), k1 = min(2 * tq + 1, T - 1), k2 = min(2 * tq + 8, T - 1), k3 = min(2 * tq + 9, T - 1);
        const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
        bf16* out = p.Cb + (r * T) * kE + (2 * hp + hh) * kDh;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int dh = 8 * j + g;
            const uint32_t b0 = (uint32_t)vb[k0 * ts + dh] | ((uint32_t)vb[k1 * ts + dh] << 16);
            const uint32_t b1 = (uint32_t)vb[k2 * ts + dh] | ((uint32_t)vb[k3 * ts + dh] << 16);
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16_16816(o, pa, b0, b1);
            const int col = 8 * j + 2 * tq;
            if (g < T) *reinterpret_cast<uint32_t*>(out + (int64_t)g * kE + col) = pack_bf16x2(o[0] * inv0, o[1] * inv0);
            if (g + 8 < T) *reinterpret_cast<uint32_t*>(out + (int64_t)(g + 8) * kE + col) = pack_bf16x2(o[2] * inv1, o[3] * inv1);
        }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // everyone is done reading before the tile is overwritten
}

template <int EPI>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmCb, const __grid_constant__ CUtensorMap tmCf, const GemmTcArgs p) {
    extern __shared__ uint8_t gt_smem_raw[];
    const uint32_t raw = smem_u32(gt_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    constexpr int NST = GtCfg<EPI>::stages;
    const uint32_t staging = base + NST * GT_STAGE_BYTES;                       // 8 epilogue warps x 8 KB (FEATURE_ATTN: 2 x 50 KB)
    const uint32_t bars = staging + GtCfg<EPI>::staging_bytes;
    const uint32_t bar_full = bars;                      // [NST]
    const uint32_t bar_empty = bars + 8 * NST;           // [NST]
    const uint32_t bar_acc_full = bars + 16 * NST;         // [2]
    const uint32_t bar_acc_empty = bar_acc_full + 16;      // [2]
    const uint32_t bar_resid = bar_acc_full + 32;          // [GT_EPI_WARPS][2] residual sub-tiles landed
    const uint32_t tmem_slot = bar_resid + 16 * GT_EPI_WARPS;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gt_smem_raw + (tmem_slot - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.K / GT_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + 8 * b, 1);
            mbar_init(bar_acc_empty + 8 * b, 128);
        }
        for (int w = 0; w < 2 * GT_EPI_WARPS; ++w) mbar_init(bar_resid + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)GT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 8) {
        // ================= TMA producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int m_blk = (int)(tile / p.n_chunks), n_blk = (int)(tile % p.n_chunks);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                    mbar_expect_tx(bar_full + 8 * s, GT_STAGE_BYTES);
                    const uint32_t dA = base + s * GT_STAGE_BYTES, dB = dA + GT_A_BYTES;
                    tma_load_2d(dA, &tmA, bar_full + 8 * s, kb * GT_BK, (int)(m_blk * p.m_stride));
                    tma_load_2d(dB, &tmW, bar_full + 8 * s, kb * GT_BK, n_blk * GT_BN);
                }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(GT_BM, GT_BN, 0, 0);
            uint32_t it = 0, lt = 0;
            for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
                const uint32_t b = lt & 1u;
                mbar_wait(bar_acc_empty + 8 * b, ((lt >> 1) & 1u) ^ 1u);  // epilogue has drained accumulator b
                tc_fence_after();
                const uint32_t d = tmem + b * GT_ACC_STRIDE;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint64_t dA = umma_desc_sw128(base + s * GT_STAGE_BYTES);
                    const uint64_t dB = umma_desc_sw128(base + s * GT_STAGE_BYTES + GT_A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GT_BK / 16; ++kk)  // 32 bytes per K16 step inside the 128-byte swizzle atom
                        umma_ss(d, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idesc, (kb > 0) || (kk > 0));
                    tc_commit(bar_empty + 8 * s);
                }
                tc_commit(bar_acc_full + 8 * b);
            }
        }
    } else {
        // ================= epilogue warpgroups =================
        const uint32_t e = (uint32_t)warp >> 2;  // warpgroup 0 / 1 <-> accumulator 0 / 1
        const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
        const uint32_t stage = staging + (uint32_t)warp * GT_STAGING_BYTES;
        const uint32_t rbar = bar_resid + 16 * (uint32_t)warp;
        const EpiMaps mp{&tmCb, &tmCf};
        uint32_t ruse[2] = {0u, 0u};
        uint32_t lt = 0;
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
            if ((lt & 1u) != e) continue;
            const int m_blk = (int)(tile / p.n_chunks), n_blk = (int)(tile % p.n_chunks);
            const int row0 = m_blk * GT_BM + (int)lane_base;
            if (EPI == EPI_RESID_LN && lane == 0) {
                bulk_wait_read0();  // the previous tile's stores no longer read the staging buffers
                mbar_expect_tx(rbar, 4096);
                tma_load_2d(stage, &tmCf, rbar, 0, row0);  // first residual slice: overlaps the accumulator wait
            }
            mbar_wait(bar_acc_full + 8 * e, (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem + (lane_base << 16) + e * GT_ACC_STRIDE;
            if constexpr (EPI == EPI_FEATURE_ATTN) {
                gemm_tc_feature_attn(p, tacc, gt_smem_raw + (staging - raw) + (size_t)e * GT_FA_WG_BYTES, 1 + (int)e,
                                     bar_acc_empty + 8 * e, m_blk, n_blk, warp & 3, lane);
            } else {
                gemm_tc_epilogue<EPI>(p, mp, tacc, row0, n_blk * GT_BN, stage, rbar, ruse, lane);
                tc_fence_before();
                mbar_arrive(bar_acc_empty + 8 * e);
            }
        }
        if (lane == 0) bulk_wait0();  // all TMA stores of this warp have completed before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)GT_TMEM_COLS)
                     : "memory");
    }
}

// 2-D bf16 map (inner = K elements), 128-byte swizzle, box = 64 x rows
static inline bool make_map2_sw128(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_bytes,
                                   uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)GT_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D output / residual map over a row-major matrix, box = 32 columns x 32 rows
static inline bool make_map2_out(CUtensorMap* m, const void* base, bool fp32, uint64_t cols, uint64_t rows,
                                 uint64_t row_bytes) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, fp32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int EPI>
static inline cudaError_t launch_gemm_tc(const GemmArgs& a, int num_sms, cudaStream_t st) {
    static bool configured_dev[64] = {};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GtCfg<EPI>::smem_bytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    static_assert(EPI != EPI_FEATURE_ATTN, "use launch_qkv_feature_attn_tc");
    if (a.K % GT_BK != 0 || (EPI == EPI_RESID_LN && a.N != GT_BN)) return cudaErrorInvalidValue;
    CUtensorMap mA, mW, mCb, mCf;
    if (!make_map2_sw128(&mA, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda * 2, GT_BM)) return cudaErrorInvalidValue;
    if (!make_map2_sw128(&mW, a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * 2, GT_BN)) return cudaErrorInvalidValue;
    const bool need_b = EPI != EPI_BIAS_SCALE_F32, need_f = EPI == EPI_BIAS_SCALE_F32 || EPI == EPI_RESID_LN;
    if (need_b && !make_map2_out(&mCb, a.Cb, false, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldcb * 2))
        return cudaErrorInvalidValue;
    if (need_f && !make_map2_out(&mCf, a.Cf, true, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldcf * 4))
        return cudaErrorInvalidValue;
    if (!need_b) mCb = mCf;
    if (!need_f) mCf = mCb;
    GemmTcArgs p{};
    p.M = a.M; p.N = a.N; p.K = a.K; p.Cb = a.Cb; p.ldcb = a.ldcb; p.Cf = a.Cf; p.ldcf = a.ldcf; p.bias = a.bias;
    p.scale = a.scale; p.ln_eps = a.ln_eps;
    p.n_chunks = (a.N + GT_BN - 1) / GT_BN;
    p.tiles = ceil_div(a.M, GT_BM) * p.n_chunks;
    p.m_stride = GT_BM;
    const unsigned grid = (unsigned)std::min<int64_t>(p.tiles, num_sms);
    gemm_tc_kernel<EPI><<<grid, GT_THREADS, GT_SMEM_BYTES, st>>>(mA, mW, mCb, mCf, p);
    return cudaGetLastError();
}

// Fused feature-attention half step: X[rows * T, 192] (bf16, contiguous tokens) x Wperm[576, 192]^T -> attention between the
// T tokens of every row -> out[rows * T, 192] bf16.  Wperm = head-pair-major permutation of the layer's [Wq; Wk; Wv].
static inline cudaError_t launch_qkv_feature_attn_tc(const bf16* X, const bf16* Wperm, int64_t rows, int T, bf16* out, int num_sms,
                                                     cudaStream_t st) {
    if (T < 1 || T > 16) return cudaErrorInvalidValue;
    static bool configured_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI_FEATURE_ATTN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             GT_FA_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap mA, mW;
    if (!make_map2_sw128(&mA, X, (uint64_t)kE, (uint64_t)(rows * T), (uint64_t)kE * 2, GT_BM)) return cudaErrorInvalidValue;
    if (!make_map2_sw128(&mW, Wperm, (uint64_t)kE, (uint64_t)(3 * kE), (uint64_t)kE * 2, GT_BN)) return cudaErrorInvalidValue;
    GemmTcArgs p{};
    p.M = rows * T; p.N = 3 * kE; p.K = kE; p.Cb = out; p.ldcb = kE;
    p.T = T; p.rows_per_tile = GT_BM / T; p.R = rows;
    p.m_stride = (int64_t)p.rows_per_tile * T;
    p.n_chunks = 3;
    p.tiles = ceil_div(rows, p.rows_per_tile) * p.n_chunks;
    const unsigned grid = (unsigned)std::min<int64_t>(p.tiles, num_sms);
    gemm_tc_kernel<EPI_FEATURE_ATTN><<<grid, GT_THREADS, GT_FA_SMEM_BYTES, st>>>(mA, mW, mA, mA, p);
    return cudaGetLastError();
}

}  // namespace pfn
