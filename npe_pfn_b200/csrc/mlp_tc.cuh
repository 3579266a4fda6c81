// Fused MLP sub-layer of the per-feature transformer on tcgen05 / TMEM / TMA:
//     x_out = LayerNorm(x + W2 · gelu(W1 · x_bf16))        (no biases)
// in ONE kernel, so the 768-wide hidden activation never leaves the SM (the two-kernel path writes and re-reads
// 3 KB per token through HBM / L2).  Per 128-token tile:
//   up-projection in 12 blocks of 64 hidden units:  acc_h[b] (64 TMEM columns, double-buffered) = X · W1[blk]^T
//   GELU epilogue (two warpgroups, alternating blocks): acc_h -> gelu -> bf16 -> shared memory, written directly in
//       the K-major 128-byte-swizzle layout the next MMA wants as its A operand (a [128 x 64] k-block);
//   down-projection: acc_out[o] (192 TMEM columns, double-buffered over tiles) += H[blk] · W2[:, blk]^T;
//   LayerNorm epilogue (third warpgroup): gemm_tc_epilogue<EPI_RESID_LN> on acc_out while the next tile is computed.
// Warp roles (704 threads): warps 0-15 four GELU warpgroups (warpgroup g drains accumulator g & 1, columns
// [32 (g >> 1), +32) of each block: the GELU epilogue is latency bound, so two warpgroups share every block),
// warps 16-19 LayerNorm warpgroup, warp 20 TMA producer (X tile once per tile; W1 block / W2 k-block stages of 24 KB
// through a 4-stage ring, in the order the MMA thread consumes them), warp 21 TMEM allocation + single-thread MMA issue.
// TMEM: 2 x 64 + 2 x 192 = 512 columns.
#pragma once
#include "gemm_tc.cuh"

namespace pfn {

constexpr int ML_GELU_WARPS = 16, ML_THREADS = (ML_GELU_WARPS + 6) * 32, ML_NB = 12, ML_BH = 64;
constexpr int ML_W_LN = ML_GELU_WARPS, ML_W_TMA = ML_GELU_WARPS + 4, ML_W_MMA = ML_GELU_WARPS + 5;
constexpr int ML_X_BYTES = 3 * GT_A_BYTES;             // 128 x 192 bf16 as three [128 x 64] k-blocks
constexpr int ML_H_BYTES = GT_A_BYTES, ML_H_BUFS = 3;  // [128 x 64] bf16 blocks of gelu(hidden)
constexpr int ML_W_BYTES = 24576, ML_W_STAGES = 4;     // W1 block = 3 x [64 x 64], W2 k-block = [192 x 64]
constexpr int ML_STAGING_BYTES = 4 * GT_STAGING_BYTES;
constexpr int ML_SMEM_BYTES = 1024 + ML_X_BYTES + ML_H_BUFS * ML_H_BYTES + ML_W_STAGES * ML_W_BYTES + ML_STAGING_BYTES + 512;
constexpr uint32_t ML_ACC_OUT = 128, ML_ACC_OUT_STRIDE = 192;

__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmCb,
              const __grid_constant__ CUtensorMap tmCf, const GemmTcArgs p) {
    extern __shared__ uint8_t ml_smem_raw[];
    const uint32_t raw = smem_u32(ml_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sX = base;
    const uint32_t sH = sX + ML_X_BYTES;
    const uint32_t sW = sH + ML_H_BUFS * ML_H_BYTES;
    const uint32_t staging = sW + ML_W_STAGES * ML_W_BYTES;
    const uint32_t bars = staging + ML_STAGING_BYTES;
    const uint32_t bar_x_full = bars, bar_x_empty = bars + 8;
    const uint32_t bar_w_full = bars + 16;                       // [4]
    const uint32_t bar_w_empty = bar_w_full + 8 * ML_W_STAGES;   // [4]
    const uint32_t bar_ah_full = bar_w_empty + 8 * ML_W_STAGES;  // [2]  acc_h ready      MMA -> GELU warpgroup
    const uint32_t bar_ah_empty = bar_ah_full + 16;              // [2]  acc_h drained    GELU -> MMA
    const uint32_t bar_h_full = bar_ah_empty + 16;               // [3]  H block written  GELU -> MMA
    const uint32_t bar_h_empty = bar_h_full + 8 * ML_H_BUFS;     // [3]  H block consumed MMA -> GELU
    const uint32_t bar_ao_full = bar_h_empty + 8 * ML_H_BUFS;    // [2]  acc_out ready    MMA -> LayerNorm warpgroup
    const uint32_t bar_ao_empty = bar_ao_full + 16;              // [2]
    const uint32_t bar_resid = bar_ao_empty + 16;                // [4 warps][2]
    const uint32_t tmem_slot = bar_resid + 16 * 4;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(ml_smem_raw + (tmem_slot - raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar_x_full, 1);
        mbar_init(bar_x_empty, 1);
        for (int s = 0; s < ML_W_STAGES; ++s) { mbar_init(bar_w_full + 8 * s, 1); mbar_init(bar_w_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_ah_full + 8 * b, 1);
            mbar_init(bar_ah_empty + 8 * b, ML_GELU_WARPS * 16);
            mbar_init(bar_ao_full + 8 * b, 1);
            mbar_init(bar_ao_empty + 8 * b, 128);
        }
        for (int h = 0; h < ML_H_BUFS; ++h) { mbar_init(bar_h_full + 8 * h, ML_GELU_WARPS * 16); mbar_init(bar_h_empty + 8 * h, 1); }
        for (int w = 0; w < 8; ++w) mbar_init(bar_resid + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ML_W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == ML_W_TMA) {
        // ================= TMA producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
            uint32_t wit = 0, lt = 0;
            for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
                mbar_wait(bar_x_empty, (lt & 1u) ^ 1u);  // the previous tile's up-projections have finished reading X
                mbar_expect_tx(bar_x_full, ML_X_BYTES);
                for (int kb = 0; kb < 3; ++kb) tma_load_2d(sX + kb * GT_A_BYTES, &tmX, bar_x_full, kb * GT_BK, (int)tile * GT_BM);
                for (int step = 0; step <= ML_NB; ++step) {
                    if (step < ML_NB) {  // W1 rows [64 step, 64 step + 64): three [64 x 64] k-blocks
                        const uint32_t s = wit % ML_W_STAGES, ph = (wit / ML_W_STAGES) & 1u;
                        mbar_wait(bar_w_empty + 8 * s, ph ^ 1u);
                        mbar_expect_tx(bar_w_full + 8 * s, ML_W_BYTES);
                        for (int kb = 0; kb < 3; ++kb)
                            tma_load_2d(sW + s * ML_W_BYTES + kb * 8192, &tmW1, bar_w_full + 8 * s, kb * GT_BK, step * ML_BH);
                        ++wit;
                    }
                    if (step >= 1) {  // W2 columns [64 (step-1), +64): one [192 x 64] k-block
                        const uint32_t s = wit % ML_W_STAGES, ph = (wit / ML_W_STAGES) & 1u;
                        mbar_wait(bar_w_empty + 8 * s, ph ^ 1u);
                        mbar_expect_tx(bar_w_full + 8 * s, ML_W_BYTES);
                        tma_load_2d(sW + s * ML_W_BYTES, &tmW2, bar_w_full + 8 * s, (step - 1) * ML_BH, 0);
                        ++wit;
                    }
                }
            }
        }
    } else if (warp == ML_W_MMA) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            constexpr uint32_t idesc_up = umma_idesc_bf16(GT_BM, ML_BH, 0, 0);
            constexpr uint32_t idesc_dn = umma_idesc_bf16(GT_BM, GT_BN, 0, 0);
            uint32_t wit = 0, lt = 0, ucount = 0, hcount = 0;
            for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
                const uint32_t o = lt & 1u;
                mbar_wait(bar_x_full, lt & 1u);
                tc_fence_after();
                for (int step = 0; step <= ML_NB; ++step) {
                    if (step < ML_NB) {
                        const uint32_t b = ucount & 1u;
                        mbar_wait(bar_ah_empty + 8 * b, ((ucount >> 1) & 1u) ^ 1u);
                        const uint32_t s = wit % ML_W_STAGES;
                        mbar_wait(bar_w_full + 8 * s, (wit / ML_W_STAGES) & 1u);
                        tc_fence_after();
#pragma unroll
                        for (int kb = 0; kb < 3; ++kb) {
                            const uint64_t dA = umma_desc_sw128(sX + kb * GT_A_BYTES);
                            const uint64_t dB = umma_desc_sw128(sW + s * ML_W_BYTES + kb * 8192);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_ss(tmem + b * ML_BH, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idesc_up, (kb > 0) || (kk > 0));
                        }
                        tc_commit(bar_w_empty + 8 * s);
                        tc_commit(bar_ah_full + 8 * b);
                        if (step == ML_NB - 1) tc_commit(bar_x_empty);
                        ++wit; ++ucount;
                    }
                    if (step >= 1) {
                        const int nb = step - 1;
                        const uint32_t hb = hcount % ML_H_BUFS;
                        mbar_wait(bar_h_full + 8 * hb, (hcount / ML_H_BUFS) & 1u);
                        if (nb == 0) mbar_wait(bar_ao_empty + 8 * o, ((lt >> 1) & 1u) ^ 1u);
                        const uint32_t s = wit % ML_W_STAGES;
                        mbar_wait(bar_w_full + 8 * s, (wit / ML_W_STAGES) & 1u);
                        tc_fence_after();
                        const uint64_t dA = umma_desc_sw128(sH + hb * ML_H_BYTES);
                        const uint64_t dB = umma_desc_sw128(sW + s * ML_W_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_ss(tmem + ML_ACC_OUT + o * ML_ACC_OUT_STRIDE, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idesc_dn,
                                    (nb > 0) || (kk > 0));
                        tc_commit(bar_w_empty + 8 * s);
                        tc_commit(bar_h_empty + 8 * hb);
                        if (nb == ML_NB - 1) tc_commit(bar_ao_full + 8 * o);
                        ++wit; ++hcount;
                    }
                }
            }
        }
    } else if (warp < ML_GELU_WARPS) {
        // ================= GELU warpgroups: thread = token row = TMEM lane, 32 of the block's 64 columns =================
        const uint32_t e = ((uint32_t)warp >> 2) & 1u, half = (uint32_t)warp >> 3;  // accumulator, column half
        const int row = (warp & 3) * 32 + lane;
        const uint32_t tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + e * ML_BH + half * 32;
        uint32_t lt = 0, gcount = 0;
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
            for (int nb = (int)e; nb < ML_NB; nb += 2, ++gcount) {
                const uint32_t G = lt * ML_NB + (uint32_t)nb, hb = G % ML_H_BUFS;
                mbar_wait(bar_ah_full + 8 * e, gcount & 1u);
                tc_fence_after();
                mbar_wait(bar_h_empty + 8 * hb, ((G / ML_H_BUFS) & 1u) ^ 1u);
                const uint32_t hrow = sH + hb * ML_H_BYTES;
                uint32_t v[32];
                tmem_ld32(tacc, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = gelu_fast(__uint_as_float(v[8 * q + i]));
                    sts128(hrow + sw128_off(row, (int)half * 4 + q), pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                           pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                }
                fence_async_smem();  // generic-proxy writes of H -> visible to the tensor core (async proxy)
                tc_fence_before();
                mbar_arrive(bar_h_full + 8 * hb);
                mbar_arrive(bar_ah_empty + 8 * e);
            }
        }
    } else {
        // ================= LayerNorm warpgroup =================
        const int wq = warp - ML_W_LN;
        const uint32_t lane_base = (uint32_t)wq * 32u;
        const uint32_t stage = staging + (uint32_t)wq * GT_STAGING_BYTES;
        const uint32_t rbar = bar_resid + 16 * (uint32_t)wq;
        const EpiMaps mp{&tmCb, &tmCf};
        uint32_t ruse[2] = {0u, 0u};
        uint32_t lt = 0;
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
            const uint32_t o = lt & 1u;
            const int row0 = (int)tile * GT_BM + (int)lane_base;
            if (lane == 0) {
                bulk_wait_read0();
                mbar_expect_tx(rbar, 4096);
                tma_load_2d(stage, &tmCf, rbar, 0, row0);
            }
            mbar_wait(bar_ao_full + 8 * o, (lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem + (lane_base << 16) + ML_ACC_OUT + o * ML_ACC_OUT_STRIDE;
            gemm_tc_epilogue<EPI_RESID_LN>(p, mp, tacc, row0, 0, stage, rbar, ruse, lane);
            tc_fence_before();
            mbar_arrive(bar_ao_empty + 8 * o);
        }
        if (lane == 0) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ML_W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// 2-D bf16 map, 128-byte swizzle, box = 64 elements x box_rows
static inline bool make_map2_box(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t row_bytes,
                                 uint32_t box_rows) {
    return make_map2_sw128(m, base, inner, rows, row_bytes, box_rows);
}

// X[M, 192] (bf16, row stride lda) -> Cf / Cb [M, 192] (fp32 residual stream in / out, bf16 copy), W1 [768, 192], W2 [192, 768]
static inline cudaError_t launch_mlp_tc(const bf16* X, int64_t lda, const bf16* W1, const bf16* W2, bf16* Cb, int64_t ldcb,
                                        float* Cf, int64_t ldcf, int64_t M, float ln_eps, int num_sms, cudaStream_t st) {
    static bool configured_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap mX, mW1, mW2, mCb, mCf;
    if (!make_map2_box(&mX, X, kE, (uint64_t)M, (uint64_t)lda * 2, GT_BM)) return cudaErrorInvalidValue;
    if (!make_map2_box(&mW1, W1, kE, kHid, (uint64_t)kE * 2, ML_BH)) return cudaErrorInvalidValue;
    if (!make_map2_box(&mW2, W2, kHid, kE, (uint64_t)kHid * 2, GT_BN)) return cudaErrorInvalidValue;
    if (!make_map2_out(&mCb, Cb, false, kE, (uint64_t)M, (uint64_t)ldcb * 2)) return cudaErrorInvalidValue;
    if (!make_map2_out(&mCf, Cf, true, kE, (uint64_t)M, (uint64_t)ldcf * 4)) return cudaErrorInvalidValue;
    GemmTcArgs p{};
    p.M = M; p.N = kE; p.K = kHid; p.Cb = Cb; p.ldcb = ldcb; p.Cf = Cf; p.ldcf = ldcf; p.bias = nullptr;
    p.scale = 1.0f; p.ln_eps = ln_eps; p.n_chunks = 1;
    p.tiles = ceil_div(M, GT_BM);
    const unsigned grid = (unsigned)std::min<int64_t>(p.tiles, num_sms);
    mlp_tc_kernel<<<grid, ML_THREADS, ML_SMEM_BYTES, st>>>(mX, mW1, mW2, mCb, mCf, p);
    return cudaGetLastError();
}

}  // namespace pfn
