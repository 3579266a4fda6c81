// Attention between items, warp-level (mma.sync) implementation.
//
// For one token column t and one query head h: queries = rows' (t, h) vectors (dh = 32), keys/values =
// the context rows of the same column.  Context rows (prefill) use K/V head h of the freshly projected
// qkv buffer; test rows use the cached head-0 K/V for all six query heads
// (multiquery_item_attention_for_test_set).  Flash-attention-2 style: one warp owns 16 query rows,
// a CTA 128; K/V stream through shared memory in 64-key tiles (3-stage cp.async ring); online softmax
// in fp32 registers, P in bf16 for the second contraction.
#pragma once
#include "common.cuh"

namespace pfn {

struct AttnArgs {
    const bf16* Q;        // query (r, t, h) at Q + r*q_row + t*q_tok + h*32
    int64_t q_row, q_tok;
    const bf16* K;        // key j of (t, h) at K + t*k_tok + j*k_row + h*k_head ; value at + v_off
    int64_t k_tok, k_row;
    int k_head, v_off;
    bf16* O;              // out (r, t, h) at O + r*o_row + t*o_tok + h*32
    int64_t o_row, o_tok;
    int64_t R;            // query rows
    int64_t N;            // keys
};

constexpr int AT_BM = 128, AT_BN = 64, AT_STAGES = 3, AT_THREADS = 256;
constexpr int AT_TILE_BYTES = AT_BN * kDh * 2;            // 4 KB (K or V)
constexpr int AT_SMEM_BYTES = AT_STAGES * 2 * AT_TILE_BYTES;  // 24 KB

__device__ __forceinline__ uint32_t at_swz(int row, int chunk) {  // [64][32] bf16 tile, 64-byte rows
    return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(AT_THREADS, 2) attn_mma_kernel(AttnArgs p) {
    __shared__ __align__(128) uint8_t smem[AT_SMEM_BYTES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int h = blockIdx.y, t = blockIdx.z;
    const int64_t m0 = (int64_t)blockIdx.x * AT_BM;
    const uint32_t sbase = smem_u32(smem);

    const bf16* Kb = p.K + (int64_t)t * p.k_tok + (int64_t)h * p.k_head;
    const bf16* Vb = Kb + p.v_off;
    const int ntiles = (int)((p.N + AT_BN - 1) / AT_BN);

    auto load_tile = [&](int stage, int kt) {
        // 64 keys x (4 K chunks + 4 V chunks) of 16 B = 512 chunks, 2 per thread
        const uint32_t sk = sbase + stage * 2 * AT_TILE_BYTES;
        const uint32_t sv = sk + AT_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * AT_THREADS;
            const int row = idx >> 3, part = idx & 7;
            const int64_t j = (int64_t)kt * AT_BN + row;
            const bool ok = j < p.N;
            const int ch = part & 3;
            const bf16* src = (part < 4 ? Kb : Vb) + (ok ? j : 0) * p.k_row + ch * 8;
            cp_async16((part < 4 ? sk : sv) + at_swz(row, ch), src, ok ? 16 : 0);
        }
    };

    // Q fragments (A operand, 2 k-steps of 16) straight from global memory
    uint32_t qa[2][4];
    {
        const int64_t r0 = m0 + warp * 16 + g, r1 = r0 + 8;
        const bf16* q0 = p.Q + (r0 < p.R ? r0 : 0) * p.q_row + (int64_t)t * p.q_tok + h * kDh;
        const bf16* q1 = p.Q + (r1 < p.R ? r1 : 0) * p.q_row + (int64_t)t * p.q_tok + h * kDh;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            qa[kk][0] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * tq);
            qa[kk][1] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 2 * tq);
            qa[kk][2] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8 + 2 * tq);
            qa[kk][3] = *reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8 + 2 * tq);
        }
    }

    float o[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[j][c] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const float sc = 1.0f;  // queries arrive pre-scaled by kItemScaleLog2 (folded into the projection weights)

#pragma unroll
    for (int s = 0; s < AT_STAGES - 1; ++s) {
        if (s < ntiles) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < ntiles; ++kt) {
        cp_async_wait<AT_STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + AT_STAGES - 1;
            if (nk < ntiles) load_tile(nk % AT_STAGES, nk);
            cp_async_commit();
        }
        const uint32_t sk = sbase + (kt % AT_STAGES) * 2 * AT_TILE_BYTES;
        const uint32_t sv = sk + AT_TILE_BYTES;

        // S = Q K^T : 8 key n-tiles x 2 k-steps
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, sk + at_swz(8 * j + (lane & 7), lane >> 3));
            mma_bf16_16816(s[j], qa[0], b0, b1);
            mma_bf16_16816(s[j], qa[1], b2, b3);
        }
        // mask keys beyond N in the last tile
        if ((int64_t)(kt + 1) * AT_BN > p.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t key = (int64_t)kt * AT_BN + 8 * j + 2 * tq;
                if (key >= p.N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (key + 1 >= p.N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            }
        }
        // online softmax (rows g and g+8 of this warp)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
        }
        float corr[2], mnew[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mnew[i] = fmaxf(mrow[i], mx[i] * sc);
            corr[i] = fast_exp2(mrow[i] - mnew[i]);  // first tile: exp2(-inf) = 0
            mrow[i] = mnew[i];
            lrow[i] *= corr[i];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j][0] *= corr[0]; o[j][1] *= corr[0];
            o[j][2] *= corr[1]; o[j][3] *= corr[1];
        }
        uint32_t pa[4][4];  // P as A operand: 4 k-steps of 16 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = fast_exp2(fmaf(s[j][0], sc, -mnew[0]));
            const float p1 = fast_exp2(fmaf(s[j][1], sc, -mnew[0]));
            const float p2 = fast_exp2(fmaf(s[j][2], sc, -mnew[1]));
            const float p3 = fast_exp2(fmaf(s[j][3], sc, -mnew[1]));
            const uint32_t lo = pack_bf16x2(p0, p1), hi = pack_bf16x2(p2, p3);
            // row sums from the rounded values that enter the second contraction
            lrow[0] += bf16_lo(lo) + bf16_hi(lo);
            lrow[1] += bf16_lo(hi) + bf16_hi(hi);
            pa[j >> 1][(j & 1) * 2 + 0] = lo;
            pa[j >> 1][(j & 1) * 2 + 1] = hi;
        }
        // O += P V : 4 k-steps (16 keys) x 4 dh n-tiles
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {
                uint32_t b0, b1, b2, b3;
                const int row = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int ch = 2 * jp + (lane >> 4);
                ldmatrix_x4_trans(b0, b1, b2, b3, sv + at_swz(row, ch));
                mma_bf16_16816(o[2 * jp], pa[kk], b0, b1);
                mma_bf16_16816(o[2 * jp + 1], pa[kk], b2, b3);
            }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        lrow[i] += __shfl_xor_sync(0xffffffffu, lrow[i], 1);
        lrow[i] += __shfl_xor_sync(0xffffffffu, lrow[i], 2);
    }
    const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
    const int64_t r0 = m0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = h * kDh + 8 * j + 2 * tq;
        if (r0 < p.R)
            *reinterpret_cast<uint32_t*>(p.O + r0 * p.o_row + (int64_t)t * p.o_tok + col) =
                pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
        if (r1 < p.R)
            *reinterpret_cast<uint32_t*>(p.O + r1 * p.o_row + (int64_t)t * p.o_tok + col) =
                pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

static inline cudaError_t launch_attn_mma(const AttnArgs& a, int heads, int T, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(a.R, AT_BM), (unsigned)heads, (unsigned)T);
    attn_mma_kernel<<<grid, AT_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace pfn
