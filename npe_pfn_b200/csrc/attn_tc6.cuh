// Item attention of TEST rows against the cached K/V, version 6 (round 2): the row sum leaves the softmax threads.
//
// What limits attn_tc v5 (attn_tc.cuh) is issue slots of the softmax warps, not a pipe (profiles/r1_ncu_attn_tc_v5_metrics.txt);
// per pair of scores a softmax thread issues {2 x MUFU.EX2, FADD2 (row sum), F2FP (pack)} or the FMA-pipe form plus
// the same FADD2 + F2FP.  The register-only loop drops from 7.4 to 6.75 clk per warp-element without the row sum and to
// 6.5 with four softmax warps per scheduler (profiles/r2_softmax_loop_rowsum_f16x2_microbench.txt).  So here
//   * the row sum l = sum_k P is a tensor-core product: besides O += P V the MMA thread issues L += P 1 (M128 N16 K16 per
//     16 keys, B = a constant all-ones tile), into 16 more TMEM columns; l is column 0 of L.  It sums the bf16-ROUNDED P -
//     exactly the weights O was built with - where v5 summed the unrounded values;
//   * one CTA owns TWO query tiles (the same 128 rows, heads 2p and 2p + 1) that share ONE K/V stream (test rows use
//     head 0's K/V for all six query heads), with S single-buffered per tile: TMEM per CTA = 2 x (64 S/P + 32 O + 16 L)
//     = 224 columns -> one 256-column allocation, two CTAs per SM = 16 softmax warps per SM (four per scheduler)
//     against v5's twelve, and half the K/V shared-memory traffic per query;
//   * because S is single-buffered, Q K^T of tile j + 1 is issued after the softmax threads have finished tile j, so a
//     reference maximum published during tile j is baked into tile j + 1 already: v5's two-tile lag, its per-buffer
//     bookkeeping and its "stale tile" general-path trips disappear.  While one tile waits for its next S, the three
//     other softmax warps of the scheduler keep the MUFU / FMA pipes busy.
//   * ONE thread issues the tensor-core work of both units in fixed order (unit 0, unit 1 per tile).
//
// MEASURED (B200, 37 888 draws, N = 10 000, profiles/r2_attn_tc6_experiments.txt): parity with the oracle is the same as
// v5's, but the kernel is SLOWER than v5 - 488 vs 512 TFLOP/s - so v5 stays the default and this file is an opt-in
// (`attn_impl = 2`).  The register loop gain is real, the single-buffered S is what costs: a unit's four warps idle for a
// full tensor-core round trip (P V + P 1 + Q K^T) every tile.  Variants tried: the MMA thread serving whichever unit's P
// arrives first with hinted barrier probes (425), un-hinted polling (460), each unit's row-0 softmax thread issuing its
// own products (340), exponential split k = 4..6 (within 1 %).  Double-buffering S with the row sum on the tensor core
// needs 176 TMEM columns per query tile, i.e. only 2.9 tiles per SM - the fix would be 48-key tiles or 3 tiles/SM with
// an asymmetric S layout; neither was tried.
// Overflow detection without a register row sum: bit 14 of a packed bf16 is set exactly when the value is >= 2, so the
// OR of all packed words (one 3-input LOP3 per two pairs) tells whether some P >= 2 (a NaN has the bit set too); the
// sign bits catch a wrapped exponent of the FMA-pipe form.  The rest is v5: scale folded into the Q projection,
// reference maximum folded into the Q K^T contraction as a third K = 16 block, integer-valued references, 5 of 16
// exponential pairs on the FMA pipes.
#pragma once
#include "attn_tc.cuh"

namespace pfn {

constexpr int T6_BM = 128, T6_BN = 64, T6_THREADS = 320, T6_UNITS = 2;
#ifndef PFN_ATTN6_STAGES
#define PFN_ATTN6_STAGES 8
#endif
constexpr int T6_STAGES = PFN_ATTN6_STAGES;
constexpr int T6_Q_BYTES = T6_BM * kDh * 2;            // 8 KB per query tile
constexpr int T6_TILE_BYTES = T6_BN * kDh * 2;         // 4 KB (K or V tile)
constexpr int T6_STAGE_BYTES = 2 * T6_TILE_BYTES;
constexpr int T6_QX_BYTES = T6_BM * 32, T6_KX_BYTES = T6_BN * 32, T6_ONES_BYTES = 512;
constexpr int T6_SMEM_BYTES = 1024 + T6_UNITS * T6_Q_BYTES + T6_STAGES * T6_STAGE_BYTES + T6_UNITS * T6_QX_BYTES + T6_KX_BYTES +
                              T6_ONES_BYTES + 256;
constexpr int T6_TMEM_COLS = 256;  // per unit: S/P 64 | O 32 | L 16 (+ 16 unused); unit u at column 128 u
constexpr int T6_COL_S = 0, T6_COL_O = 64, T6_COL_L = 96, T6_UNIT_COLS = 128;

// no-swizzle K-major descriptor (8-row x 16-byte core matrices, 128 B apart along K, 256 B apart along N): used for the
// constant all-ones B tile of the L += P 1 product, where every address inside the 512-byte region reads 1.0
__device__ __forceinline__ uint64_t umma_desc_ones(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;  // layout type 0: SWIZZLE_NONE
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}

// P = 2^x for 32 scores that arrive as x = s - m (16 packed bf16 pairs), no row sum; `ovf` ORs every packed word
template <int POLY16, int DEG>
__device__ __forceinline__ void softmax_exp32_v6(const uint32_t (&s)[32], uint32_t* pk, uint32_t& ovf) {
    const uint64_t CM = pk2(kExpMagic, kExpMagic), NEG1 = pk2(-1.0f, -1.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool poly = ((i + 1) * POLY16) / 16 != (i * POLY16) / 16;
        float p0, p1;
        if (!poly) {
            p0 = fast_exp2(__uint_as_float(s[2 * i]));
            p1 = fast_exp2(__uint_as_float(s[2 * i + 1]));
        } else {
            const uint64_t X2 = pk2(fmaxf(__uint_as_float(s[2 * i]), -125.0f), fmaxf(__uint_as_float(s[2 * i + 1]), -125.0f));
            const uint64_t t2 = fadd2(X2, CM);                   // magic + n, n = round(x)
            const uint64_t f2 = fadd2(X2, ffma2(t2, NEG1, CM));  // x - n
            uint64_t q2;
            if (DEG == 2) {
                q2 = ffma2(pk2(0.23842893540859222f, 0.23842893540859222f), f2, pk2(0.7034479975700378f, 0.7034479975700378f));
                q2 = ffma2(q2, f2, pk2(1.0004431009292603f, 1.0004431009292603f));
            } else {
                q2 = ffma2(pk2(0.05517163127660751f, 0.05517163127660751f), f2, pk2(0.2426111251115799f, 0.2426111251115799f));
                q2 = ffma2(q2, f2, pk2(0.6932609677314758f, 0.6932609677314758f));
                q2 = ffma2(q2, f2, pk2(0.9999280571937561f, 0.9999280571937561f));
            }
            float q0, q1, t0, t1;
            upk2(q2, q0, q1);
            upk2(t2, t0, t1);
            p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
            p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
        }
        pk[i] = pack_bf16x2(p0, p1);
#ifndef PFN_T6_NOOVF
        ovf |= pk[i];
#endif
    }
}

struct Tc6Args {
    bf16* O;
    int64_t o_row, o_tok;
    int64_t R, N;
    int v_dx;                 // x coordinate of V inside a cached key row (K at 0)
    int head_pairs, n_qtiles; // item = (t * n_qtiles + qt) * head_pairs + p; the CTA's units are heads 2p and 2p + 1
    uint32_t probe_ns;        // suspend-time hint of the MMA thread's barrier probes
    unsigned long long* dbg;  // optional counters: [0] tiles redone after the overflow check, [1] reference changes after
                              // a row's first tile, [2] tiles on the general path
};

template <int POLY16, int DEG>
__global__ void __launch_bounds__(T6_THREADS, 2)
attn_tc6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Tc6Args p) {
    extern __shared__ uint8_t t6_smem_raw[];
    const uint32_t raw = smem_u32(t6_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sQ = base;                                           // [2] query tiles
    const uint32_t sKV = sQ + T6_UNITS * T6_Q_BYTES;                    // K/V ring
    const uint32_t sQx = sKV + T6_STAGES * T6_STAGE_BYTES;              // [2] (-m_hi, -m_lo) columns of the query side
    const uint32_t sKx = sQx + T6_UNITS * T6_QX_BYTES;                  // constant (1, 1, 0, ...) of the key side
    const uint32_t sOnes = sKx + T6_KX_BYTES;                           // 512 B of bf16 1.0
    const uint32_t bars = sOnes + T6_ONES_BYTES;
    const uint32_t bar_kv_full = bars;                        // [T6_STAGES]  TMA -> MMA
    const uint32_t bar_kv_empty = bars + 8 * T6_STAGES;       // [T6_STAGES]  MMA -> TMA
    const uint32_t bar_q = bars + 16 * T6_STAGES;             // both query tiles landed
    const uint32_t bar_s = bar_q + 8;                         // [2] S of unit u holds Q K_j^T         MMA -> softmax
    const uint32_t bar_p = bar_q + 24;                        // [2] P_j of unit u written             softmax -> MMA
    const uint32_t bar_pv = bar_q + 40;                       // [2] O/L += P_j V_j of unit u finished MMA -> softmax
    const uint32_t bar_o = bar_q + 56;                        // [2] all products of unit u finished
    const uint32_t tmem_slot = bar_q + 72;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(t6_smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (int)((p.N + T6_BN - 1) / T6_BN);
    const int64_t item = blockIdx.x;
    const int per_col = p.n_qtiles * p.head_pairs;
    const int t = (int)(item / per_col), rem = (int)(item % per_col);
    const int qt = rem / p.head_pairs, hp = rem % p.head_pairs;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T6_STAGES; ++s) {
            mbar_init(bar_kv_full + 8 * s, 1);
            mbar_init(bar_kv_empty + 8 * s, 1);
        }
        mbar_init(bar_q, 1);
        for (int u = 0; u < T6_UNITS; ++u) {
            mbar_init(bar_s + 8 * u, 1);
            mbar_init(bar_p + 8 * u, 128);
            mbar_init(bar_pv + 8 * u, 1);
            mbar_init(bar_o + 8 * u, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        uint32_t* ext = reinterpret_cast<uint32_t*>(t6_smem_raw + (sQx - raw));
        for (int i = threadIdx.x; i < (T6_UNITS * T6_QX_BYTES + T6_KX_BYTES) / 4; i += T6_THREADS) ext[i] = 0u;
        uint32_t* ones = reinterpret_cast<uint32_t*>(t6_smem_raw + (sOnes - raw));
        for (int i = threadIdx.x; i < T6_ONES_BYTES / 4; i += T6_THREADS) ones[i] = 0x3F803F80u;
        __syncthreads();
        if (threadIdx.x < T6_BN)  // key side: ones in the two columns that carry -m_hi and -m_lo
            *reinterpret_cast<uint32_t*>(t6_smem_raw + (sKx - raw) + sw32_chunk0(threadIdx.x)) = 0x3F803F80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)T6_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot_ptr[0];

    if (warp == 8) {
        // ================= TMA producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
            mbar_expect_tx(bar_q, T6_UNITS * T6_Q_BYTES);
            for (int u = 0; u < T6_UNITS; ++u)
                tma_load_3d(sQ + u * T6_Q_BYTES, &tmQ, bar_q, (2 * hp + u) * kDh, t, qt * T6_BM);
            for (int j = 0; j < ntiles; ++j) {
                const uint32_t s = (uint32_t)j % T6_STAGES, ph = ((uint32_t)j / T6_STAGES) & 1u;
                mbar_wait_hint(bar_kv_empty + 8 * s, ph ^ 1u, 1000);
                mbar_expect_tx(bar_kv_full + 8 * s, T6_STAGE_BYTES);
                const uint32_t dstK = sKV + s * T6_STAGE_BYTES, dstV = dstK + T6_TILE_BYTES;
                tma_load_3d(dstK, &tmKV, bar_kv_full + 8 * s, 0, j * T6_BN, t);
                tma_load_3d(dstV, &tmKV, bar_kv_full + 8 * s, p.v_dx, j * T6_BN, t);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer (one thread) =================
        // order: QK_0(0), QK_1(0), then per tile j and unit u: [wait P_u(j)] P V, P 1, QK_u(j + 1).  tcgen05.mma executes in
        // issue order, so Q K_{j+1}^T (which overwrites S_u = P_j) cannot pass P_j V_j / P_j 1.
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(T6_BM, T6_BN, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(T6_BM, kDh, 0, 1);
            constexpr uint32_t idesc_pl = umma_idesc_bf16(T6_BM, 16, 0, 0);
            const uint64_t descKx = umma_desc_sw32(sKx), descOnes = umma_desc_ones(sOnes);
            auto issue_qk = [&](int u, int j) {
                const uint32_t s = (uint32_t)j % T6_STAGES;
                const uint64_t descQ = umma_desc_sw64(sQ + u * T6_Q_BYTES), descQx = umma_desc_sw32(sQx + u * T6_QX_BYTES);
                const uint64_t descK = umma_desc_sw64(sKV + s * T6_STAGE_BYTES);
                const uint32_t d = tmem + u * T6_UNIT_COLS + T6_COL_S;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                    umma_ss(d, descQ + (uint64_t)(kk * 2), descK + (uint64_t)(kk * 2), idesc_qk, kk > 0);
                umma_ss(d, descQx, descKx, idesc_qk, 1);  // S -= m (per-row reference maximum)
                tc_commit(bar_s + 8 * u);
            };
            mbar_wait_hint(bar_q, 0, 1000);
            mbar_wait_hint(bar_kv_full, 0, 1000);
            tc_fence_after();
            issue_qk(0, 0);
            issue_qk(1, 0);
            for (int j = 0; j < ntiles; ++j) {
                const uint32_t s = (uint32_t)j % T6_STAGES;
                const uint64_t descV = umma_desc_sw64(sKV + s * T6_STAGE_BYTES + T6_TILE_BYTES);
#pragma unroll
                for (int u = 0; u < T6_UNITS; ++u) {
                    mbar_wait_hint(bar_p + 8 * u, (uint32_t)j & 1u, p.probe_ns);
                    tc_fence_after();
                    const uint32_t ub = tmem + u * T6_UNIT_COLS;
#pragma unroll
                    for (int kk = 0; kk < T6_BN / 16; ++kk) {  // 16 keys per step: 8 TMEM columns of P, 1 KB of V
                        umma_ts(ub + T6_COL_O, ub + T6_COL_S + kk * 8, descV + (uint64_t)(kk * 64), idesc_pv, (j > 0) || (kk > 0));
                        umma_ts(ub + T6_COL_L, ub + T6_COL_S + kk * 8, descOnes, idesc_pl, (j > 0) || (kk > 0));
                    }
                    if (u == T6_UNITS - 1) tc_commit(bar_kv_empty + 8 * s);  // both units have used the stage's V tile
                    tc_commit(bar_pv + 8 * u);
                    if (j + 1 == ntiles) {
                        tc_commit(bar_o + 8 * u);
                    } else {
                        if (u == 0) {  // unit 1 follows on the same tile: the stage has landed by then
                            const uint32_t s1 = (uint32_t)(j + 1) % T6_STAGES;
                            mbar_wait_hint(bar_kv_full + 8 * s1, ((uint32_t)(j + 1) / T6_STAGES) & 1u, 1000);
                            tc_fence_after();
                        }
                        issue_qk(u, j + 1);
                    }
                }
            }
        }
    } else {
        // ================= softmax warps: warp w -> unit w / 4, TMEM lanes 32 (w % 4) .. + 31; thread = query row ====
        const int u = warp >> 2, wq = warp & 3;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const uint32_t scol = tmem + lane_base + u * T6_UNIT_COLS + T6_COL_S;
        const uint32_t orow = tmem + lane_base + u * T6_UNIT_COLS + T6_COL_O;
        const uint32_t lrow = tmem + lane_base + u * T6_UNIT_COLS + T6_COL_L;
        constexpr float kMargin = 7.0f;  // reference = rounded running maximum + 7: P <= 2^-6.5 until the maximum grows by 8
        uint8_t* qx_row = t6_smem_raw + (sQx - raw) + u * T6_QX_BYTES + sw32_chunk0(wq * 32 + lane);
        const int h = 2 * hp + u;
        // m_cur: integer-valued reference of the O and L accumulators; m_q: value baked into the NEXT tile's S (last
        // value written to this row's Q extension; 0 before the first write)
        float m_cur = 0.f, m_q = 0.f;
        auto publish = [&](float m) {  // Q K_{j+1}^T is issued only after this thread's arrival on bar_p: no race
            const float hi = 256.0f * rintf(m * (1.0f / 256.0f));
            *reinterpret_cast<uint32_t*>(qx_row) = pack_bf16x2(-hi, -(m - hi));
            m_q = m;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        };
        auto rescale_acc = [&](float corr, int j) {  // O and l of this row *= corr (after O/L += P_{j-1} V_{j-1} landed)
            mbar_wait(bar_pv + 8 * u, (uint32_t)(j - 1) & 1u);
            tc_fence_after();
            uint32_t ov[32], lv;
            tmem_ld32(orow, ov);
            tmem_ld1(lrow, lv);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
            tmem_st32(orow, ov);
            tmem_st1(lrow, __float_as_uint(__uint_as_float(lv) * corr));
        };
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(bar_s + 8 * u, (uint32_t)j & 1u);
            tc_fence_after();
            const int nvalid = (int)min((int64_t)T6_BN, p.N - (int64_t)j * T6_BN);
            const float baked = m_q;  // what Q K_j^T subtracted
            uint32_t pk[32];
            bool fast = !__any_sync(0xffffffffu, j == 0 || nvalid != T6_BN);
            if (fast) {
                // a reference published without rescaling (large-but-finite tile below): adopt it now
                const bool adopt = baked != m_cur;
                if (__any_sync(0xffffffffu, adopt)) {
                    rescale_acc(adopt ? fast_exp2(m_cur - baked) : 1.0f, j);
                    m_cur = baked;
                }
                uint32_t sa[32], ovf = 0;
                tmem_ld32(scol, sa);
                tmem_wait_ld();
                softmax_exp32_v6<POLY16, DEG>(sa, pk, ovf);
                tmem_ld32(scol + 32, sa);
                tmem_wait_ld();
                softmax_exp32_v6<POLY16, DEG>(sa, pk + 16, ovf);
                if (__any_sync(0xffffffffu, (ovf & 0xC000C000u) != 0u)) {
                    // some P >= 2 (or NaN, or a wrapped exponent): rare.  Below 2^64 the tile is kept as it is (P is only
                    // large, not wrong) and a higher reference is published for the next tile; otherwise redo.
                    uint32_t mx = 0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = max(mx, max(pk[i] & 0xffffu, pk[i] >> 16));
                    if (__any_sync(0xffffffffu, mx >= 0x5F80u)) {  // >= 2^64, inf, NaN or sign bit
                        fast = false;                            // S is still intact in TMEM
                        if (p.dbg && lane == 0) atomicAdd(p.dbg + 0, 1ull);
                    } else if (mx >= 0x4000u) {
                        publish(m_cur + (float)((int)(mx >> 7) - 127 + 8));
                        if (p.dbg) atomicAdd(p.dbg + 1, 1ull);
                    }
                }
            }
            if (!fast) {
                if (p.dbg && lane == 0) atomicAdd(p.dbg + 2, 1ull);
                // general path: explicit tile maximum, lazy rescaling (first tile of an item, partial tiles, redone tiles)
                uint32_t sa[32], sb[32];
                tmem_ld32(scol, sa);
                tmem_ld32(scol + 32, sb);
                tmem_wait_ld();
                if (nvalid < T6_BN) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= nvalid) sa[i] = 0xff800000u;  // -inf
                        if (32 + i >= nvalid) sb[i] = 0xff800000u;
                    }
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    mx0 = fmax3(mx0, __uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1]));
                    mx1 = fmax3(mx1, __uint_as_float(sb[2 * i]), __uint_as_float(sb[2 * i + 1]));
                }
                const float mt = fmaxf(mx0, mx1) + baked;  // tile maximum in absolute (scaled) units
                const bool need = (j == 0) || (mt > m_cur + (8.0f - kMargin)) || (baked != m_cur);
                if (__any_sync(0xffffffffu, need)) {
                    float corr = 1.0f;
                    if (need) {
                        const float m_new = fmaxf(rintf(mt) + kMargin, j == 0 ? -INFINITY : fmaxf(m_cur, baked));
                        if (j > 0) {
                            if (p.dbg && m_new != m_cur) atomicAdd(p.dbg + 1, 1ull);
                            corr = fast_exp2(m_cur - m_new);
                        }
                        m_cur = m_new;
                    }
                    if (j > 0) rescale_acc(corr, j);
                }
                // P = exp2(S - (m_cur - baked)) as bf16 pairs (row sums come from the tensor core); the second piece is
                // re-read instead of being kept live across the first (register budget of two CTAs per SM)
                const float delta = m_cur - baked;
                const float cm = kExpMagic - delta, smin = delta - 125.0f;
                uint64_t l2 = 0;
                softmax_exp32<POLY16, DEG>(sa, pk, 1.0f, -delta, cm, smin, l2);
                tmem_ld32(scol + 32, sa);
                tmem_wait_ld();
                if (nvalid < T6_BN) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (32 + i >= nvalid) sa[i] = 0xff800000u;
                }
                softmax_exp32<POLY16, DEG>(sa, pk + 16, 1.0f, -delta, cm, smin, l2);
                if (m_cur != m_q) publish(m_cur);
            }
            tmem_st32(scol, reinterpret_cast<uint32_t(&)[32]>(pk));
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p + 8 * u);
        }
        // ---- epilogue: O / l -> bf16 -> global ----
        mbar_wait(bar_o + 8 * u, 0);
        tc_fence_after();
        uint32_t ov[32], lv;
        tmem_ld32(orow, ov);
        tmem_ld1(lrow, lv);
        tmem_wait_ld();
        const int64_t r = (int64_t)qt * T6_BM + wq * 32 + lane;
        if (r < p.R) {
            const float inv = 1.0f / __uint_as_float(lv);
            uint4* dst = reinterpret_cast<uint4*>(p.O + r * p.o_row + (int64_t)t * p.o_tok + h * kDh);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(ov[8 * q + 0]) * inv, __uint_as_float(ov[8 * q + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(ov[8 * q + 2]) * inv, __uint_as_float(ov[8 * q + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(ov[8 * q + 4]) * inv, __uint_as_float(ov[8 * q + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(ov[8 * q + 6]) * inv, __uint_as_float(ov[8 * q + 7]) * inv);
                dst[q] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)T6_TMEM_COLS) : "memory");
}

template <int POLY16, int DEG>
static inline cudaError_t launch_attn_tc6_impl(const CUtensorMap& mq, const CUtensorMap& mkv, const Tc6Args& p, dim3 grid,
                                               cudaStream_t st) {
    static bool configured_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc6_kernel<POLY16, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, T6_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    attn_tc6_kernel<POLY16, DEG><<<grid, T6_THREADS, T6_SMEM_BYTES, st>>>(mq, mkv, p);
    return cudaGetLastError();
}

// test rows against the cached head-0 K/V only (a.k_head == 0); `heads` must be even
static inline cudaError_t launch_attn_tc6(const AttnArgs& a, int heads, int T, int poly, uint32_t probe_ns, unsigned long long* dbg,
                                          cudaStream_t st) {
    if (a.k_head != 0 || (heads & 1)) return cudaErrorInvalidValue;
    CUtensorMap mq, mkv;
    if (!make_map3(&mq, a.Q, (uint64_t)a.q_tok, (uint64_t)T, (uint64_t)a.R, (uint64_t)a.q_tok * 2, (uint64_t)a.q_row * 2, kDh, 1, T6_BM))
        return cudaErrorInvalidValue;
    if (!make_map3(&mkv, a.K, (uint64_t)kKvRow, (uint64_t)a.N, (uint64_t)T, (uint64_t)a.k_row * 2, (uint64_t)a.k_tok * 2, kDh, T6_BN, 1))
        return cudaErrorInvalidValue;
    Tc6Args p{};
    p.O = a.O; p.o_row = a.o_row; p.o_tok = a.o_tok; p.R = a.R; p.N = a.N; p.v_dx = a.v_off;
    p.head_pairs = heads / 2;
    p.n_qtiles = (int)ceil_div(a.R, T6_BM);
    p.dbg = dbg;
    p.probe_ns = probe_ns;
    dim3 grid((unsigned)((int64_t)T * p.n_qtiles * p.head_pairs));
    switch (poly) {
        case 0: return launch_attn_tc6_impl<0, 3>(mq, mkv, p, grid, st);
        case 4: return launch_attn_tc6_impl<4, 3>(mq, mkv, p, grid, st);
        case 5: return launch_attn_tc6_impl<5, 3>(mq, mkv, p, grid, st);
        case 6: return launch_attn_tc6_impl<6, 3>(mq, mkv, p, grid, st);
        case 7: return launch_attn_tc6_impl<7, 3>(mq, mkv, p, grid, st);
        case 8: return launch_attn_tc6_impl<8, 3>(mq, mkv, p, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pfn
