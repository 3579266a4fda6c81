// Attention between items on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Same contract as attn_mma.cuh (AttnArgs): one CTA = 128 query vectors (dh = 32) of one (column t, head h)
// against all N keys of that column, streamed in tiles of 64 keys.  Warp roles (192 threads):
//   warps 0-3  softmax: thread i owns query row i = TMEM lane i.  Reads S from TMEM (tcgen05.ld), online
//              softmax in fp32 registers with lazy rescaling, writes P (bf16) back into the S columns
//              (tcgen05.st), so the second contraction takes P straight from TMEM.
//   warp 4     TMA producer: Q tile once, then K/V tiles through an 8-stage mbarrier ring
//              (cp.async.bulk.tensor, 64-byte swizzle = the canonical K-major / MN-major UMMA layouts).
//   warp 5     TMEM allocation + single-thread MMA issue: S_j = Q K_j^T (M128 N64 K32, operands in shared
//              memory) and O += P_j V_j (M128 N32 K64, A from TMEM, B = V tile MN-major); completion is
//              signalled with tcgen05.commit on mbarriers.
// S is double-buffered in TMEM and the MMA thread runs one tile ahead (QK_{j+1} is issued before the softmax of
// tile j finishes), so the exponential pipe (MUFU, the real bound at dh = 32: 128 FLOP per ex2) never waits
// for the tensor pipe.  Three CTAs are resident per SM (128 + 32 TMEM columns each), so twelve softmax warps
// keep the MUFU pipe busy through each other's TMEM-load / max / store / barrier phases.
// r1 profile of the first version (single S buffer, 128-key tiles): profiles/r1_ncu_attn_tc_v1_metrics.txt.
#pragma once
#include <cuda.h>

#include "attn_mma.cuh"
#include "common.cuh"

namespace pfn {

#ifndef PFN_ATTN_BN
#define PFN_ATTN_BN 64   // keys per tile (64 or 32)
#endif
#ifndef PFN_ATTN_CTAS
#define PFN_ATTN_CTAS 3  // resident CTAs per SM the kernel is compiled for
#endif
// PFN_ATTN_WG256 (experiment, round 2): CTA of two warpgroups, warps 4-7 (TMA, MMA, two idle) give registers back with
// setmaxnreg so that the softmax warpgroup runs with 120 instead of 96 registers (the pool is per CTA: 128 x 120 + 128 x 40 = 256 x 80) at the same 3 CTAs per SM
#ifdef PFN_ATTN_WG256
constexpr int TC_THREADS = 256;
#else
constexpr int TC_THREADS = 192;
#endif
constexpr int TC_BM = 128, TC_BN = PFN_ATTN_BN, TC_STAGES = TC_BN == 64 ? 7 : 8;
static_assert(TC_BN == 64 || TC_BN == 32, "key tile must be 32 or 64");
constexpr int TC_Q_BYTES = TC_BM * kDh * 2;           // 8 KB: 128 rows x 64 B
constexpr int TC_TILE_BYTES = TC_BN * kDh * 2;        // 4 KB: 64 keys x 64 B
constexpr int TC_STAGE_BYTES = 2 * TC_TILE_BYTES;     // K tile + V tile
// 72 960 B: three CTAs per SM fit in shared memory (219 KB), a fourth does not.  TMEM: each CTA takes 128 columns
// (S0/P0 [0,64), S1/P1 [64,128)) plus a separate 32-column allocation for O: 3 x 160 = 480 of the 512 columns.
// + the "lean" softmax's third K = 16 block of the Q K^T contraction (32-byte-swizzle rows of 16 bf16): a per-row
// column pair (-m_hi, -m_lo) on the Q side and a constant (1, 1, 0, ...) on the key side, so S arrives as s - m.
constexpr int TC_QX_BYTES = TC_BM * 32, TC_KX_BYTES = TC_BN * 32;
constexpr int TC_SMEM_BYTES = 1024 + TC_Q_BYTES + TC_STAGES * TC_STAGE_BYTES + TC_QX_BYTES + TC_KX_BYTES + 256;
constexpr int TC_TMEM_COLS_S = 2 * TC_BN, TC_TMEM_COLS_O = 32;

struct TcArgs {
    bf16* O;
    int64_t o_row, o_tok;
    int64_t R, N;
    int kv_ctx_mode;  // 0: K/V map dims (64, N, T) box (32,64,1); 1: dims (ld, T, N) box (32,1,64)
    int kx0, kx_head, v_dx;  // x coordinate of K for head 0, per-head step, V = K + v_dx
    int heads, n_qtiles;     // work item = (column t, query tile, head): item = (t * n_qtiles + qt) * heads + h
    int64_t items;
    uint32_t wait_ticks;     // suspend-time hint of the producer / MMA threads' barrier waits (0 = none)
    uint32_t stagger_ns;     // persistent mode: residency slot k (blockIdx / #SMs) starts k * stagger_ns later
    uint32_t num_sms;
    unsigned long long* dbg;  // optional event counters: [0] fast-path tiles redone after the overflow check fired,
                              // [1] reference-maximum changes after the first tile, [2] tiles on the general path
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// same, with a suspend-time hint (ns) so the hardware parks the thread longer before reporting "not yet"
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t ticks) {
    if (ticks == 0) { mbar_wait(bar, parity); return; }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITH_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAITH_DONE;\n\t"
        "bra WAITH_LOOP;\n\t"
        "WAITH_DONE:\n\t}" ::"r"(bar), "r"(parity), "r"(ticks) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc),
        "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 64-byte swizzle, tile = rows of 64 bytes, 8-row groups 512 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64))
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;            // leading byte offset (unused for one swizzle atom in the leading dim)
    d |= (uint64_t)(512 >> 4) << 32;   // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;            // SWIZZLE_64B
    return d;
}
// same for a K = 16 block stored as rows of 32 bytes with the 32-byte swizzle (8-row groups 256 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;  // SWIZZLE_32B
    return d;
}
// byte offset of logical 16-byte chunk 0 of row r in such a block (Swizzle<1,4,3>: address bit 4 ^= bit 7)
__device__ __forceinline__ uint32_t sw32_chunk0(int r) { return (uint32_t)(r * 32 + (((r >> 2) & 1) << 4)); }
// instruction descriptor kind::f16: D fp32, A/B bf16 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 do two lanes per issue slot) and the 3-input maximum ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr float kExpMagic = 12582912.0f;  // 1.5 * 2^23: adding it leaves round(x) in the low mantissa bits

// P = 2^(S * sc - m) for one 32-column piece of a score tile, as 16 bf16 pairs, row sum accumulated in l2 (two lanes).
// The reference maximum m is INTEGER valued (see the caller), so both forms below are exact in their argument:
//   MUFU pair:  x = fma(S, sc, -m) (one FFMA2), two ex2.approx;
//   FMA pair:   Cody-Waite on the FMA pipes, two lanes per instruction.  With C = magic - m (an exact integer),
//               t = fma(S, sc, C) = magic + n, n = round(S sc - m);  w = C - t = -(m + n) exactly;
//               f = fma(S, sc, w) in [-0.5, 0.5];  2^f by a minimax polynomial (degree 3: 7.5e-5, degree 2: 1.7e-3
//               relative; P is rounded to bf16 = 3.9e-3 afterwards);  the exponent n is added with one integer
//               shift-add per lane ((magic + n) << 23 = n << 23 mod 2^32).  S is clamped below at smin so that
//               n >= -125 (also maps the -inf of masked keys to 2^-125 ~ 0).
// POLY16 of every 16 pairs take the FMA form, spread evenly, so the MUFU pipe (16 ex2 / clk / SM, the bound of
// this kernel) and the FMA pipes work side by side.
// NPAIR pairs starting at pair I0 of a 16-pair pattern (the FMA-form pairs are those i with
// ((I0 + i + 1) * POLY16) / 16 != ((I0 + i) * POLY16) / 16).
// `between(i)` runs after pair i (hook for work to interleave with the exponentials; tools/softmax_bench.cu).
struct NoBetween { __device__ __forceinline__ void operator()(int) const {} };
template <int POLY16, int DEG, int NPAIR, int I0, typename Between = NoBetween>
__device__ __forceinline__ void softmax_exp(const uint32_t* s, uint32_t* pk, float sc, float neg_m, float cm, float smin,
                                            uint64_t& l2, Between between = Between()) {
    const uint64_t SC = pk2(sc, sc), NM = pk2(neg_m, neg_m), CM = pk2(cm, cm), NEG1 = pk2(-1.0f, -1.0f);
#pragma unroll
    for (int i = 0; i < NPAIR; ++i) {
        const bool poly = ((I0 + i + 1) * POLY16) / 16 != ((I0 + i) * POLY16) / 16;
        float p0, p1;
        if (!poly) {
            float x0, x1;
            upk2(ffma2(pk2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), SC, NM), x0, x1);
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
        } else {
            const uint64_t S2 = pk2(fmaxf(__uint_as_float(s[2 * i]), smin), fmaxf(__uint_as_float(s[2 * i + 1]), smin));
            const uint64_t t2 = ffma2(S2, SC, CM);
            const uint64_t w2 = ffma2(t2, NEG1, CM);
            const uint64_t f2 = ffma2(S2, SC, w2);
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
        l2 = fadd2(l2, pk2(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
        between(i);
    }
}
template <int POLY16, int DEG, typename Between = NoBetween>
__device__ __forceinline__ void softmax_exp32(const uint32_t (&s)[32], uint32_t* pk, float sc, float neg_m, float cm,
                                              float smin, uint64_t& l2, Between between = Between()) {
    softmax_exp<POLY16, DEG, 16, 0>(s, pk, sc, neg_m, cm, smin, l2, between);
}

// "Lean" form: the scores arrive as x = s - m already (scale folded into the Q projection, -m contributed by the
// third K = 16 block of the Q K^T MMA), so a MUFU pair is two ex2 and nothing else, and the tile maximum is replaced by
// an overflow check: the tile's own row sum (accumulated from zero) reaching 2 <=> some P >= 2 is possible <=> a score
// may have exceeded the reference by >= 1 (a NaN sum counts as overflow); `ovf` ORs the packed pairs of the FMA form,
// whose exponent can wrap into the sign bit (mask 0x80008000).  The caller then redoes / re-references the tile.
template <int POLY16, int DEG>
__device__ __forceinline__ void softmax_exp32_lean(const uint32_t (&s)[32], uint32_t* pk, uint64_t& l2, uint32_t& ovf) {
    const uint64_t CM = pk2(kExpMagic, kExpMagic), NEG1 = pk2(-1.0f, -1.0f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool poly = ((i + 1) * POLY16) / 16 != (i * POLY16) / 16;
        float p0, p1;
        if (!poly) {
            p0 = fast_exp2(__uint_as_float(s[2 * i]));
            p1 = fast_exp2(__uint_as_float(s[2 * i + 1]));
        } else {
            // The clamp at -125 costs 20 FMNMX per tile.  Dropping it (-DPFN_ATTN_LEAN_NOCLAMP) measured +1.5 % (528 vs 519
            // TFLOP/s, profiles/r2_attn_tc6_experiments.txt) but is only safe while no score lies 512 or more log2 units below
            // the reference: beyond that the 9 exponent bits kept by `<< 23` wrap to a small POSITIVE value that neither the
            // sign bits nor the row-sum check see (the sharp-score test reaches such spreads).  Kept.
#ifndef PFN_ATTN_LEAN_NOCLAMP
            const uint64_t X2 = pk2(fmaxf(__uint_as_float(s[2 * i]), -125.0f), fmaxf(__uint_as_float(s[2 * i + 1]), -125.0f));
#else
            const uint64_t X2 = pk2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
#endif
            const uint64_t t2 = fadd2(X2, CM);            // magic + n, n = round(x)
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
        l2 = fadd2(l2, pk2(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
        if (poly) ovf |= pk[i];  // only the FMA form can wrap its exponent into the sign bit; P >= 2 shows in the tile sum
    }
}

// POLY16 = 0: every exponential on MUFU; k > 0: k of every 16 pairs on the FMA pipes (softmax_exp32).
// LEAN = 1: attn_tc v5 (see softmax_exp32_lean); the reference maximum m_cur (integer valued, = rounded running maximum
// + 7) is baked into S two tiles ahead, a tile takes the fast path when the value baked into it equals m_cur, and the
// general path (explicit maximum pass, as v4) handles the first two tiles of an item, partial tiles, and the two tiles
// after a reference change.
template <int POLY16, int DEG, int LEAN>
#ifdef PFN_ATTN_MAXNREG
__global__ void __maxnreg__(PFN_ATTN_MAXNREG)
#else
__global__ void __launch_bounds__(TC_THREADS, PFN_ATTN_CTAS)
#endif
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const TcArgs p) {
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sKV = base + TC_Q_BYTES;
    const uint32_t sQx = sKV + TC_STAGES * TC_STAGE_BYTES;
    const uint32_t sKx = sQx + TC_QX_BYTES;
    const uint32_t bars = sKx + TC_KX_BYTES;
    // barrier slots (8 B each)
    const uint32_t bar_kv_full = bars;                     // [TC_STAGES]  TMA -> MMA
    const uint32_t bar_kv_empty = bars + 8 * TC_STAGES;    // [TC_STAGES]  MMA -> TMA
    const uint32_t bar_q = bars + 16 * TC_STAGES;          // Q tile landed
    const uint32_t bar_s = bar_q + 8;                      // [2] S buffer b holds Q K_j^T          MMA -> softmax
    const uint32_t bar_p = bar_q + 24;                     // [2] P_j written into S buffer b       softmax -> MMA
    const uint32_t bar_pv = bar_q + 40;                    // O += P_j V_j finished (rescale guard) MMA -> softmax
    const uint32_t bar_o = bar_q + 48;                     // last O += P V finished
    const uint32_t bar_q_empty = bar_q + 56;               // the item's S MMAs have finished reading Q
    const uint32_t tmem_slot = bar_q + 64;                 // [2]: S/P columns, O columns
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(tc_smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (int)((p.N + TC_BN - 1) / TC_BN);
    const int per_col = p.n_qtiles * p.heads;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_kv_full + 8 * s, 1);
            mbar_init(bar_kv_empty + 8 * s, 1);
        }
        mbar_init(bar_q, 1);
        mbar_init(bar_q_empty, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_s + 8, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_p + 8, 128);
        mbar_init(bar_pv, 1);
        mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (LEAN) {
        uint32_t* ext = reinterpret_cast<uint32_t*>(tc_smem_raw + (sQx - raw));
        for (int i = threadIdx.x; i < (TC_QX_BYTES + TC_KX_BYTES) / 4; i += TC_THREADS) ext[i] = 0u;
        __syncthreads();
        if (threadIdx.x < TC_BN)  // key side: ones in the two columns that carry -m_hi and -m_lo
            *reinterpret_cast<uint32_t*>(tc_smem_raw + (sKx - raw) + sw32_chunk0(threadIdx.x)) = 0x3F803F80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)TC_TMEM_COLS_S) : "memory");
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot + 4),
                     "r"((uint32_t)TC_TMEM_COLS_O) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot_ptr[0];     // S/P double buffer
    const uint32_t tmem_o = tmem_slot_ptr[1];   // O accumulator
#ifdef PFN_ATTN_WG256
#define PFN_WG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 40;")
#define PFN_WG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 120;")
    if (warp >= 4) PFN_WG_DEC();  // the whole helper warpgroup executes ONE setmaxnreg (it is warpgroup-aligned)
#else
#define PFN_WG_DEC()
#define PFN_WG_INC()
#endif

    if (p.stagger_ns) {  // de-phase the CTAs that share an SM (they would otherwise run their tiles in lockstep)
        const uint32_t slot = blockIdx.x / p.num_sms;
        if (slot) __nanosleep(slot * p.stagger_ns);
    }
    // Persistent: every role walks the same sequence of work items; barrier phases are carried across items through
    // running counters (g = key tiles processed so far, qn = items processed so far).
    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
            uint32_t g = 0, qn = 0;
            for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x, ++qn) {
                const int t = (int)(item / per_col), rem = (int)(item % per_col);
                const int qt = rem / p.heads, h = rem % p.heads;
                mbar_wait_hint(bar_q_empty, (qn & 1u) ^ 1u, p.wait_ticks);  // the previous item's S = Q K^T MMAs have finished reading Q
                mbar_expect_tx(bar_q, TC_Q_BYTES);
                tma_load_3d(sQ, &tmQ, bar_q, h * kDh, t, qt * TC_BM);
                const int kx = p.kx0 + h * p.kx_head;
                for (int j = 0; j < ntiles; ++j, ++g) {
                    const uint32_t s = g % TC_STAGES, ph = (g / TC_STAGES) & 1u;
                    mbar_wait_hint(bar_kv_empty + 8 * s, ph ^ 1u, p.wait_ticks);
                    mbar_expect_tx(bar_kv_full + 8 * s, TC_STAGE_BYTES);
                    const uint32_t dstK = sKV + s * TC_STAGE_BYTES, dstV = dstK + TC_TILE_BYTES;
                    const int key0 = j * TC_BN;
                    if (p.kv_ctx_mode) {
                        tma_load_3d(dstK, &tmKV, bar_kv_full + 8 * s, kx, t, key0);
                        tma_load_3d(dstV, &tmKV, bar_kv_full + 8 * s, kx + p.v_dx, t, key0);
                    } else {
                        tma_load_3d(dstK, &tmKV, bar_kv_full + 8 * s, kx, key0, t);
                        tma_load_3d(dstV, &tmKV, bar_kv_full + 8 * s, kx + p.v_dx, key0, t);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer (one thread) =================
        // Issue order inside an item: QK_0, QK_1, then per tile j: [wait P_j] PV_j, QK_{j+2}.  tcgen05.mma executes in
        // issue order, so QK_{j+2} (which overwrites S buffer = P_j) cannot pass PV_j, and the softmax warps always
        // find S_{j+1} ready when they finish tile j.  The next item's first PV overwrites O only after its own P_0
        // barrier, which every softmax thread reaches after it has read the previous item's O.
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(TC_BM, TC_BN, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(TC_BM, kDh, 0, 1);
            const uint64_t descQ = umma_desc_sw64(sQ);
            const uint64_t descQx = umma_desc_sw32(sQx), descKx = umma_desc_sw32(sKx);
            uint32_t g0 = 0, qn = 0;  // g0 = global index of this item's first key tile
            for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x, ++qn, g0 += (uint32_t)ntiles) {
                auto issue_qk = [&](int j) {
                    const uint32_t g = g0 + (uint32_t)j, s = g % TC_STAGES;
                    mbar_wait_hint(bar_kv_full + 8 * s, (g / TC_STAGES) & 1u, p.wait_ticks);
                    tc_fence_after();
                    const uint64_t descK = umma_desc_sw64(sKV + s * TC_STAGE_BYTES);
                    const uint32_t d = tmem + (g & 1u) * TC_BN;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)  // dh = 32 = 2 x K16; 32 bytes per step inside the swizzle atom
                        umma_ss(d, descQ + (uint64_t)(kk * 2), descK + (uint64_t)(kk * 2), idesc_qk, kk > 0);
                    if (LEAN) umma_ss(d, descQx, descKx, idesc_qk, 1);  // S -= m (per-row reference maximum)
                    tc_commit(bar_s + 8 * (g & 1u));
                    if (j == ntiles - 1) tc_commit(bar_q_empty);
                };
                mbar_wait_hint(bar_q, qn & 1u, p.wait_ticks);
                issue_qk(0);
                if (ntiles > 1) issue_qk(1);
                for (int j = 0; j < ntiles; ++j) {
                    const uint32_t g = g0 + (uint32_t)j, s = g % TC_STAGES, b = g & 1u;
                    mbar_wait_hint(bar_p + 8 * b, (g >> 1) & 1u, p.wait_ticks);
                    tc_fence_after();
                    const uint64_t descV = umma_desc_sw64(sKV + s * TC_STAGE_BYTES + TC_TILE_BYTES);
                    const uint32_t pa = tmem + b * TC_BN;
#pragma unroll
                    for (int kk = 0; kk < TC_BN / 16; ++kk)  // 16 keys per step: 8 TMEM columns of P, 1 KB of V
                        umma_ts(tmem_o, pa + kk * 8, descV + (uint64_t)(kk * 64), idesc_pv, (j > 0) || (kk > 0));
                    tc_commit(bar_kv_empty + 8 * s);
                    tc_commit(bar_pv);
                    if (j + 2 < ntiles) issue_qk(j + 2);
                }
                tc_commit(bar_o);
            }
        }
    } else if (warp < 4) {
        PFN_WG_INC();
        // ================= softmax warps: thread = query row = TMEM lane =================
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t orow = tmem_o + ((uint32_t)(warp * 32) << 16);
        const float sc = 1.0f;  // queries arrive pre-scaled by kItemScaleLog2 (folded into the projection weights)
        constexpr float kMargin = LEAN ? 7.0f : 0.0f;  // m_cur = rounded maximum + margin; a new reference is taken when a
                                                       // tile maximum exceeds m_cur + 8 - margin (lazy rescaling, 2^8)
        uint8_t* qx_row = tc_smem_raw + (sQx - raw) + sw32_chunk0(warp * 32 + lane);
        uint32_t g = 0, qn = 0;
        for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x, ++qn) {
            const int t = (int)(item / per_col), rem = (int)(item % per_col);
            const int qt = rem / p.heads, h = rem % p.heads;
            const int64_t m0 = (int64_t)qt * TC_BM;
            // m_cur: integer-valued reference of O and l (scaled units).  LEAN: m_q = value last written to this row's Q
            // extension, mb[b] = value baked into the S tile that buffer b currently holds (0 until the first write).
            float m_cur = 0.f, m_q = 0.f, mb0 = 0.f, mb1 = 0.f;
            uint64_t l2 = 0;
            for (int j = 0; j < ntiles; ++j, ++g) {
                const uint32_t b = g & 1u;
                const uint32_t scol = trow + b * TC_BN;
                mbar_wait(bar_s + 8 * b, (g >> 1) & 1u);
                tc_fence_after();
                constexpr bool kTwo = TC_BN == 64;
                const int nvalid = (int)min((int64_t)TC_BN, p.N - (int64_t)j * TC_BN);
                const float baked = LEAN ? (b ? mb1 : mb0) : 0.f;
                uint32_t pk[kTwo ? 32 : 16];
                bool fast = false;
                // publish a new reference for the tiles whose Q K^T has not been issued yet (j + 2 onwards).  Q K_{j+1}^T
                // may still be reading the old value: wait for its completion barrier first.
                auto publish = [&](float m) {
                    if (j + 1 < ntiles) mbar_wait(bar_s + 8 * (b ^ 1u), ((g + 1u) >> 1) & 1u);
                    const float hi = 256.0f * rintf(m * (1.0f / 256.0f));
                    *reinterpret_cast<uint32_t*>(qx_row) = pack_bf16x2(-hi, -(m - hi));
                    m_q = m;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                };
                if constexpr (LEAN && kTwo) {
                    // LEAN 1: any row whose baked reference differs from m_cur sends the warp to the general path.
                    // LEAN 2: a row that only LAGS behind a reference it published itself (m_cur != m_q) adopts the
                    //         baked value with one rescale of O and l and stays on the fast path.
                    const bool stale = baked != m_cur && (LEAN == 1 || m_cur == m_q);
                    fast = !__any_sync(0xffffffffu, j == 0 || nvalid != TC_BN || stale);
                    if (fast) {
                        if constexpr (LEAN == 2) {
                            const bool adopt = baked != m_cur;
                            if (__any_sync(0xffffffffu, adopt)) {
                                float corr = 1.0f;
                                if (adopt) {
                                    corr = fast_exp2(m_cur - baked);
                                    float l0, l1;
                                    upk2(l2, l0, l1);
                                    l2 = pk2(l0 * corr, l1 * corr);
                                    m_cur = baked;
                                }
                                mbar_wait(bar_pv, (g - 1u) & 1u);  // O += P_{j-1} V_{j-1} must have landed
                                tc_fence_after();
                                uint32_t ov[32];
                                tmem_ld32(orow, ov);
                                tmem_wait_ld();
#pragma unroll
                                for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
                                tmem_st32(orow, ov);
                            }
                        }
                        uint32_t sa[32], ovf = 0;
                        uint64_t lt2 = 0;  // this tile's row sum, from zero
                        tmem_ld32(scol, sa);
                        tmem_wait_ld();
                        softmax_exp32_lean<POLY16, DEG>(sa, pk, lt2, ovf);
                        tmem_ld32(scol + 32, sa);
                        tmem_wait_ld();
                        softmax_exp32_lean<POLY16, DEG>(sa, pk + 16, lt2, ovf);
                        float ts0, ts1;
                        upk2(lt2, ts0, ts1);
                        const uint64_t l2n = fadd2(l2, lt2);
                        if (__any_sync(0xffffffffu, !(ts0 + ts1 < 2.0f) || (ovf & 0x80008000u) != 0u)) {
                            // some P >= 2 (or an exponent wrapped): rare.  LEAN 1 redoes the tile on the general path.
                            // LEAN 2 looks at the row's largest P: below 2^64 the tile is kept as it is (P is only
                            // large, not wrong) and a higher reference is published for two tiles ahead; otherwise redo.
                            bool redo = true;
                            if constexpr (LEAN == 2) {
                                uint32_t mx = 0;
#pragma unroll
                                for (int i = 0; i < 32; ++i) mx = max(mx, max(pk[i] & 0xffffu, pk[i] >> 16));
                                redo = __any_sync(0xffffffffu, mx >= 0x5F80u);  // >= 2^64, inf, NaN or sign bit
                                if (!redo && mx >= 0x4000u) {
                                    const float m_new = m_cur + (float)((int)(mx >> 7) - 127 + 8);
                                    if (m_new > m_q) {
                                        publish(m_new);
                                        if (p.dbg) atomicAdd(p.dbg + 1, 1ull);
                                    }
                                }
                            }
                            if (redo) {
                                fast = false;  // S is still intact in TMEM
                                if (p.dbg && lane == 0) atomicAdd(p.dbg + 0, 1ull);
                            } else {
                                l2 = l2n;
                            }
                        } else {
                            l2 = l2n;
                        }
                    }
                }
                if (!fast) {
                    if (p.dbg && lane == 0) atomicAdd(p.dbg + 2, 1ull);
                    // general path.  pass 1: the tile's 32-column pieces -> tile maximum (with 64-key tiles the second
                    // piece is re-read later instead of being kept live: register budget of 3 CTAs per SM)
                    uint32_t sa[32], sb[kTwo ? 32 : 1];
                    tmem_ld32(scol, sa);
                    if constexpr (kTwo) tmem_ld32(scol + 32, sb);
                    tmem_wait_ld();
                    if (nvalid < TC_BN) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i >= nvalid) sa[i] = 0xff800000u;  // -inf
                            if constexpr (kTwo)
                                if (32 + i >= nvalid) sb[i] = 0xff800000u;
                        }
                    }
                    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if constexpr (kTwo) {
                            mx0 = fmax3(mx0, __uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1]));
                            mx1 = fmax3(mx1, __uint_as_float(sb[2 * i]), __uint_as_float(sb[2 * i + 1]));
                        } else {
                            if (i & 1) mx1 = fmax3(mx1, __uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1]));
                            else mx0 = fmax3(mx0, __uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1]));
                        }
                    }
                    const float mt = fmaxf(mx0, mx1) * sc + baked;  // tile maximum in absolute (scaled) units
                    const bool need = (j == 0) || (mt > m_cur + (8.0f - kMargin));
                    if (__any_sync(0xffffffffu, need)) {
                        float corr = 1.0f;
                        if (need) {
                            const float m_new = rintf(mt) + kMargin;
                            if (j > 0) {
                                if (p.dbg) atomicAdd(p.dbg + 1, 1ull);
                                corr = fast_exp2(m_cur - m_new);
                                float l0, l1;
                                upk2(l2, l0, l1);
                                l2 = pk2(l0 * corr, l1 * corr);
                            }
                            m_cur = m_new;
                        }
                        if (j > 0) {
                            mbar_wait(bar_pv, (g - 1u) & 1u);  // O += P_{j-1} V_{j-1} must have landed
                            tc_fence_after();
                            uint32_t ov[32];
                            tmem_ld32(orow, ov);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
                            tmem_st32(orow, ov);
                        }
                    }
                    // pass 2: P = exp2(S - (m_cur - baked)) as bf16 pairs
                    const float delta = m_cur - baked;
                    const float cm = kExpMagic - delta, smin = (delta - 125.0f) * (1.0f / sc);
                    softmax_exp32<POLY16, DEG>(sa, pk, sc, -delta, cm, smin, l2);
                    if constexpr (kTwo) {
                        tmem_ld32(scol + 32, sa);
                        tmem_wait_ld();
                        if (nvalid < TC_BN) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (32 + i >= nvalid) sa[i] = 0xff800000u;
                        }
                        softmax_exp32<POLY16, DEG>(sa, pk + 16, sc, -delta, cm, smin, l2);
                    }
                    if (LEAN && need && m_cur != m_q) publish(m_cur);  // only a reference changed HERE is published (a
                                                                       // lagging LEAN-2 row keeps its pending, higher one)
                }
                if constexpr (kTwo) tmem_st32(scol, reinterpret_cast<uint32_t(&)[32]>(pk));
                else tmem_st16(scol, pk);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(bar_p + 8 * b);
                if (b) mb1 = m_q; else mb0 = m_q;  // buffer b next holds tile j + 2, issued after this arrival
            }
            // ---- epilogue: O / l -> bf16 -> global ----
            mbar_wait(bar_o, qn & 1u);
            tc_fence_after();
            uint32_t ov[32];
            tmem_ld32(orow, ov);
            tmem_wait_ld();
            const int64_t r = m0 + warp * 32 + lane;
            if (r < p.R) {
                float l0, l1;
                upk2(l2, l0, l1);
                const float inv = 1.0f / (l0 + l1);
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_TMEM_COLS_S)
                     : "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_o), "r"((uint32_t)TC_TMEM_COLS_O)
                     : "memory");
    }
}

// ---- host side: tensor maps + launch ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// 3-D bf16 map; dims/box innermost first; strides (bytes) of dims 1 and 2
static inline bool make_map3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                             uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int POLY16, int DEG, int LEAN>
static inline cudaError_t launch_attn_tc_impl(const CUtensorMap& mq, const CUtensorMap& mkv, const TcArgs& p, dim3 grid,
                                              cudaStream_t st) {
    static bool configured_dev[64] = {};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<POLY16, DEG, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TC_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    attn_tc_kernel<POLY16, DEG, LEAN><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(mq, mkv, p);
    return cudaGetLastError();
}

static inline cudaError_t launch_attn_tc(const AttnArgs& a, int heads, int T, int poly_mod, int num_sms, int persist,
                                         uint32_t wait_ticks, uint32_t stagger_ns, unsigned long long* dbg, cudaStream_t st) {
    CUtensorMap mq, mkv;
    TcArgs p{};
    p.O = a.O; p.o_row = a.o_row; p.o_tok = a.o_tok; p.R = a.R; p.N = a.N;
    // queries: element (x, t, r) at Q + r*q_row + t*q_tok + x
    if (!make_map3(&mq, a.Q, (uint64_t)a.q_tok, (uint64_t)T, (uint64_t)a.R, (uint64_t)a.q_tok * 2, (uint64_t)a.q_row * 2,
                   kDh, 1, TC_BM))
        return cudaErrorInvalidValue;
    if (a.k_head) {
        // context rows: keys live in the projected qkv buffer, element (x, t, j) at base + j*k_row + t*k_tok + x
        const bf16* basep = a.K - kE;  // a.K points at the K block (column kE) of the qkv rows
        if (!make_map3(&mkv, basep, (uint64_t)a.k_tok, (uint64_t)T, (uint64_t)a.N, (uint64_t)a.k_tok * 2,
                       (uint64_t)a.k_row * 2, kDh, 1, TC_BN))
            return cudaErrorInvalidValue;
        p.kv_ctx_mode = 1; p.kx0 = kE; p.kx_head = a.k_head; p.v_dx = a.v_off;
    } else {
        // cached head-0 K/V: element (x, j, t) at base + t*k_tok + j*k_row + x, rows of 64 (K | V)
        if (!make_map3(&mkv, a.K, (uint64_t)kKvRow, (uint64_t)a.N, (uint64_t)T, (uint64_t)a.k_row * 2,
                       (uint64_t)a.k_tok * 2, kDh, TC_BN, 1))
            return cudaErrorInvalidValue;
        p.kv_ctx_mode = 0; p.kx0 = 0; p.kx_head = 0; p.v_dx = a.v_off;
    }
    p.heads = heads;
    p.n_qtiles = (int)ceil_div(a.R, TC_BM);
    p.items = (int64_t)T * p.n_qtiles * heads;
    p.wait_ticks = wait_ticks;
    p.stagger_ns = persist ? stagger_ns : 0;
    p.num_sms = (uint32_t)num_sms;
    p.dbg = dbg;
    // persist: 3 CTAs per SM walk the items; otherwise one CTA per item (the hardware scheduler staggers them)
    dim3 grid((unsigned)(persist ? std::min<int64_t>(p.items, 3 * (int64_t)num_sms) : p.items));
    // poly_mod = k + 100 * (degree == 2) + 1000 * lean: k of every 16 exponential pairs on the FMA pipes
#define PFN_ATTN_CASE(K)                                                            \
    case K: return launch_attn_tc_impl<K, 3, 0>(mq, mkv, p, grid, st);              \
    case 1000 + K: return launch_attn_tc_impl<K, 3, 1>(mq, mkv, p, grid, st);
    switch (poly_mod) {
        PFN_ATTN_CASE(0) PFN_ATTN_CASE(3) PFN_ATTN_CASE(4) PFN_ATTN_CASE(5) PFN_ATTN_CASE(6) PFN_ATTN_CASE(7) PFN_ATTN_CASE(8)
        case 106: return launch_attn_tc_impl<6, 2, 0>(mq, mkv, p, grid, st);
        case 1106: return launch_attn_tc_impl<6, 2, 1>(mq, mkv, p, grid, st);
        case 2004: return launch_attn_tc_impl<4, 3, 2>(mq, mkv, p, grid, st);
        case 2005: return launch_attn_tc_impl<5, 3, 2>(mq, mkv, p, grid, st);
        case 2006: return launch_attn_tc_impl<6, 3, 2>(mq, mkv, p, grid, st);
        case 2000: return launch_attn_tc_impl<0, 3, 2>(mq, mkv, p, grid, st);
        default: return cudaErrorInvalidValue;
    }
#undef PFN_ATTN_CASE
}

}  // namespace pfn
