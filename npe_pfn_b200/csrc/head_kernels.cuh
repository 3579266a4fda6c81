// Bar-distribution head, round-2 kernels (same arithmetic spec as oracle/bar_head.c and head_kernel in small_kernels.cuh,
// hence the same bits: bucket masses are integers, so every partition of the prefix sums gives the same bucket).
//
//  head_row_kernel   one CTA (8 warps) per logits row, the row lives in REGISTERS (5 coalesced float4 loads per
//                    thread, 20 000 B read exactly once, no shared-memory staging: head_kernel staged 20 KB per warp and
//                    ran 8 warps per SM at 10 % of the HBM peak).  exp_det + quantisation once per element; 128-element
//                    block totals by two REDUX (20-bit halves of the 40-bit masses) instead of shuffle trees; only the
//                    one block that holds the target is scanned.
//  head_cdf_kernel + head_shared_kernel   when MANY draws / targets share one logits row (dimension 0 of `sample`: all
//                    M rows are the same observation; `sample_batched`: n draws per observation; dimension 0 of
//                    `log_prob`): the integer CDF, max and log-partition of each distinct row are computed ONCE, then
//                    one thread per draw does a 13-step binary search (the old path recomputed 5 000 exponentials per draw).
#pragma once
#include "small_kernels.cuh"

namespace pfn {

#ifndef PFN_HEAD_THREADS
#define PFN_HEAD_THREADS 256
#endif
constexpr int HR_THREADS = PFN_HEAD_THREADS, HR_WARPS = HR_THREADS / 32, HR_VEC = 5120 / (HR_THREADS * 4), HR_MAX_B = HR_THREADS * HR_VEC * 4;  // 5120 buckets

__device__ __forceinline__ int float_ordered(float f) {  // monotone float -> int map (for REDUX max)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ unsigned long long warp_scan_u64(unsigned long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// sum over the warp of a value < 2^42 that is itself a sum of <= 4 masses <= 2^40: two 32-bit REDUX on its 20-bit halves
__device__ __forceinline__ unsigned long long warp_sum_q(unsigned long long s) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(s & 0xFFFFFull));
    const unsigned hi = __reduce_add_sync(0xffffffffu, (unsigned)(s >> 20));
    return ((unsigned long long)hi << 20) + (unsigned long long)lo;
}

template <bool SAMPLE>
__global__ void __launch_bounds__(HR_THREADS, 1024 / HR_THREADS) head_row_kernel(HeadArgs a) {
    __shared__ int s_max[HR_WARPS];
    __shared__ unsigned long long s_w[HR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = a.B;
    const int nvec = B >> 2;  // B % 4 == 0
    const double kLn2x40 = 40.0 * 0.6931471805599453;
    for (int64_t r = blockIdx.x; r < a.M; r += gridDim.x) {
        const float* lg = a.logits + (r / a.group) * a.ld_logits;
        const float4* lg4 = reinterpret_cast<const float4*>(lg);
        // element order: warp w owns float4 indices [32 HR_VEC w, 32 HR_VEC (w + 1)), block k of the warp = indices 32 HR_VEC w + 32 k + lane
        float v[4 * HR_VEC];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < HR_VEC; ++k) {
            const int f = warp * (32 * HR_VEC) + k * 32 + lane;
            float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (f < nvec) x = __ldg(lg4 + f);
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
            mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
        }
        const int wmx = __reduce_max_sync(0xffffffffu, float_ordered(mx));
        __syncthreads();  // previous row's readers of s_max / s_w are done
        if (lane == 0) s_max[warp] = wmx;
        __syncthreads();
        int mi = s_max[0];
#pragma unroll
        for (int w = 1; w < HR_WARPS; ++w) mi = max(mi, s_max[w]);
        const float m = ordered_float(mi);
        // block totals T[k] (uniform across the warp) and the warp total
        unsigned long long T[HR_VEC], wsum = 0ull;
#pragma unroll
        for (int k = 0; k < HR_VEC; ++k) {
            const int f = warp * (32 * HR_VEC) + k * 32 + lane;
            unsigned long long s = 0ull;
            if (f < nvec)
                s = quantize_q40(exp_det(v[4 * k] - m)) + quantize_q40(exp_det(v[4 * k + 1] - m)) +
                    quantize_q40(exp_det(v[4 * k + 2] - m)) + quantize_q40(exp_det(v[4 * k + 3] - m));
            T[k] = warp_sum_q(s);
            wsum += T[k];
        }
        if (lane == 0) s_w[warp] = wsum;
        __syncthreads();
        unsigned long long Z = 0ull;
#pragma unroll
        for (int w = 0; w < HR_WARPS; ++w) Z += s_w[w];

        if (!SAMPLE) {
            if (threadIdx.x == 0) {
                const double logZ = log((double)Z) - kLn2x40;
                const float yv = a.y[r * a.ld_y];
                const double logp = bar_logp(lg, B, a.borders, m, logZ, yv);
                if (a.out_nll) a.out_nll[r] = (float)(-logp);
                if (a.out_logp) {
                    float lp = (float)logp;
                    if (lp == -INFINITY) lp = a.log_eps;
                    a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
                }
            }
            continue;
        }

        float u;
        if (a.uniforms) u = a.uniforms[r];
        else {
            uint32_t c[4];
            philox4x32_10(a.seed, a.row0 + (uint64_t)r, a.offset, c);
            u = ((float)(c[0] >> 9) + 0.5f) * 0x1.0p-23f;
        }
        const double target = (double)u * (double)Z;
        // the warp that holds the target: first w whose inclusive total is not below it
        int wsel = HR_WARPS;
        unsigned long long run = 0ull;
        {
            unsigned long long c = 0ull;
#pragma unroll
            for (int w = 0; w < HR_WARPS; ++w) {
                const unsigned long long cw = c + s_w[w];
                if (wsel == HR_WARPS && !((double)cw < target)) { wsel = w; run = c; }
                c = cw;
            }
        }
        if (wsel == HR_WARPS) {  // unreachable for u < 1 (spec: last bucket, no mass)
            if (threadIdx.x == 0) {
                const int idx = B - 1;
                a.out_theta[r * a.ld_theta] = a.borders[idx];
                if (a.out_bin) a.out_bin[r] = idx;
                if (a.out_u) a.out_u[r] = u;
                if (a.out_logp) {
                    const double logZ = log((double)Z) - kLn2x40;
                    float lp = (float)bar_logp(lg, B, a.borders, m, logZ, a.borders[idx]);
                    if (lp == -INFINITY) lp = a.log_eps;
                    a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
                }
            }
            continue;
        }
        if (warp != wsel) continue;
        // block of the warp that holds the target
        int ksel = HR_VEC - 1;
        {
            unsigned long long c = run;
            bool found = false;
#pragma unroll
            for (int k = 0; k < HR_VEC; ++k) {
                if (!found) {
                    if (!((double)(c + T[k]) < target)) { ksel = k; found = true; }
                    else c += T[k];
                }
            }
            run = c;
        }
        // the four masses of this lane in that block (compile-time register indices, one predicated branch taken)
        unsigned long long q[4] = {0ull, 0ull, 0ull, 0ull};
        float lv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < HR_VEC; ++k) {
            if (k == ksel) {
                const int f = warp * (32 * HR_VEC) + k * 32 + lane;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    lv[c] = v[4 * k + c];
                    q[c] = f < nvec ? quantize_q40(exp_det(v[4 * k + c] - m)) : 0ull;
                }
            }
        }
        const unsigned long long s4 = q[0] + q[1] + q[2] + q[3];
        const unsigned long long inc = warp_scan_u64(s4, lane) + run;  // inclusive prefix up to this lane's last element
        const unsigned below = __ballot_sync(0xffffffffu, (double)inc < target);
        const int lsel = __popc(below);  // first lane whose inclusive prefix is not below the target (< 32 by construction)
        if (lane == min(lsel, 31)) {
            unsigned long long Cprev = inc - s4, qsel = 0ull;
            int j = 3;
            float lsel_v = lv[3];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (j == 3 && c < 3) {
                    if (!((double)(Cprev + q[c]) < target)) { j = c; }
                    else Cprev += q[c];
                }
            }
            qsel = q[j];
            lsel_v = lv[j];
            const int idx = 4 * (warp * (32 * HR_VEC) + ksel * 32 + lane) + j;
            double frac = qsel ? (target - (double)Cprev) / (double)qsel : 0.0;
            frac = fmin(fmax(frac, 0.0), 1.0);
            const double lo = (double)a.borders[idx], hi = (double)a.borders[idx + 1];
            const float th = (float)(lo + (hi - lo) * frac);
            a.out_theta[r * a.ld_theta] = th;
            if (a.out_bin) a.out_bin[r] = idx;
            if (a.out_u) a.out_u[r] = u;
            if (a.out_logp) {
                const double logZ = log((double)Z) - kLn2x40;
                (void)lsel_v;
                float lp = (float)bar_logp(lg, B, a.borders, m, logZ, th);
                if (lp == -INFINITY) lp = a.log_eps;
                a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
            }
        }
    }
}

// ---- round-2 second generation: the same row kernel with (a) the NEXT row prefetched by one bulk copy (TMA,
// `cp.async.bulk` + mbarrier) into a 20 KB shared-memory landing buffer while the current row is processed in registers,
// persistent CTAs (4 per SM); (b) the bucket mass in 15 instructions instead of 28, with the SAME bits as the specification
// in oracle/bar_head.c (tools/head_arith_check.c walks every fp32 argument in [-64, 0]):
//   floor(x)           x <= 0, x >= -93:  tr = RD(x + (1.5 * 2^23 + 40)) is exactly 1.5 * 2^23 + 40 + floor(x) (ulp = 1 in
//                      [2^23, 2^24)), so n = tr - magic is floorf(x) as a float and the low bits of tr are n + 40 as an integer
//                      (FRND + F2I -> FADD.RM + FADD);
//   p * 2^n * 2^40     p in [1, 2] and n + 40 + 127 > 0: multiplying by a power of two only moves the exponent field, i.e.
//                      bits(e * 2^40) = bits(p) + (tr_bits << 23)   (the magic's own low 9 bits are zero);
//   trunc(e * 2^40)    one F2I.U64.TRUNC of that fp32 value (< 2^42) instead of mantissa / exponent shifts.
__device__ __forceinline__ unsigned long long mass_q40(float v, float m) {
    const float t = fmaxf(__fsub_rn(v, m), -64.0f);  // NaN -> -64 like the specification's `!(t > -64)`
    const float x = __fmul_rn(t, 0x1.715476p+0f);
    const float kMagic = 12582912.0f + 40.0f;
    const float tr = __fadd_rd(x, kMagic);
    const float n = __fsub_rn(tr, kMagic);
    const float f = __fsub_rn(x, n);
    float p = 0x1.ca8f0ap-13f;
    p = __fmaf_rn(p, f, 0x1.44d4d2p-10f);
    p = __fmaf_rn(p, f, 0x1.3d54d8p-7f);
    p = __fmaf_rn(p, f, 0x1.c67f50p-5f);
    p = __fmaf_rn(p, f, 0x1.ebfdf2p-3f);
    p = __fmaf_rn(p, f, 0x1.62e428p-1f);
    p = __fmaf_rn(p, f, 0x1.000000p+0f);
    return __float2ull_rz(__uint_as_float(__float_as_uint(p) + (__float_as_uint(tr) << 23)));
}

__device__ __forceinline__ void bulk_load_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void head_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "HB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HB_DONE;\n\t"
        "bra HB_WAIT;\n\t"
        "HB_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}


// Comparisons of the specification are `(double)C < target` with target = (double)u * (double)Z.  C is an integer below
// 2^53, so C < target  <=>  C < ceil(target): every search below compares integers against tceil = (u64)ceil(target).
// THREADS per logits row (a CTA), CTAS resident per SM: <256, 4> = 5 float4 per thread, 64 registers; <128, 5> = 10 float4 per
// thread (twice the independent work per warp, half the per-warp bookkeeping per row, four instead of eight warps per barrier).
template <bool SAMPLE, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) head_row2_kernel(HeadArgs a) {
    constexpr int HR_THREADS = THREADS, HR_WARPS = THREADS / 32, HR_VEC = 5120 / (THREADS * 4);
    static_assert(HR_VEC * THREADS * 4 == HR_MAX_B && (THREADS & (THREADS - 1)) == 0, "threads per row must divide 1280 float4 evenly");
    __shared__ __align__(128) float s_row[HR_MAX_B];  // landing buffer of the bulk copy: the row AFTER the one in registers
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) int s_max[2][HR_WARPS];                 // double-buffered by row parity: two barriers per row
    __shared__ __align__(16) unsigned long long s_w[2][HR_WARPS];
    __shared__ float s_u[HR_THREADS];  // uniforms of this CTA's next HR_THREADS rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = a.B;
    const uint32_t row_bytes = (uint32_t)B * 4u, bar = smem_u32(&s_bar), dst = smem_u32(s_row);
    const double kLn2x40 = 40.0 * 0.6931471805599453;
    const bool grouped = a.group != 1;  // 64-bit division only when draws share logits rows
    const int64_t stride = gridDim.x;
    // the copy fills s_row[0, B); the tail stays -inf for the whole kernel, so loads and masses need no bounds checks
    // (exp_det clamps -inf to its floor, whose mass is 0)
    for (int i = B + (int)threadIdx.x; i < HR_MAX_B; i += HR_THREADS) s_row[i] = -INFINITY;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int64_t r0 = blockIdx.x;
        if (r0 < a.M) bulk_load_row(dst, a.logits + (grouped ? r0 / a.group : r0) * a.ld_logits, row_bytes, bar);
    }
    __syncthreads();
    uint32_t phase = 0;
    int par = 0;
    const float4* s4 = reinterpret_cast<const float4*>(s_row) + warp * (32 * HR_VEC) + lane;
    int it = 0;
    for (int64_t r = blockIdx.x; r < a.M; r += stride, par ^= 1, ++it) {
        // in the shadow of the copy: address of the next row (thread 0); every HR_THREADS rows of this CTA, one uniform per
        // thread for the next HR_THREADS rows (Philox is ~90 instructions: once per thread instead of once per warp and row,
        // and no warp arrives late at the first barrier because of it)
        const float* next = nullptr;
        if (threadIdx.x == 0 && r + stride < a.M) next = a.logits + (grouped ? (r + stride) / a.group : r + stride) * a.ld_logits;
        if (SAMPLE && (it & (HR_THREADS - 1)) == 0) {
            __syncthreads();  // readers of the previous batch are done
            const int64_t rr = r + (int64_t)threadIdx.x * stride;
            float u0 = 0.f;
            if (rr < a.M) {
                if (a.uniforms) u0 = a.uniforms[rr];
                else {
                    uint32_t c[4];
                    philox4x32_10(a.seed, a.row0 + (uint64_t)rr, a.offset, c);
                    u0 = ((float)(c[0] >> 9) + 0.5f) * 0x1.0p-23f;
                }
            }
            s_u[threadIdx.x] = u0;  // published by the first barrier below
        }
        head_bar_wait(bar, phase);
        phase ^= 1u;
        // element order: warp w owns float4 indices [32 HR_VEC w, 32 HR_VEC (w + 1)), block k of the warp = indices 32 HR_VEC w + 32 k + lane
        float v[4 * HR_VEC];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < HR_VEC; ++k) {
            const float4 x = s4[k * 32];
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
            mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
        }
        const int wmx = __reduce_max_sync(0xffffffffu, float_ordered(mx));
        if (lane == 0) s_max[par][warp] = wmx;
        __syncthreads();  // the row has left s_row and the warp maxima are in place
        if (next) bulk_load_row(dst, next, row_bytes, bar);
        int mi = s_max[par][0];
#pragma unroll
        for (int w = 1; w < HR_WARPS; ++w) mi = max(mi, s_max[par][w]);
        const float m = ordered_float(mi);
        unsigned long long sk[HR_VEC], tsum = 0ull;  // this thread's mass per block and in total
#pragma unroll
        for (int k = 0; k < HR_VEC; ++k) {
            sk[k] = mass_q40(v[4 * k], m) + mass_q40(v[4 * k + 1], m) + mass_q40(v[4 * k + 2], m) + mass_q40(v[4 * k + 3], m);
            tsum += sk[k];
        }
        // warp total: tsum < 20 * 2^40 + ..., split at bit 22 so that both 32-lane REDUX sums stay below 2^32
        const unsigned long long wsum = ((unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)(tsum >> 22)) << 22) +
                                        (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)(tsum & 0x3FFFFFull));
        if (lane == 0) s_w[par][warp] = wsum;
        __syncthreads();
        unsigned long long Z = 0ull, run = 0ull;  // partition sum; mass of the warps before this one
#pragma unroll
        for (int w = 0; w < HR_WARPS; ++w) {
            const unsigned long long x = s_w[par][w];
            if (w < warp) run += x;
            Z += x;
        }

        if (!SAMPLE) {
            if (threadIdx.x == 0) {
                const float* lg = a.logits + (grouped ? r / a.group : r) * a.ld_logits;
                const double logZ = log((double)Z) - kLn2x40;
                const double logp = bar_logp(lg, B, a.borders, m, logZ, a.y[r * a.ld_y]);
                if (a.out_nll) a.out_nll[r] = (float)(-logp);
                if (a.out_logp) {
                    float lp = (float)logp;
                    if (lp == -INFINITY) lp = a.log_eps;
                    a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
                }
            }
            continue;
        }

        const float u = s_u[it & (HR_THREADS - 1)];
        const double target = (double)u * (double)Z;
        const unsigned long long tceil = __double2ull_ru(target);  // target >= 0
        // this warp holds the target iff its inclusive total is the first one not below it
        const bool mine = run + wsum >= tceil && (warp == 0 || run < tceil);
        int idx = -1;
        double frac = 0.0;
        if (tceil > Z) {  // unreachable for u < 1 (spec: last bucket, no mass)
            if (threadIdx.x == 0) idx = B - 1;
        } else if (mine) {
            int ksel = HR_VEC - 1;  // block of the warp that holds the target
            {
                bool found = false;
#pragma unroll
                for (int k = 0; k < HR_VEC; ++k) {
                    const unsigned long long Tk = warp_sum_q(sk[k]);
                    if (!found) {
                        if (run + Tk >= tceil) { ksel = k; found = true; }
                        else run += Tk;
                    }
                }
            }
            float lv0 = 0.f, lv1 = 0.f, lv2 = 0.f, lv3 = 0.f;  // this lane's four logits of that block
            unsigned long long s4q = 0ull;
#pragma unroll
            for (int k = 0; k < HR_VEC; ++k)
                if (k == ksel) { lv0 = v[4 * k]; lv1 = v[4 * k + 1]; lv2 = v[4 * k + 2]; lv3 = v[4 * k + 3]; s4q = sk[k]; }
            const unsigned long long inc = warp_scan_u64(s4q, lane) + run;  // inclusive prefix up to this lane's last element
            const int lsel = __popc(__ballot_sync(0xffffffffu, inc < tceil));  // first lane not below the target
            if (lane == min(lsel, 31)) {
                const unsigned long long q0 = mass_q40(lv0, m), q1 = mass_q40(lv1, m), q2 = mass_q40(lv2, m), q3 = mass_q40(lv3, m);
                unsigned long long Cprev = inc - s4q, qsel = q3;
                int j = 3;
                if (Cprev + q0 >= tceil) { j = 0; qsel = q0; }
                else if (Cprev + q0 + q1 >= tceil) { j = 1; qsel = q1; Cprev += q0; }
                else if (Cprev + q0 + q1 + q2 >= tceil) { j = 2; qsel = q2; Cprev += q0 + q1; }
                else Cprev += q0 + q1 + q2;
                idx = 4 * (warp * (32 * HR_VEC) + ksel * 32 + lane) + j;
                frac = qsel ? (target - (double)Cprev) / (double)qsel : 0.0;
                frac = fmin(fmax(frac, 0.0), 1.0);
            }
        }
        if (idx >= 0) {  // exactly one thread of the CTA
            const double lo = (double)a.borders[idx], hi = (double)a.borders[idx + 1];
            const float th = (float)(lo + (hi - lo) * frac);
            a.out_theta[r * a.ld_theta] = th;
            if (a.out_bin) a.out_bin[r] = idx;
            if (a.out_u) a.out_u[r] = u;
            if (a.out_logp) {
                const float* lg = a.logits + (grouped ? r / a.group : r) * a.ld_logits;
                const double logZ = log((double)Z) - kLn2x40;
                float lp = (float)bar_logp(lg, B, a.borders, m, logZ, th);
                if (lp == -INFINITY) lp = a.log_eps;
                a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
            }
        }
    }
}

// ---- shared logits rows: integer CDF once per distinct row, then one thread per draw / target ---------------------------
struct HeadCdf {            // per distinct logits row
    unsigned long long* cdf;  // [rows][B] inclusive prefix sums of the bucket masses
    float* row_max;           // [rows]
    double* logZ;             // [rows]
};

constexpr int HC_THREADS = 256;
__global__ void __launch_bounds__(HC_THREADS) head_cdf_kernel(const float* __restrict__ logits, int64_t ld, int B, HeadCdf o) {
    __shared__ float s_mx[HC_THREADS / 32];
    __shared__ unsigned long long s_ws[HC_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x;
    const float* lg = logits + row * ld;
    const int per = (B + HC_THREADS - 1) / HC_THREADS;  // contiguous buckets per thread
    const int i0 = min((int)threadIdx.x * per, B), i1 = min(i0 + per, B);
    float mx = -INFINITY;
    for (int i = i0; i < i1; ++i) mx = fmaxf(mx, lg[i]);
    mx = warp_max(mx);
    if (lane == 0) s_mx[warp] = mx;
    __syncthreads();
    float m = s_mx[0];
#pragma unroll
    for (int w = 1; w < HC_THREADS / 32; ++w) m = fmaxf(m, s_mx[w]);
    unsigned long long s = 0ull;
    for (int i = i0; i < i1; ++i) s += quantize_q40(exp_det(lg[i] - m));
    const unsigned long long inc = warp_scan_u64(s, lane);
    if (lane == 31) s_ws[warp] = inc;
    __syncthreads();
    unsigned long long before = 0ull, Z = 0ull;
#pragma unroll
    for (int w = 0; w < HC_THREADS / 32; ++w) {
        if (w < warp) before += s_ws[w];
        Z += s_ws[w];
    }
    unsigned long long c = before + inc - s;
    unsigned long long* out = o.cdf + row * (int64_t)B;
    for (int i = i0; i < i1; ++i) {
        c += quantize_q40(exp_det(lg[i] - m));
        out[i] = c;
    }
    if (threadIdx.x == 0) {
        o.row_max[row] = m;
        o.logZ[row] = log((double)Z) - 40.0 * 0.6931471805599453;
    }
}

// log density of y given the row's max and log partition, reading the one logit it needs from global memory
__device__ __forceinline__ double bar_logp_shared(const float* __restrict__ lg, int B, const float* __restrict__ borders, float m,
                                                  double logZ, float y) {
    return bar_logp(lg, B, borders, m, logZ, y);
}

template <bool SAMPLE>
__global__ void __launch_bounds__(256) head_shared_kernel(HeadArgs a, HeadCdf o) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.M) return;
    const int B = a.B;
    const int64_t row = a.ld_logits == 0 ? 0 : r / a.group;
    const float* lg = a.logits + row * a.ld_logits;
    const float m = o.row_max[row];
    const double logZ = o.logZ[row];
    if (!SAMPLE) {
        const double logp = bar_logp(lg, B, a.borders, m, logZ, a.y[r * a.ld_y]);
        if (a.out_nll) a.out_nll[r] = (float)(-logp);
        if (a.out_logp) {
            float lp = (float)logp;
            if (lp == -INFINITY) lp = a.log_eps;
            a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
        }
        return;
    }
    const unsigned long long* __restrict__ C = o.cdf + row * (int64_t)B;
    float u;
    if (a.uniforms) u = a.uniforms[r];
    else {
        uint32_t c[4];
        philox4x32_10(a.seed, a.row0 + (uint64_t)r, a.offset, c);
        u = ((float)(c[0] >> 9) + 0.5f) * 0x1.0p-23f;
    }
    const unsigned long long Z = C[B - 1];
    const double target = (double)u * (double)Z;
    int lo_i = 0, hi_i = B;  // first i with !(C_i < target)
    while (lo_i < hi_i) {
        const int mid = (lo_i + hi_i) >> 1;
        if ((double)C[mid] < target) lo_i = mid + 1; else hi_i = mid;
    }
    int idx = lo_i;
    unsigned long long Cprev, qsel;
    if (idx >= B) { idx = B - 1; Cprev = Z; qsel = 0ull; }
    else { Cprev = idx ? C[idx - 1] : 0ull; qsel = C[idx] - Cprev; }
    double frac = qsel ? (target - (double)Cprev) / (double)qsel : 0.0;
    frac = fmin(fmax(frac, 0.0), 1.0);
    const double lo = (double)a.borders[idx], hi = (double)a.borders[idx + 1];
    const float th = (float)(lo + (hi - lo) * frac);
    a.out_theta[r * a.ld_theta] = th;
    if (a.out_bin) a.out_bin[r] = idx;
    if (a.out_u) a.out_u[r] = u;
    if (a.out_logp) {
        float lp = (float)bar_logp(lg, B, a.borders, m, logZ, th);
        if (lp == -INFINITY) lp = a.log_eps;
        a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
    }
}

}  // namespace pfn
