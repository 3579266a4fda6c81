// HBM-bound kernels around the transformer: context statistics (fit), the per-cell feature encoder,
// attention between features (tiny T), the head-0 K/V cache writer, the bar-distribution head
// (softmax -> integer CDF -> inverse-CDF sample / log density) and support check + ordered compaction.
//
// The head follows the arithmetic spec of oracle/bar_head.c exactly (deterministic exp, 2^40 fixed
// point probabilities, integer prefix sums) so bucket indices and samples are bit-identical to the
// oracle for identical logits and uniforms.
#pragma once
#include "common.cuh"

namespace pfn {

constexpr int kMaxFeat = 128;  // 2 * max_groups
constexpr int kEncFloats = 2 * kMaxFeat + kMaxFeat / 2 + 4;
// slot "enc" buffer layout (floats): mean[128] | std[128] | scale[64] | y_mean, y_std, y_fill, pad
constexpr int kEncMean = 0, kEncStd = kMaxFeat, kEncScale = 2 * kMaxFeat, kEncY = 2 * kMaxFeat + kMaxFeat / 2;

// ---------------------------------------------------------------------------------------------
// fit statistics (oracle/tabpfn_oracle.py::EncoderStats,
// oracle/estimator.py::y_standardise).  One block per padded feature column + one for y.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_double(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += sh[w];  // fixed order: deterministic
    return t;
}

__global__ void __launch_bounds__(256) fit_stats_kernel(const float* __restrict__ X, int64_t ldx,
                                                        const float* __restrict__ y, int64_t N, int F, int G,
                                                        float* __restrict__ enc, int standardize_y) {
    __shared__ double sh[8];
    const int c = blockIdx.x;
    const int Fp = 2 * G;
    if (c < Fp) {
        if (c >= F) {  // zero padding column: constant by construction
            if (threadIdx.x == 0) { enc[kEncMean + c] = 0.f; enc[kEncStd + c] = 0.f; }
            return;
        }
        double s = 0.0, cnt = 0.0;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            float v = X[i * ldx + c];
            if (isfinite(v)) { s += (double)v; cnt += 1.0; }
        }
        s = block_sum_double(s, sh);
        cnt = block_sum_double(cnt, sh);
        const double mean = s / fmax(cnt, 1.0);
        double q = 0.0;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            float v = X[i * ldx + c];
            double d = isfinite(v) ? (double)v - mean : 0.0;
            q += d * d;
        }
        q = block_sum_double(q, sh);
        if (threadIdx.x == 0) {
            double var = N > 1 ? q / (double)(N - 1) : 0.0;
            enc[kEncMean + c] = (float)mean;
            enc[kEncStd + c] = (float)sqrt(var);
        }
        return;
    }
    // y column
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s += (double)y[i];
    s = block_sum_double(s, sh);
    const double mean = s / (double)N;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        double d = (double)y[i] - mean;
        q += d * d;
    }
    q = block_sum_double(q, sh);
    // population standard deviation + 1e-20 like upstream's `np.std(y) + 1e-20`; a constant target keeps unit scale
    const double std_d = sqrt(q / (double)N);
    float mean32 = (float)mean;
    float std32 = std_d > 0.0 ? (float)(std_d + 1e-20) : 0.f;
    if (!isfinite(std32) || std32 == 0.f) std32 = 1.f;
    if (!standardize_y) { mean32 = 0.f; std32 = 1.f; }  // classifier: class indices enter the y-encoder as they are
    double z = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) z += (double)((y[i] - mean32) / std32);
    z = block_sum_double(z, sh);
    if (threadIdx.x == 0) {
        enc[kEncY + 0] = mean32;
        enc[kEncY + 1] = std32;
        enc[kEncY + 2] = (float)(z / (double)N);
        enc[kEncY + 3] = 0.f;
    }
}

// group scale sqrt(2 / #non-constant features) and the bucket borders in original units
__global__ void fit_finalize_kernel(float* __restrict__ enc, int G, const float* __restrict__ borders, int nb1,
                                    float* __restrict__ borders_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) {
        int used = (enc[kEncStd + 2 * i] != 0.f) + (enc[kEncStd + 2 * i + 1] != 0.f);
        enc[kEncScale + i] = sqrtf(2.0f / (float)max(used, 1));
    }
    if (i < nb1) borders_out[i] = fmaf(borders[i], enc[kEncY + 1], enc[kEncY + 0]);
}

// ---------------------------------------------------------------------------------------------
// cell encoder: rows x F raw features (+ y for context rows) -> token grid [R, T, E] (fp32 + bf16)
// (oracle/tabpfn_oracle.py::encode_x / encode_y_ctx / encode_y_test)
// ---------------------------------------------------------------------------------------------
// Thread layout (round 2): a block handles ENC_ROWS rows at once, 48 threads per row, each thread FOUR consecutive output
// features of every token - so the fp32 state goes out as float4 (a warp writes 512 contiguous bytes) and the bf16 copy as
// 8-byte words instead of the scalar 4- / 2-byte stores of round 1 (0.8 TB/s).  Same fmaf sequence per element: same bits.
constexpr int ENC_ROWS = 4, ENC_TPR = kE / 4;  // 48 threads per row
__global__ void __launch_bounds__(ENC_ROWS * ENC_TPR) encode_kernel(const float* __restrict__ X, int64_t ldx, int F, int G,
                                                    const float* __restrict__ y /* null: test rows */,
                                                    int64_t R, const float* __restrict__ enc,
                                                    const float* __restrict__ enc_x_w, const float* __restrict__ enc_y_w,
                                                    const float* __restrict__ enc_y_b, const float* __restrict__ pos_emb,
                                                    float* __restrict__ xf, bf16* __restrict__ xb) {
    const int rr = threadIdx.x / ENC_TPR, e0 = 4 * (threadIdx.x % ENC_TPR);
    const int T = G + 1;
    const int64_t r = (int64_t)blockIdx.x * ENC_ROWS + rr;
    if (r >= R) return;
    float4 wx[4];
    float2 wy[4];
    float by[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        wx[i] = *reinterpret_cast<const float4*>(enc_x_w + 4 * (e0 + i));
        wy[i] = *reinterpret_cast<const float2*>(enc_y_w + 2 * (e0 + i));
        by[i] = enc_y_b[e0 + i];
    }
    const float y_mean = enc[kEncY + 0], y_std = enc[kEncY + 1], y_fill = enc[kEncY + 2];
    const float* xr = X + r * ldx;
    float* of = xf + r * T * kE + e0;
    bf16* ob = xb + r * T * kE + e0;
    auto store = [&](int tok, const float (&v)[4]) {
        *reinterpret_cast<float4*>(of + tok * kE) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint2*>(ob + tok * kE) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    };
    for (int g = 0; g < G; ++g) {
        float f[2], ind[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 2 * g + j;
            float v = c < F ? xr[c] : 0.f;
            float id = 0.f;
            if (isnan(v)) id = -2.f;
            else if (isinf(v)) id = v > 0.f ? 2.f : 4.f;
            const float mean = enc[kEncMean + c], sd = enc[kEncStd + c];
            const float filled = isfinite(v) ? v : mean;
            float xn = (filled - mean) / (sd + 1e-16f);
            if (sd == 0.f) xn = 0.f;
            xn = fminf(fmaxf(xn, -100.f), 100.f);
            f[j] = xn * enc[kEncScale + g];
            ind[j] = id;
        }
        const float4 pe = *reinterpret_cast<const float4*>(pos_emb + g * kE + e0);
        const float pev[4] = {pe.x, pe.y, pe.z, pe.w};
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = f[0] * wx[i].x;
            v = fmaf(f[1], wx[i].y, v);
            v = fmaf(ind[0], wx[i].z, v);
            v = fmaf(ind[1], wx[i].w, v);
            out[i] = v + pev[i];
        }
        store(g, out);
    }
    float y0, y1;
    if (y) { y0 = (y[r] - y_mean) / y_std; y1 = 0.f; }
    else   { y0 = y_fill; y1 = -2.f; }
    float out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = fmaf(y1, wy[i].y, y0 * wy[i].x) + by[i];
    store(G, out);
}

// ---------------------------------------------------------------------------------------------
// attention between features: for every row, MHA over its T tokens (T = G+1 is tiny).
// One warp per (row, head); fp32 math on bf16 q/k/v.  qkv [R*T, 576] = q | k | v, head h at h*32.
// ---------------------------------------------------------------------------------------------
constexpr int FA_WARPS = 6;
__global__ void __launch_bounds__(FA_WARPS * 32) feature_attn_kernel(const bf16* __restrict__ qkv, int64_t R, int T,
                                                                     bf16* __restrict__ out) {
    extern __shared__ float fa_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sK = fa_smem + (size_t)warp * 2 * T * 33;
    float* sV = sK + T * 33;
    const float scale_log2 = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
    const int nj = (T + 31) >> 5;                                          // key groups of 32 (<= 3)
    for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
        const int h = warp;
        const bf16* base = qkv + r * T * 3 * kE + h * kDh;
        __syncwarp();
        for (int j = 0; j < T; ++j) {
            sK[j * 33 + lane] = __bfloat162float(base[(int64_t)j * 3 * kE + kE + lane]);
            sV[j * 33 + lane] = __bfloat162float(base[(int64_t)j * 3 * kE + 2 * kE + lane]);
        }
        __syncwarp();
        for (int i = 0; i < T; ++i) {
            const float qd = __bfloat162float(base[(int64_t)i * 3 * kE + lane]);
            float s[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
            for (int d = 0; d < kDh; ++d) {
                const float q = __shfl_sync(0xffffffffu, qd, d);
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int j = jj * 32 + lane;
                    if (jj < nj && j < T) s[jj] = fmaf(q, sK[j * 33 + d], s[jj]);
                }
            }
            float m = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                const int j = jj * 32 + lane;
                s[jj] = (jj < nj && j < T) ? s[jj] * scale_log2 : -INFINITY;
                m = fmaxf(m, s[jj]);
            }
            m = warp_max(m);
            float p[3], l = 0.f;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) { p[jj] = exp2f(s[jj] - m); l += p[jj]; }
            l = warp_sum(l);
            float o = 0.f;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                if (jj < nj) {
                    const int jmax = min(32, T - jj * 32);
                    for (int j = 0; j < jmax; ++j) {
                        const float pj = __shfl_sync(0xffffffffu, p[jj], j);
                        o = fmaf(pj, sV[(jj * 32 + j) * 33 + lane], o);
                    }
                }
            }
            out[(r * T + i) * kE + h * kDh + lane] = __float2bfloat16_rn(o / l);
        }
    }
}

// Tensor-core version for T <= 16 (F <= 30 features): one warp per (row, head), the whole T x T attention is
// one 16 x 16 tile: S = Q K^T is 2 n-tiles x 2 k-steps of m16n8k16, O = P V is 4 n-tiles x 1 k-step.  Fragments
// are loaded straight from global memory (the row's q/k/v are 64-byte runs), softmax on the accumulator quads.
constexpr int FAM_WARPS = 8;
__global__ void __launch_bounds__(FAM_WARPS * 32) feature_attn_mma_kernel(const bf16* __restrict__ qkv, int64_t R, int T,
                                                                          bf16* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int64_t w = (int64_t)blockIdx.x * FAM_WARPS + warp;
    if (w >= R * kHeads) return;
    const int64_t r = w / kHeads;
    const int h = (int)(w % kHeads);
    const bf16* base = qkv + r * T * 3 * kE + h * kDh;
    const int64_t ts = 3 * kE;  // token stride
    const int q0 = min(g, T - 1), q1 = min(g + 8, T - 1);
    auto ld32 = [](const bf16* ptr) { return *reinterpret_cast<const uint32_t*>(ptr); };
    // S = Q K^T
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
            const uint32_t b0 = ld32(base + key * ts + kE + 16 * kk + 2 * tq);
            const uint32_t b1 = ld32(base + key * ts + kE + 16 * kk + 8 + 2 * tq);
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
    // O = P V : B fragment (k = key, n = dh) gathered as two 16-bit loads per register
    const unsigned short* vb = reinterpret_cast<const unsigned short*>(base + 2 * kE);
    const int k0 = min(2 * tq, T - 1), k1 = min(2 * tq + 1, T - 1), k2 = min(2 * tq + 8, T - 1), k3 = min(2 * tq + 9, T - 1);
    const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int dh = 8 * j + g;
        const uint32_t b0 = (uint32_t)vb[k0 * ts + dh] | ((uint32_t)vb[k1 * ts + dh] << 16);
        const uint32_t b1 = (uint32_t)vb[k2 * ts + dh] | ((uint32_t)vb[k3 * ts + dh] << 16);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(o, pa, b0, b1);
        const int col = h * kDh + 8 * j + 2 * tq;
        if (g < T) *reinterpret_cast<uint32_t*>(out + (r * T + g) * kE + col) = pack_bf16x2(o[0] * inv0, o[1] * inv0);
        if (g + 8 < T) *reinterpret_cast<uint32_t*>(out + (r * T + g + 8) * kE + col) = pack_bf16x2(o[2] * inv1, o[3] * inv1);
    }
}

// head-0 K/V of context rows [r0, r0 + nr) -> cache [T][N][64] (K 0..31 | V 32..63) for one layer
__global__ void kv_cache_kernel(const bf16* __restrict__ qkv, int64_t nr, int T, int64_t r0, int64_t N,
                                bf16* __restrict__ cache) {
    // one thread per 16 bytes: (n, t, part in 0..7): parts 0-3 = K, 4-7 = V
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = nr * T * 8;
    if (idx >= total) return;
    const int part = (int)(idx & 7);
    const int64_t nt = idx >> 3;
    const int t = (int)(nt % T);
    const int64_t n = nt / T;
    const bf16* src = qkv + (n * T + t) * 3 * kE + (part < 4 ? kE + part * 8 : 2 * kE + (part - 4) * 8);
    bf16* dst = cache + ((int64_t)t * N + r0 + n) * kKvRow + part * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

// bf16 copy of the Q rows (first E output features) of every layer's item-attention projection, scaled by `scale`
__global__ void scale_item_q_kernel(const float* __restrict__ wf, bf16* __restrict__ wb, int64_t off, int L, float scale) {
    const int64_t per = (int64_t)kE * kE;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per * L) return;
    const int64_t e = off + (i / per) * 3 * per + (i % per);
    wb[e] = __float2bfloat16_rn(wf[e] * scale);
}

// ---------------------------------------------------------------------------------------------
// bar-distribution head (oracle/bar_head.c is the arithmetic spec)
// ---------------------------------------------------------------------------------------------
constexpr int kQShift = 40;

__device__ __forceinline__ float exp_det(float t) {
    if (!(t > -64.0f)) t = -64.0f;
    const float x = __fmul_rn(t, 0x1.715476p+0f);
    const float n = floorf(x);
    const float f = __fsub_rn(x, n);
    float p = 0x1.ca8f0ap-13f;
    p = __fmaf_rn(p, f, 0x1.44d4d2p-10f);
    p = __fmaf_rn(p, f, 0x1.3d54d8p-7f);
    p = __fmaf_rn(p, f, 0x1.c67f50p-5f);
    p = __fmaf_rn(p, f, 0x1.ebfdf2p-3f);
    p = __fmaf_rn(p, f, 0x1.62e428p-1f);
    p = __fmaf_rn(p, f, 0x1.000000p+0f);
    const float s = __uint_as_float((uint32_t)((int)n + 127) << 23);
    return __fmul_rn(p, s);
}

__device__ __forceinline__ unsigned long long quantize_q40(float e) {
    const uint32_t b = __float_as_uint(e);
    const int ex = (int)((b >> 23) & 0xff);
    if (ex == 0) return 0ull;
    const unsigned long long m = (unsigned long long)((b & 0x7fffffu) | 0x800000u);
    const int sh = ex - 127 - 23 + kQShift;
    if (sh >= 0) return m << sh;
    if (sh <= -64) return 0ull;
    return m >> (-sh);
}

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t row, uint64_t offset, uint32_t (&c)[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)row; c[1] = (uint32_t)(row >> 32); c[2] = (uint32_t)offset; c[3] = (uint32_t)(offset >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// first k in [0, B] with borders[k] >= y, minus one, clamped to [0, B-1]
__device__ __forceinline__ int bucket_of(const float* __restrict__ borders, int B, float y) {
    int lo = 0, hi = B + 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (borders[mid] < y) lo = mid + 1; else hi = mid;
    }
    return min(max(lo - 1, 0), B - 1);
}

__device__ __forceinline__ double halfnormal_logpdf(double v, double scale) {
    return log(2.0) - log(scale) - 0.5 * log(2.0 * 3.14159265358979323846) - 0.5 * (v / scale) * (v / scale);
}

// log density of y under the bar distribution given the row's max logit m and log partition
__device__ double bar_logp(const float* __restrict__ lg, int B, const float* __restrict__ borders, float m,
                           double logZ, float y) {
    const int idx = bucket_of(borders, B, y);
    const double width = (double)borders[idx + 1] - (double)borders[idx];
    double logp = ((double)lg[idx] - (double)m) - logZ - log(width);
    const double kIcdfHalf = 0.6744897501960817;
    if (idx == 0) {
        const double w0 = (double)borders[1] - (double)borders[0];
        double v = (double)borders[1] - (double)y;
        if (v < 1e-8) v = 1e-8;
        logp += halfnormal_logpdf(v, w0 / kIcdfHalf) + log(w0);
    } else if (idx == B - 1) {
        const double wl = (double)borders[B] - (double)borders[B - 1];
        double v = (double)y - (double)borders[B - 1];
        if (v < 1e-8) v = 1e-8;
        logp += halfnormal_logpdf(v, wl / kIcdfHalf) + log(wl);
    }
    return logp;
}

struct HeadArgs {
    const float* logits;
    int64_t ld_logits;  // 0 = every row reads the same logits row
    int64_t group;      // row r reads logits row r / group (sample_batched dimension 0: `group` draws per observation)
    int64_t M;
    int B;
    const float* borders;  // [B+1], original units
    const float* uniforms;
    uint64_t seed, row0, offset;
    float* out_theta;
    int64_t ld_theta;
    int32_t* out_bin;
    float* out_u;
    float* out_logp;
    float log_eps;
    int accumulate;
    const float* y;  // nll mode
    int64_t ld_y;
    float* out_nll;
};

constexpr int HEAD_WARPS = 4;

// one warp per row: stage the row in shared memory, max, integer partition sum, inverse CDF
template <bool SAMPLE>
__global__ void __launch_bounds__(HEAD_WARPS * 32) head_kernel(HeadArgs a) {
    extern __shared__ float head_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = a.B;
    float* row = head_smem + (size_t)warp * B;
    const double kLn2x40 = 40.0 * 0.6931471805599453;
    for (int64_t r = (int64_t)blockIdx.x * HEAD_WARPS + warp; r < a.M; r += (int64_t)gridDim.x * HEAD_WARPS) {
        const float* lg = a.logits + (r / a.group) * a.ld_logits;
        __syncwarp();
        float m = -INFINITY;
        for (int i = lane; i < B; i += 32) {
            const float v = lg[i];
            row[i] = v;
            m = fmaxf(m, v);
        }
        m = warp_max(m);
        __syncwarp();
        // bucket masses are integers, so partial sums are order-independent: every lane sums a CONTIGUOUS segment
        // of buckets (no shuffles in the pass over the row); the inclusive scan of the 32 segment sums gives the
        // partition sum and, when sampling, the segment that holds the target.
        const int seg = (B + 31) >> 5;
        const int s0 = min(lane * seg, B), s1 = min(s0 + seg, B);
        unsigned long long ssum = 0ull;
        for (int i = s0; i < s1; ++i) ssum += quantize_q40(exp_det(row[i] - m));
        unsigned long long sinc = ssum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, sinc, o);
            if (lane >= o) sinc += t;
        }
        const unsigned long long Z = __shfl_sync(0xffffffffu, sinc, 31);
        const double logZ = log((double)Z) - kLn2x40;

        if (!SAMPLE) {
            if (lane == 0) {
                const float yv = a.y[r * a.ld_y];
                const double logp = bar_logp(row, B, a.borders, m, logZ, yv);
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
            u = ((float)(c[0] >> 9) + 0.5f) * 0x1.0p-23f;  // strictly inside (0, 1)
        }
        const double target = (double)u * (double)Z;
        // first segment whose inclusive sum is not below the target (sums are non-decreasing)
        const unsigned sbal = __ballot_sync(0xffffffffu, (double)sinc < target);
        const int sseg = __popc(sbal);
        int idx = -1;
        unsigned long long Cprev = 0ull, qsel = 0ull;
        if (sseg < 32) {
            const unsigned long long segprev = __shfl_sync(0xffffffffu, sinc - ssum, sseg);  // mass before the segment
            const int b0 = min(sseg * seg, B), b1 = min(b0 + seg, B);
            unsigned long long run = segprev;
            for (int base = b0; base < b1 && idx < 0; base += 32) {
                const int i = base + lane;
                const unsigned long long q = i < b1 ? quantize_q40(exp_det(row[i] - m)) : 0ull;
                unsigned long long c = q;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long t = __shfl_up_sync(0xffffffffu, c, o);
                    if (lane >= o) c += t;
                }
                c += run;
                const bool below = (i < b1) && ((double)c < target);
                const int nbelow = __popc(__ballot_sync(0xffffffffu, below));
                const int nvalid = min(32, b1 - base);
                if (nbelow < nvalid) {
                    idx = base + nbelow;
                    const unsigned long long csel = __shfl_sync(0xffffffffu, c, nbelow);
                    qsel = __shfl_sync(0xffffffffu, q, nbelow);
                    Cprev = csel - qsel;
                }
                run = __shfl_sync(0xffffffffu, c, 31);
            }
        }
        if (idx < 0) { idx = B - 1; Cprev = Z; qsel = 0ull; }
        if (lane == 0) {
            double frac = qsel ? (target - (double)Cprev) / (double)qsel : 0.0;
            frac = fmin(fmax(frac, 0.0), 1.0);
            const double lo = (double)a.borders[idx], hi = (double)a.borders[idx + 1];
            const float th = (float)(lo + (hi - lo) * frac);
            a.out_theta[r * a.ld_theta] = th;
            if (a.out_bin) a.out_bin[r] = idx;
            if (a.out_u) a.out_u[r] = u;
            if (a.out_logp) {
                float lp = (float)bar_logp(row, B, a.borders, m, logZ, th);
                if (lp == -INFINITY) lp = a.log_eps;
                a.out_logp[r] = a.accumulate ? a.out_logp[r] + lp : lp;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// support check + ordered stream compaction (accept_reject_sampler.py:54-62, support_posterior.py:152-156)
//
// ONE kernel, one pass: every thread evaluates the accept predicate of its rows once, a warp ballot + popcount gives
// the position inside the warp, a shared-memory scan the position inside the block, and a decoupled look-back over the
// per-block aggregates the position in the whole stream (blocks take their index from an atomic ticket, so a block
// only ever waits for blocks that are already running).  Accepted rows are APPENDED to the output at a device-resident
// cursor, in proposal order, and the cursor / proposal counters are advanced by the last block - a rejection loop can
// enqueue round after round without the host ever reading a count (north_star 4).
// ---------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256, CP_ITEMS = 4, CP_TILE = CP_THREADS * CP_ITEMS;

struct CompactArgs {
    const float* theta; int64_t ld; int64_t M; int dim;
    const float* lo; const float* hi;       // box (any may be null)
    const uint8_t* mask;                    // optional precomputed predicate
    const float* score; const float* thr;   // optional: accept only rows with score[r] > *thr (device scalar)
    const float* logp;                      // optional per-row payload carried along with accepted rows
    int64_t* out_idx; float* out_rows; float* out_logp;
    int64_t out_ld;                         // row stride of out_rows (floats)
    int64_t capacity;                       // rows that fit into out_* (rows past it are counted, not written)
    int64_t* cursor;                        // device: [0] accepted so far (append position), [1] proposed so far
    int64_t* out_count;                     // optional: number accepted in THIS call
    unsigned long long* tile_state;         // [nblocks] (flag << 62 | value), zeroed before the launch
    unsigned int* ticket;                   // zeroed before the launch
};

__device__ __forceinline__ bool row_accepted(const CompactArgs& a, int64_t r) {
    bool ok = a.mask ? a.mask[r] != 0 : true;
    for (int j = 0; j < a.dim; ++j) {
        const float v = a.theta[r * a.ld + j];
        ok = ok && isfinite(v);
        if (a.lo) ok = ok && (v >= a.lo[j]);
        if (a.hi) ok = ok && (v <= a.hi[j]);
    }
    if (a.score) ok = ok && (a.score[r] > *a.thr);
    return ok;
}

__global__ void __launch_bounds__(CP_THREADS) compact_append_kernel(const CompactArgs a) {
    __shared__ int wsum[CP_THREADS / 32];
    __shared__ unsigned int s_tile;
    __shared__ long long s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const unsigned int ntiles = gridDim.x;
    // thread t owns CP_ITEMS CONSECUTIVE rows: order inside the tile = (thread, item)
    const int64_t r0 = (int64_t)tile * CP_TILE + (int64_t)threadIdx.x * CP_ITEMS;
    bool ok[CP_ITEMS];
    int mine = 0;
#pragma unroll
    for (int i = 0; i < CP_ITEMS; ++i) {
        ok[i] = (r0 + i < a.M) && row_accepted(a, r0 + i);
        mine += ok[i];
    }
    // exclusive scan of `mine` over the block: warp shuffle scan + scan of the warp sums
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        if (w < warp) before += wsum[w];
        total += wsum[w];
    }
    const int excl = before + inc - mine;
    // decoupled look-back: publish the aggregate, then sum the predecessors' states
    if (threadIdx.x == 0) {
        long long prefix = 0;
        if (tile == 0) {
            atomicExch(a.tile_state + 0, (2ull << 62) | (unsigned long long)total);
        } else {
            atomicExch(a.tile_state + tile, (1ull << 62) | (unsigned long long)total);
            for (long long p = (long long)tile - 1; p >= 0; --p) {
                unsigned long long st;
                do { st = atomicAdd(a.tile_state + p, 0ull); } while ((st >> 62) == 0ull);
                prefix += (long long)(st & ((1ull << 62) - 1ull));
                if ((st >> 62) == 2ull) break;
            }
            atomicExch(a.tile_state + tile, (2ull << 62) | (unsigned long long)(prefix + total));
        }
        s_prefix = prefix;
    }
    __syncthreads();
    const int64_t base = a.cursor ? a.cursor[0] : 0;
    int64_t pos = base + s_prefix + excl;
#pragma unroll
    for (int i = 0; i < CP_ITEMS; ++i) {
        if (ok[i]) {
            if (pos < a.capacity) {
                const int64_t r = r0 + i;
                if (a.out_idx) a.out_idx[pos] = r;
                if (a.out_rows)
                    for (int j = 0; j < a.dim; ++j) a.out_rows[pos * a.out_ld + j] = a.theta[r * a.ld + j];
                if (a.out_logp && a.logp) a.out_logp[pos] = a.logp[r];
            }
            ++pos;
        }
    }
    // the last tile in TICKET order knows the grand total; it may only advance the cursor after every block has read
    // it, so the cursor update is done by whichever block finishes LAST (second ticket counter)
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(a.ticket + 1, 1u) == ntiles - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        unsigned long long st;
        do { st = atomicAdd(a.tile_state + (ntiles - 1), 0ull); } while ((st >> 62) != 2ull);
        const long long grand = (long long)(st & ((1ull << 62) - 1ull));
        if (a.out_count) *a.out_count = grand;
        if (a.cursor) { a.cursor[0] = base + grand; a.cursor[1] += a.M; }
    }
}

// row r of dst[rows, ld] <- src[n] in its first n columns (the observation repeated down the joint test matrix)
__global__ void broadcast_rows_kernel(const float* __restrict__ src, int n, int64_t rows, float* __restrict__ dst, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n) return;
    dst[(i / n) * ld + (i % n)] = src[i % n];
}

// uniform proposals on a box (`BoxUniform.sample`, support_posterior.py:137 / :305-309): out[r, j] = lo[j] + u (hi[j] - lo[j]),
// u from Philox4x32-10 keyed by (seed; counter = (row0 + r, j / 4)), 23-bit mantissa, strictly inside (0, 1)
__global__ void uniform_box_kernel(const float* __restrict__ lo, const float* __restrict__ hi, int64_t M, int dim, uint64_t seed,
                                   uint64_t row0, float* __restrict__ out, int64_t ld) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    for (int j0 = 0; j0 < dim; j0 += 4) {
        uint32_t c[4];
        philox4x32_10(seed, row0 + (uint64_t)r, 0x5EED0000ull + (uint64_t)(j0 >> 2), c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = j0 + k;
            if (j < dim) {
                const float u = ((float)(c[k] >> 9) + 0.5f) * 0x1.0p-23f;
                out[r * ld + j] = fmaf(u, hi[j] - lo[j], lo[j]);
            }
        }
    }
}

}  // namespace pfn
