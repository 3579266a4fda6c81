// Ensemble path (n_estimators > 1) of the regressor the reference constructs with default kwargs
// (/root/reference/npe_pfn/npe_pfn.py:48; upstream tabpfn default: 8 members):
//   member_transform_kernel  per-member feature pipeline of the TEST rows (constant-feature removal, quantile-uniform
//                            transform with the original columns appended and truncated-SVD components, or
//                            standardise -> Yeo-Johnson -> standardise; fingerprint feature; feature shuffle).
//                            The statistics (quantile tables, lambdas, SVD basis) are fitted on the context by the
//                            host mirror (npe_pfn_b200/ensemble.py) and arrive here as device tables.
//   ensemble_combine_kernel  per row: softmax of every member's logits, re-binning of members whose target was
//                            transformed onto the common bucket borders (CDF interpolation under the member's
//                            piecewise-uniform density), mean over members, log.
// Both are HBM-bound streaming kernels: 4 F bytes in + 4 F' bytes out per row, and (E + 1) * 4 B bytes per row.
// Specification: oracle/ensemble.py (sklearn's QuantileTransformer / PowerTransformer are the oracle's arithmetic).
#pragma once
#include "common.cuh"

namespace pfn {

constexpr int kMaxBase = 160;  // columns of a member's feature matrix before the shuffle

struct MemberXformArgs {
    const float* X; int64_t ldx; int64_t M; int F_in;
    float* out; int64_t ld_out;
    int n_keep; const int32_t* keep;      // kept (non-constant) raw columns
    int kind;                             // 0 = quantile-uniform + original (+ SVD), 1 = safepower, 2 = kept columns as they are
    int nq; const float* quantiles;       // [n_keep][nq] ascending
    const float* sp;                      // [5][n_keep]: in_mean | in_inv_std | lambda | out_mean | out_inv_std
    int svd_k; const float* svd_inv_scale; const float* svd_vt;  // [2 n_keep], [svd_k][2 n_keep]
    int fingerprint;
    int n_out; const int32_t* perm;       // out column c = base column perm[c]
};

__device__ __forceinline__ float yeo_johnson_f(float x, float lam) {
    if (x >= 0.f) {
        if (fabsf(lam) < 1e-12f) return log1pf(x);
        return (powf(x + 1.0f, lam) - 1.0f) / lam;
    }
    if (fabsf(lam - 2.0f) < 1e-12f) return -log1pf(-x);
    return -(powf(1.0f - x, 2.0f - lam) - 1.0f) / (2.0f - lam);
}

// sklearn QuantileTransformer._transform_col (uniform output): bounds with threshold 1e-7, otherwise the mean of the
// forward and the mirrored np.interp (they differ only on repeated quantile values)
__device__ __forceinline__ float quantile_uniform(float x, const float* q, int nq) {
    if (isnan(x)) return x;
    const double xd = (double)x;
    if (xd - 1e-7 < (double)q[0]) return 0.0f;
    if (xd + 1e-7 > (double)q[nq - 1]) return 1.0f;
    int lo = 0, hi = nq;  // upper bound: first index with q > x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (q[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int j = min(max(lo - 1, 0), nq - 2);  // last index with q[j] <= x
    lo = 0; hi = nq;  // lower bound: first index with q >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (q[mid] < x) lo = mid + 1; else hi = mid;
    }
    const int i = min(max(lo, 1), nq - 1);
    const float step = 1.0f / (float)(nq - 1);
    const float qj = q[j], qj1 = q[j + 1], qi = q[i], qi0 = q[i - 1];
    const float fwd = ((float)j + (qj1 > qj ? (x - qj) / (qj1 - qj) : 0.0f)) * step;
    const float bwd = ((float)i - (qi > qi0 ? (qi - x) / (qi - qi0) : 0.0f)) * step;
    return 0.5f * (fwd + bwd);
}

__device__ __forceinline__ float row_fingerprint(const float* x, int F) {
    unsigned long long h = 0x9E3779B97F4A7C15ull;
    for (int f = 0; f < F; ++f) {
        const float v = x[f];
        unsigned int b = __float_as_uint(v);
        if (v == 0.0f) b = 0u;
        if (isnan(v)) b = 0x7FC00000u;
        h = (h ^ (unsigned long long)b) * 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
    }
    h ^= h >> 29;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 32;
    return (float)(h >> 40) * (1.0f / 16777216.0f);
}

constexpr int XF_WARPS = 8;
__global__ void __launch_bounds__(XF_WARPS * 32) member_transform_kernel(MemberXformArgs a) {
    __shared__ float base_s[XF_WARPS][kMaxBase];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* base = base_s[warp];
    const int nk = a.n_keep;
    const int n_el = a.kind == 0 ? 2 * nk : nk;  // elementwise columns
    for (int64_t r = (int64_t)blockIdx.x * XF_WARPS + warp; r < a.M; r += (int64_t)gridDim.x * XF_WARPS) {
        const float* x = a.X + r * a.ldx;
        __syncwarp();
        for (int c = lane; c < n_el; c += 32) {
            if (a.kind == 0) {
                const int f = c < nk ? c : c - nk;
                const float v = x[a.keep[f]];
                base[c] = c < nk ? quantile_uniform(v, a.quantiles + (size_t)f * a.nq, a.nq) : v;
            } else if (a.kind == 2) {
                base[c] = x[a.keep[c]];
            } else {
                const float z = (x[a.keep[c]] - a.sp[c]) * a.sp[nk + c];
                base[c] = (yeo_johnson_f(z, a.sp[2 * nk + c]) - a.sp[3 * nk + c]) * a.sp[4 * nk + c];
            }
        }
        __syncwarp();
        int n_base = n_el;
        if (a.kind == 0 && a.svd_k > 0) {
            for (int c = lane; c < a.svd_k; c += 32) {
                const float* v = a.svd_vt + (size_t)c * n_el;
                float acc = 0.f;
                for (int j = 0; j < n_el; ++j) acc = fmaf(base[j] * a.svd_inv_scale[j], v[j], acc);
                base[n_el + c] = acc;
            }
            n_base += a.svd_k;
        }
        if (a.fingerprint) {
            if (lane == 0) base[n_base] = row_fingerprint(x, a.F_in);
            n_base += 1;
        }
        __syncwarp();
        float* o = a.out + r * a.ld_out;
        for (int c = lane; c < a.n_out; c += 32) o[c] = base[a.perm[c]];
    }
}

// ---- combine -------------------------------------------------------------------------------------------------------
struct CombineArgs {
    const float* logits; int64_t ld_logits; int64_t member_stride;  // member e, row r: logits + e*member_stride + r*ld
    int E; int B; int64_t M;
    const int32_t* idx;    // [E][B+1]; idx[e][0] < 0: member e already lives on the common borders
    const float* frac;     // [E][B+1]
    const uint8_t* valid;  // [E][B]
    float* out; int64_t ld_out;
};

constexpr int CB_THREADS = 256;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < CB_THREADS / 32; ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < CB_THREADS / 32; ++i) r += red[i];
    return r;
}

// one CTA per row; shared: acc[B] fp32 | p[B] fp32 | C[B+1] fp64 (exclusive prefix sums in double: the re-binned
// bucket mass is a difference of two CDF values)
__global__ void __launch_bounds__(CB_THREADS) ensemble_combine_kernel(CombineArgs a) {
    extern __shared__ __align__(16) uint8_t cb_smem[];
    __shared__ float red_f[CB_THREADS / 32];
    __shared__ double red_d[CB_THREADS / 32];
    __shared__ double part[CB_THREADS];
    const int B = a.B, tid = threadIdx.x;
    double* C = reinterpret_cast<double*>(cb_smem);
    float* acc = reinterpret_cast<float*>(cb_smem + (size_t)(B + 2) * 8);
    float* p = acc + B;
    const int seg = (B + CB_THREADS - 1) / CB_THREADS;
    for (int64_t r = blockIdx.x; r < a.M; r += gridDim.x) {
        for (int i = tid; i < B; i += CB_THREADS) acc[i] = 0.f;
        for (int e = 0; e < a.E; ++e) {
            const float* lg = a.logits + (int64_t)e * a.member_stride + r * a.ld_logits;
            const int32_t* idx = a.idx + (size_t)e * (B + 1);
            const bool rebin = idx[0] >= 0;
            const uint8_t* valid = a.valid + (size_t)e * B;
            float m = -INFINITY;
            for (int i = tid; i < B; i += CB_THREADS) {
                const float v = (!rebin || valid[i]) ? lg[i] : -INFINITY;
                p[i] = v;
                m = fmaxf(m, v);
            }
            m = block_reduce_max(m, red_f);
            double s = 0.0;
            for (int i = tid; i < B; i += CB_THREADS) {
                const float ev = expf(p[i] - m);
                p[i] = ev;
                s += (double)ev;
            }
            s = block_reduce_sum(s, red_d);
            const float inv = (float)(1.0 / s);
            if (!rebin) {
                for (int i = tid; i < B; i += CB_THREADS) acc[i] += p[i] * inv;
                __syncthreads();
                continue;
            }
            // exclusive prefix sums of the normalised masses: contiguous segment per thread + scan of the partials
            const int s0 = min(tid * seg, B), s1 = min(s0 + seg, B);
            double loc = 0.0;
            for (int i = s0; i < s1; ++i) loc += (double)(p[i] * inv);
            part[tid] = loc;
            __syncthreads();
            if (tid < 32) {  // 256 partials: each lane scans 8, then a warp scan of the lane totals
                double t8[8], run = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { t8[q] = run; run += part[tid * 8 + q]; }
                double incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                const double excl = incl - run;
#pragma unroll
                for (int q = 0; q < 8; ++q) part[tid * 8 + q] = excl + t8[q];
            }
            __syncthreads();
            double run = part[tid];
            for (int i = s0; i < s1; ++i) { C[i] = run; run += (double)(p[i] * inv); }
            __syncthreads();
            const float* frac = a.frac + (size_t)e * (B + 1);
            for (int k = tid; k < B; k += CB_THREADS) {
                const int j0 = idx[k], j1 = idx[k + 1];
                const double c0 = C[j0] + (double)(p[j0] * inv) * (double)frac[k];
                const double c1 = C[j1] + (double)(p[j1] * inv) * (double)frac[k + 1];
                acc[k] += (float)fmax(c1 - c0, 0.0);
            }
            __syncthreads();
        }
        float* o = a.out + r * a.ld_out;
        const float invE = 1.0f / (float)a.E;
        for (int i = tid; i < B; i += CB_THREADS) o[i] = logf(acc[i] * invE);
        __syncthreads();
    }
}

static inline size_t combine_smem_bytes(int B) { return (size_t)(B + 2) * 8 + (size_t)2 * B * 4; }

}  // namespace pfn
