// Context filter on the device: `standardized_euclidean_filtering`
// (/root/reference/npe_pfn/support_posterior.py:357-369): z-score every x column over ALL simulations, L2 distance
// of every simulation to the observation, keep the k nearest ordered by distance (ties by index).
//   1. column mean / unbiased std         (one block per column, fp64 accumulation)
//   2. distance scan                      (HBM-bound: 4*dx bytes per simulation)
//   3. exact k-th smallest by 4-pass MSB radix select on the fp32 bit patterns (non-negative floats order as uints)
//   4. ordered compaction of {d < v_k} plus the first k - count ties {d == v_k}
//   5. one-CTA bitonic sort of the k (distance, index) pairs in shared memory (k <= 16384)
#pragma once
#include "common.cuh"
#include "small_kernels.cuh"

namespace pfn {

constexpr int FLT_MAX_K = 16384;

__global__ void __launch_bounds__(256) filter_stats_kernel(const float* __restrict__ X, int64_t ld, int64_t N, int dx,
                                                           float* __restrict__ stats /* mean[dx] | std[dx] */) {
    __shared__ double sh[8];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s += (double)X[i * ld + c];
    s = block_sum_double(s, sh);
    const double mean = s / (double)N;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double d = (double)X[i * ld + c] - mean;
        q += d * d;
    }
    q = block_sum_double(q, sh);
    if (threadIdx.x == 0) {
        stats[c] = (float)mean;
        stats[dx + c] = (float)(N > 1 ? sqrt(q / (double)(N - 1)) : 0.0);
    }
}

// d_i = || (x_i - mean)/std - (obs - mean)/std ||_2 ; zero-variance columns are skipped (the reference yields NaN there)
__global__ void filter_dist_kernel(const float* __restrict__ X, int64_t ld, int64_t N, int dx,
                                   const float* __restrict__ obs, const float* __restrict__ stats,
                                   float* __restrict__ dist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float acc = 0.f;
    for (int c = 0; c < dx; ++c) {
        const float mu = stats[c], sd = stats[dx + c];
        if (sd > 0.f) {
            const float a = (X[i * ld + c] - mu) / sd, b = (obs[c] - mu) / sd;
            const float d = a - b;
            acc = fmaf(d, d, acc);
        }
    }
    dist[i] = sqrtf(acc);
}

// state[0] = prefix (bits fixed so far), state[1] = remaining rank (0-based) inside the prefix bucket
__global__ void __launch_bounds__(256) filter_hist_kernel(const float* __restrict__ dist, int64_t N, int shift,
                                                          const unsigned long long* __restrict__ state,
                                                          unsigned int* __restrict__ hist /* 256 bins, zeroed */) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = (uint32_t)state[0];
    const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = __float_as_uint(dist[i]);
        if ((b & himask) == (prefix & himask)) atomicAdd(&sh[(b >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void filter_pick_kernel(unsigned int* __restrict__ hist, int shift, unsigned long long* __restrict__ state) {
    if (threadIdx.x != 0) return;
    unsigned long long rank = state[1];
    uint32_t prefix = (uint32_t)state[0];
    unsigned long long below = state[2];  // elements strictly below the current prefix bucket
    for (int d = 0; d < 256; ++d) {
        const unsigned long long c = hist[d];
        if (rank < c) { prefix |= (uint32_t)d << shift; break; }
        rank -= c;
        below += c;
    }
    for (int d = 0; d < 256; ++d) hist[d] = 0;
    state[0] = prefix;
    state[1] = rank;
    state[2] = below;
}

// mask: 1 = strictly below the k-th value, 2 = equal to it
__global__ void filter_mask_kernel(const float* __restrict__ dist, int64_t N, const unsigned long long* __restrict__ state,
                                   uint8_t* __restrict__ mask_lt, uint8_t* __restrict__ mask_eq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t vk = (uint32_t)state[0];
    const uint32_t b = __float_as_uint(dist[i]);
    mask_lt[i] = b < vk;
    mask_eq[i] = b == vk;
}

// candidates = idx_lt[0 .. n_lt) ++ idx_eq[0 .. k - n_lt); bitonic sort of (distance bits, index) in shared memory
__global__ void __launch_bounds__(1024) filter_sort_kernel(const float* __restrict__ dist, const int64_t* __restrict__ idx_lt,
                                                           const int64_t* __restrict__ n_lt_p,
                                                           const int64_t* __restrict__ idx_eq, int64_t k, int npow2,
                                                           int64_t* __restrict__ out_idx, float* __restrict__ out_dist) {
    extern __shared__ unsigned long long keys[];
    const int64_t n_lt = *n_lt_p;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < k) {
            const int64_t id = i < n_lt ? idx_lt[i] : idx_eq[i - n_lt];
            key = ((unsigned long long)__float_as_uint(dist[id]) << 32) | (unsigned long long)(uint32_t)id;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < npow2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_idx[i] = (int64_t)(uint32_t)keys[i];
        if (out_dist) out_dist[i] = __uint_as_float((uint32_t)(keys[i] >> 32));
    }
}

}  // namespace pfn
