/* npe_pfn_b200 — C ABI of the B200-native (sm_100a) in-context engine.
 *
 * This is the boundary the reference's hot path binds to.  The reference
 * (`/root/reference/npe_pfn/npe_pfn.py`) reaches its arithmetic through five
 * calls into the third-party `tabpfn` estimator:
 *
 *   TabPFNRegressor(**kw)                              npe_pfn.py:48, 69     -> pfn_ctx_create
 *   model.fit(joint[:, :dx+d], joint[:, dx+d])         npe_pfn.py:140,215,502 -> pfn_prefill
 *   model.predict(X, output_type="full", quantiles=[]) npe_pfn.py:143-145,217-219,505-507
 *                                                      -> pfn_forward_logits
 *   criterion.sample(logits)                           npe_pfn.py:146, 220   -> pfn_head_sample / pfn_sample (fused)
 *   criterion(logits, y)                               npe_pfn.py:149-151,226-228,510-512
 *                                                      -> pfn_head_nll / pfn_logprob (fused)
 *   accept_reject_sample / _within_support             accept_reject_sampler.py:54-62, npe_pfn.py:581-600
 *                                                      -> pfn_accept_compact
 *
 * Conventions
 *   - plain C types only; every data pointer is a DEVICE pointer owned by the
 *     caller (e.g. a torch tensor's data_ptr) unless marked "host";
 *   - the library owns the handle, its bf16 weight copies, per-slot K/V caches
 *     and a grow-only workspace;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as
 *     void*); no host synchronisation except where a count is returned through
 *     a host pointer and in workspace growth;
 *   - every function returns 0 on success, non-zero on failure;
 *     `pfn_last_error()` returns a thread-local message.
 *   - a "slot" holds everything `fit` produces for one (context, dimension):
 *     encoder statistics, y mean/std, renormalised bucket borders and the
 *     head-0 K/V cache of the context rows for all layers.
 */
#ifndef NPE_PFN_B200_H
#define NPE_PFN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFN_ABI_VERSION 2

typedef struct pfn_ctx pfn_ctx; /* opaque */

typedef struct pfn_model_config {
    int32_t emsize;       /* 192 */
    int32_t nhead;        /* 6   */
    int32_t nlayers;      /* 12  */
    int32_t nhid;         /* 768 */
    int32_t num_buckets;  /* 5000 */
    int32_t max_groups;   /* rows of pos_emb in the blob */
    int32_t max_slots;    /* number of (context, dimension) caches kept */
    int32_t chunk_rows;   /* test rows processed per pass (0 = default) */
    float ln_eps;         /* 1e-5 */
    float softmax_temperature; /* logits are divided by this (0.9) */
} pfn_model_config;

int pfn_abi_version(void);
const char* pfn_last_error(void);

/* `weights` = flat fp32 blob on the device in the order of
 * npe_pfn_b200/weights.py::blob_layout; copied/converted, caller may free it. */
int pfn_ctx_create(const pfn_model_config* cfg, const float* weights, size_t n_floats, int device,
                   void* stream, pfn_ctx** out);
int pfn_ctx_destroy(pfn_ctx* ctx);
/* runtime switches (no reference counterpart): "attn_impl" / "gemm_impl" 0 = warp-level mma.sync kernels,
 * 1 = tcgen05/TMEM/TMA kernels (default); "attn_impl" 2 = attn_tc6 for test rows (row sum on the tensor core, two query
 * tiles per CTA; measured slower than 1, kept as a parity-tested opt-in); "mlp_fused" 1 (default) = MLP sub-layer in one kernel (hidden activation stays on chip), 0 = two GEMM launches; "chunk_rows" = test rows per pass; "standardize_y" 0 = targets enter the y-encoder unscaled (classifier head:
 * class indices, npe_pfn.py:610, 661); "attn_poly" k = k of every 16 pairs of softmax exponentials on the FMA pipes; "attn_lean" 2 (default) / 1 = reference maximum folded
 * into the Q K^T MMA + overflow check instead of the maximum pass (1: every reference change redoes / slows the warp's tile, 2: kept on the
 * fast path), 0 = explicit maximum pass per tile; "head_impl" 2 (default) = head_row2_kernel (persistent CTAs, next logits row
 * prefetched by a bulk copy, 15-instruction bucket mass), 1 = the same arithmetic without prefetch, 0 = round-1 warp-per-row
 * kernel (all three bit-identical); "head_threads" 128 (default) / 64 / 256 = threads per logits row of head_row2_kernel;
 * "dec_rows" = rows per decoder + head pass (default 16 384; logits workspace = dec_rows x num_buckets fp32);
 * "time_kernels" 1 = record a
 * CUDA event pair around every attention / GEMM launch on its stream (read back with pfn_kernel_times). */
int pfn_set_option(pfn_ctx* ctx, const char* key, int64_t value);

/* fit: context rows X[N, F] (row stride ldx floats), raw targets y[N].
 * Standardises y, computes encoder statistics, runs the context through all
 * layers and stores the head-0 K/V cache in `slot`. */
int pfn_prefill(pfn_ctx* ctx, int slot, const float* X, int64_t ldx, const float* y, int64_t N, int F,
                void* stream);

/* predict(output_type="full"): raw decoder logits / temperature for M test rows.
 * X[M, F] with row stride ldx; out[M, num_buckets] with row stride ld_out. */
int pfn_forward_logits(pfn_ctx* ctx, int slot, const float* X, int64_t ldx, int64_t M, float* out,
                       int64_t ld_out, void* stream);

/* Head on materialised logits (what `criterion.sample` / `criterion(logits, y)` do).
 * uniforms == NULL -> Philox4x32-10(seed; counter = (row0 + r, offset)).
 * Any of out_bin / out_u / out_logp may be NULL.  out_theta[r * ld_theta].
 * out_logp (if given) receives log p(theta_r) with -inf replaced by log(eps);
 * when `accumulate` != 0 it is added to the existing value.
 * Row r reads logits row r / group (group >= 1; ld_logits == 0: one row for all): `sample_batched` draws `group`
 * samples per observation from that observation's dimension-0 logits (npe_pfn.py:199, 211-220) in ONE launch. */
int pfn_head_sample(pfn_ctx* ctx, int slot, const float* logits, int64_t ld_logits, int64_t group, int64_t M,
                    const float* uniforms, uint64_t seed, uint64_t row0, uint64_t offset,
                    float* out_theta, int64_t ld_theta, int32_t* out_bin, float* out_u,
                    float* out_logp, float eps, int accumulate, void* stream);
/* out_nll[r] = -log p(y_r) (may be +inf).  If out_logp != NULL it receives the
 * clamped log-prob as in pfn_head_sample. */
int pfn_head_nll(pfn_ctx* ctx, int slot, const float* logits, int64_t ld_logits, int64_t group, int64_t M,
                 const float* y, int64_t ld_y, float* out_nll, float* out_logp, float eps,
                 int accumulate, void* stream);

/* Fused autoregressive step: forward + head, logits never leave the library.
 * pfn_sample writes theta_d for each row (e.g. straight into column dx+d of
 * the caller's joint matrix: out_theta = joint + dx + d, ld_theta = ld of joint). */
int pfn_sample(pfn_ctx* ctx, int slot, const float* X, int64_t ldx, int64_t M, const float* uniforms,
               uint64_t seed, uint64_t row0, uint64_t offset, float* out_theta, int64_t ld_theta,
               int32_t* out_bin, float* out_logp, float eps, int accumulate, void* stream);
int pfn_logprob(pfn_ctx* ctx, int slot, const float* X, int64_t ldx, int64_t M, const float* y,
                int64_t ld_y, float* out_logp, float eps, int accumulate, void* stream);

/* Support check + ordered stream compaction (accept_reject_sampler.py:54-62).
 * Row r is accepted iff lo[j] <= theta[r, j] <= hi[j] for all j (lo/hi may be
 * +-inf; NULL = unbounded) and (mask == NULL or mask[r] != 0) and all finite.
 * out_idx[0..count) = accepted row indices in increasing order; out_rows (may
 * be NULL) receives the accepted rows packed [count, dim] (row stride dim);
 * out_count is a DEVICE int64.  One kernel, predicate evaluated once, no host sync. */
int pfn_accept_compact(pfn_ctx* ctx, const float* theta, int64_t ld, int64_t M, int dim, const float* lo,
                       const float* hi, const uint8_t* mask, int64_t* out_idx, float* out_rows,
                       int64_t* out_count, void* stream);

/* The same predicate + ordered compaction as ONE link of an on-device rejection loop: accepted rows are APPENDED to
 * out_rows[capacity, dim] (row stride out_ld) at the device-resident cursor[0], in proposal order; cursor[0] += number
 * accepted, cursor[1] += M (proposals seen).  Rows past `capacity` are counted but not written, so "the first
 * num_samples accepted, in proposal order" (accept_reject_sampler.py:82) is what the buffer holds however many rounds
 * are enqueued.  Optional extras: `score`/`thr` accept only rows with score[r] > *thr (thr is a DEVICE scalar: the TSNPE
 * truncated-prior rule `log_probs > self.thr`, support_posterior.py:152-154); `logp` is a per-row payload carried into
 * out_logp.  One kernel (warp ballot + block scan + decoupled look-back), no host sync. */
int pfn_accept_append(pfn_ctx* ctx, const float* theta, int64_t ld, int64_t M, int dim, const float* lo, const float* hi,
                      const uint8_t* mask, const float* score, const float* thr, const float* logp, float* out_rows,
                      int64_t out_ld, float* out_logp, int64_t capacity, int64_t* cursor, void* stream);

/* Prior proposals on a box (`prior.sample((bs,))` for a BoxUniform, support_posterior.py:137, 305-309) drawn on the
 * device: out[r, j] = lo[j] + u (hi[j] - lo[j]), u = Philox4x32-10(seed; row0 + r, j). */
int pfn_uniform_box(pfn_ctx* ctx, const float* lo, const float* hi, int64_t M, int dim, uint64_t seed, uint64_t row0,
                    float* out, int64_t ld, void* stream);

/* The whole accept/reject loop of `NPE_PFN_Core.sample` (npe_pfn.py:284-303 -> accept_reject_sampler.py:48-77) for a box
 * (or unbounded: lo = hi = NULL) prior support, enqueued WITHOUT any host round trip: `n_rounds` proposal rounds of
 * `round_rows` rows; each round = autoregressive draw of all `dtheta` dimensions for the observation x_obs[dx]
 * (`slots` is a HOST array: slots[d] holds the prefilled context of dimension d, F = dx + d), support check, ordered
 * append at the device cursor (see pfn_accept_append).  The caller reads cursor[0] once when it needs the count.
 * Philox rows of round k: row0 + k * round_rows + r.  out_logp (optional) receives the summed per-dimension log-prob
 * (-inf -> log eps per dimension) of the accepted draws. */
int pfn_sample_rejection(pfn_ctx* ctx, const int32_t* slots, const float* x_obs, int dx, int dtheta, int n_rounds,
                         int64_t round_rows, const float* lo, const float* hi, uint64_t seed, uint64_t row0, float eps,
                         float* out_theta, int64_t out_ld, float* out_logp, int64_t capacity, int64_t* cursor,
                         void* stream);

/* Context filter (support_posterior.py:357-369, called from get_context, npe_pfn.py:739-744): z-score the columns
 * of x_train[Ntot, dx] (row stride ld), L2 distance of every row to obs[dx], indices of the k nearest rows in
 * ascending distance (ties by index) -> out_idx[k]; out_dist[k] optional.  k <= 16384.  No host sync. */
int pfn_filter_context(pfn_ctx* ctx, const float* x_train, int64_t ld, int64_t Ntot, int dx, const float* obs,
                       int64_t k, int64_t* out_idx, float* out_dist, void* stream);

/* introspection (host values) */
int pfn_slot_info(pfn_ctx* ctx, int slot, int64_t* N, int32_t* F, int32_t* T, int64_t* kv_bytes);
/* number of kernels this library launched since creation (bench.py's gpu_launches) */
int64_t pfn_launch_count(pfn_ctx* ctx);
/* per-class device time of the launches recorded while "time_kernels" was on (host arrays of n_classes >= 9:
 * 0 = item attention of test rows, 1 = item attention of context rows, 2 = projection GEMMs, 3 = other,
 * 4 = fused MLP sub-layer, 5 = bar-distribution head, 6 = cell encoder, 7 = K/V cache writer, 8 = support check +
 * compaction): summed milliseconds, launch counts, algorithmic FLOPs and (HBM-bound classes) algorithmic bytes.
 * Synchronises the device. */
int pfn_kernel_times(pfn_ctx* ctx, double* ms, int64_t* counts, double* flops, double* bytes, int n_classes, int reset);
/* debug / parity: copy a slot's derived quantities to caller device buffers (any may be NULL):
 * stats[3 * 2G] = mean | std | group scale (first G), y_stats[2] = mean, std, borders[num_buckets+1],
 * kv[L][T][N][64] bf16 (K 0..31 | V 32..63). */
int pfn_slot_export(pfn_ctx* ctx, int slot, float* stats, float* y_stats, float* borders, void* kv,
                    void* stream);
/* Slot transfer (prefills of different dimensions are independent, so ranks can split them and exchange the results
 * over NCCL): pfn_slot_state copies the raw encoder-statistics block (452 floats: mean[128] | std[128] | scale[64] |
 * y_mean, y_std, y_fill, 0) to a device buffer; pfn_slot_import installs a slot from that block, its bucket borders
 * [num_buckets + 1] and its K/V cache [L][T][N][64] bf16 (all device pointers), as pfn_slot_export produced them. */
int pfn_slot_state(pfn_ctx* ctx, int slot, float* enc_state, void* stream);
int pfn_slot_import(pfn_ctx* ctx, int slot, int64_t N, int F, const float* enc_state, const float* borders,
                    const void* kv, void* stream);
/* debug: event counters of the tcgen05 item-attention kernel since option "attn_debug" was last set to 1 (host array
 * of 3): fast-path tiles redone after the overflow check fired, reference-maximum changes after a row's first tile,
 * tiles (per warp) that took the general path.  Synchronises the device. */
int pfn_attn_debug_counts(pfn_ctx* ctx, uint64_t* out3);
/* debug / parity: final-layer states of the last forward chunk [rows, T, E] fp32 */
int pfn_debug_last_states(pfn_ctx* ctx, float* out, int64_t max_floats, void* stream);

/* ---- ensemble path (n_estimators > 1): what upstream `TabPFNRegressor()` does by default for the reference's
 * constructor call (npe_pfn.py:48; 8 members).  Each member is an ordinary slot (its own
 * transformed context); these two calls are the per-member feature pipeline of the TEST rows and the combination of
 * the members' logits.  All pointers inside the descriptor are DEVICE pointers owned by the caller. */
typedef struct pfn_member_desc {
    int32_t n_features_in;      /* raw feature count F of the rows handed in                                      */
    int32_t n_keep;             /* non-constant raw features                                                      */
    const int32_t* keep;        /* [n_keep] raw column indices                                                    */
    int32_t kind;               /* 0: quantile-uniform + original columns (+ SVD components); 1: standardise ->   */
                                /*    Yeo-Johnson -> standardise ("safepower"); 2: kept columns as they are       */
    int32_t n_quantiles;        /* kind 0                                                                         */
    const float* quantiles;     /* kind 0: [n_keep][n_quantiles] ascending (references are j / (n_quantiles - 1)) */
    const float* safepower;     /* kind 1: [5][n_keep] = in_mean | 1/in_std | lambda | out_mean | 1/out_std       */
    int32_t svd_k;              /* kind 0: appended components (0 = none)                                         */
    const float* svd_inv_scale; /* [2 n_keep]                                                                     */
    const float* svd_vt;        /* [svd_k][2 n_keep]                                                              */
    int32_t fingerprint;        /* append the row-hash feature                                                    */
    int32_t n_out;              /* columns written = elementwise + svd_k + fingerprint                            */
    const int32_t* perm;        /* [n_out] feature shuffle: out column c = pre-shuffle column perm[c]             */
} pfn_member_desc;
/* X[M, n_features_in] (row stride ldx) -> out[M, n_out] (row stride ld_out). */
int pfn_member_transform(pfn_ctx* ctx, const pfn_member_desc* desc, const float* X, int64_t ldx, int64_t M, float* out,
                         int64_t ld_out, void* stream);
/* out[r, :] = log(mean_e probs_e[r, :]) on the common bucket borders.  Member e's logits for row r start at
 * logits + e * member_stride + r * ld_logits (already divided by the temperature, as pfn_forward_logits returns them).
 * idx[e][0] < 0: member e lives on the common borders (softmax only); otherwise its softmax masses (buckets with
 * valid[e][j] == 0 dropped, renormalised) are re-binned: CDF_e(z_k) = C_e[idx[e][k]] + p_e[idx[e][k]] * frac[e][k] for the
 * B + 1 common borders z_k, C_e the exclusive prefix sum, new mass k = max(CDF_e(z_{k+1}) - CDF_e(z_k), 0).
 * idx [E][B+1] int32, frac [E][B+1], valid [E][B]. */
int pfn_ensemble_combine(pfn_ctx* ctx, const float* logits, int64_t ld_logits, int64_t member_stride, int n_members,
                         int64_t M, const int32_t* idx, const float* frac, const uint8_t* valid, float* out,
                         int64_t ld_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPE_PFN_B200_H */
