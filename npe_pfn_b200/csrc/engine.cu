// npe_pfn_b200 engine: C-ABI (include/npe_pfn_b200.h) over the sm_100a kernels.
//
// Replaces, behind the reference's five-call estimator protocol, what the third-party
// `tabpfn` package does for /root/reference/npe_pfn/npe_pfn.py:140 (fit), :143-145 (predict), :146 (sample)
// and :149-151 (log density):
//   pfn_prefill         context rows -> encoder statistics + per-layer head-0 K/V cache kept in HBM
//   pfn_forward_logits  test rows attend to that cache only -> bar-distribution logits
//   pfn_sample/logprob  the same forward fused with the head; logits stay inside the library
//   pfn_accept_compact  support check + ordered compaction of accepted draws
#include "../../include/npe_pfn_b200.h"

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "attn_mma.cuh"
#ifdef PFN_WITH_ATTN_TC
#include "attn_tc.cuh"
#include "attn_tc6.cuh"
#endif
#include "common.cuh"
#include "gemm_mma.cuh"
#ifdef PFN_WITH_ATTN_TC
#include "gemm_tc.cuh"
#include "mlp_tc.cuh"
#endif
#include "small_kernels.cuh"
#include "head_kernels.cuh"
#include "filter_kernels.cuh"
#include "ensemble_kernels.cuh"

namespace pfn {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
}  // namespace pfn

using namespace pfn;

namespace {

struct Slot {
    bool valid = false;
    int64_t N = 0;
    int F = 0, G = 0, T = 0;
    float* enc = nullptr;      // kEncFloats
    float* borders = nullptr;  // num_buckets + 1, original units
    bf16* kv = nullptr;        // [L][T][N][64]
    size_t kv_cap = 0;         // bytes
};

struct Offsets {  // element offsets into the weight blob (npe_pfn_b200/weights.py::blob_layout)
    size_t enc_x_w, enc_y_w, enc_y_b, pos_emb, feat_wqkv, feat_wo, item_wqkv, item_wo, mlp_w1, mlp_w2, dec_w1, dec_b1,
        dec_w2, dec_b2, borders, total;
};

}  // namespace

struct pfn_ctx {
    pfn_model_config cfg;
    int device = 0;
    Offsets off;
    float* wf = nullptr;  // fp32 blob
    bf16* wb = nullptr;   // bf16 copy of the blob
    bf16* wb_fqkv = nullptr;  // [L][576][192] feature-attention [Wq; Wk; Wv] in head-pair-major order (gemm_tc.cuh EPI_FEATURE_ATTN)
    int feat_fused = 1;       // 1: QKV projection + attention between features in one kernel (T <= 16, tcgen05 GEMMs)
    std::vector<Slot> slots;
    // workspace (token capacity `cap_tok`)
    int64_t cap_tok = 0;
    uint8_t* ws = nullptr;
    float* xf = nullptr;
    bf16 *xb = nullptr, *qkv = nullptr, *ob = nullptr;
    // sub-chunk scratch (reused so that it stays L2 resident): feature-attention qkv and MLP hidden
    bf16 *qkv_s = nullptr, *hb_s = nullptr;
    int64_t sub_tok = 0;      // scratch capacity in tokens
    int64_t sub_tok_opt = 0;  // option "sub_tokens": 0 = the whole chunk in one go
    // decoder / head workspace
    int dec_rows = 16384;  // rows per decoder + head pass: 328 MB of fp32 logits; the persistent head kernel needs >= 20 rows per CTA to amortise its ramp (4096: 2.4 TB/s, 16384: 3.0 TB/s)
    bf16* dech = nullptr;
    float* logits = nullptr;
    // compaction scratch
    // head: integer CDFs of logits rows that many draws share (head_kernels.cuh)
    unsigned long long* hd_cdf = nullptr;
    float* hd_max = nullptr;
    double* hd_logZ = nullptr;
    int64_t hd_rows = 0;
    int head_impl = 2;  // 0 = round-1 warp-per-row kernel (shared-memory staging), 1 = register-resident rows + shared-row CDFs,
                        // 2 = 1 with the next row prefetched by a bulk copy and the 15-instruction bucket mass (head_row2_kernel)
    int head_threads = 128;  // threads per logits row of head_row2_kernel (256: 4 CTAs per SM, 2.97 TB/s; 128: 5 CTAs per SM, 3.31 TB/s; 64: 6 CTAs per SM)
    unsigned long long* cp_state = nullptr;  // [cp_cap] look-back tile states | 2 ticket words (zeroed per launch)
    int64_t cp_cap = 0;
    // on-device rejection loop (pfn_sample_rejection): joint test matrix / log-probs of one proposal round
    float* rj_buf = nullptr;
    float* rj_logp = nullptr;
    int64_t rj_rows = 0, rj_ld = 0;
    // context-filter scratch
    uint8_t* flt_ws = nullptr;
    int64_t flt_cap = 0;
    int64_t launches = 0;
    int64_t last_rows = 0;
    int last_T = 0;
    int attn_impl = 1;  // 0 = mma.sync, 1 = tcgen05
    int gemm_impl = 1;  // 0 = mma.sync, 1 = tcgen05
    int attn_persist = 0;       // 1: persistent item-attention CTAs (3 per SM)
    int attn_wait_ticks = 1000;
    int attn_stagger_ns = 0;
    int standardize_y = 1;  // 0 for the classifier head (targets are class indices)
    int mlp_fused = 1;  // 1: MLP sub-layer (up-projection, GELU, down-projection, residual, LayerNorm) in one kernel (mlp_tc.cuh)
    int attn_debug = 0;                      // count reference-change events of the item-attention kernel
    unsigned long long* attn_dbg = nullptr;  // [3] device counters (attn_tc.cuh::TcArgs::dbg)
    int attn_lean = 2;  // 1/2: reference maximum folded into the QK^T MMA, overflow check instead of the maximum pass (attn_tc v5);
                        // 2: rows lagging behind a reference they published adopt it on the fast path, large-but-finite P tiles are kept
    int attn_poly = 5;  // k of every 16 pairs of exponentials on the FMA pipes (+100: degree-2 polynomial); 0 = all on MUFU. r1 sweep: 0 -> 437, 5 -> 473 TFLOP/s
    int num_sms = 148;
    // optional per-class kernel timing (bench.py roofline): CUDA events around each launch on its stream
    int time_kernels = 0;
    struct Timed { cudaEvent_t a, b; int cls; double flops, bytes; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> ev_pool;
};

// classes 5..7 are the HBM-bound kernels north_star wants reported against the measured copy bandwidth
enum KernelClass { KC_ATTN_TEST = 0, KC_ATTN_CTX = 1, KC_GEMM = 2, KC_OTHER = 3, KC_MLP = 4, KC_HEAD = 5, KC_ENCODE = 6, KC_KVCACHE = 7,
                   KC_COMPACT = 8, KC_COUNT = 9 };

namespace {

Offsets make_offsets(const pfn_model_config& c) {
    Offsets o{};
    size_t p = 0;
    const size_t E = c.emsize, H = c.nhid, B = c.num_buckets, L = c.nlayers;
    auto take = [&](size_t n) { size_t r = p; p += n; return r; };
    o.enc_x_w = take(E * 4);
    o.enc_y_w = take(E * 2);
    o.enc_y_b = take(E);
    o.pos_emb = take((size_t)c.max_groups * E);
    o.feat_wqkv = take(L * 3 * E * E);
    o.feat_wo = take(L * E * E);
    o.item_wqkv = take(L * 3 * E * E);
    o.item_wo = take(L * E * E);
    o.mlp_w1 = take(L * H * E);
    o.mlp_w2 = take(L * E * H);
    o.dec_w1 = take(H * E);
    o.dec_b1 = take(H);
    o.dec_w2 = take(B * H);
    o.dec_b2 = take(B);
    o.borders = take(B + 1);
    o.total = p;
    return o;
}

#define PFN_LAUNCH_OK(ctx)                                 \
    do {                                                   \
        (ctx)->launches++;                                 \
        PFN_CUDA_OK(cudaGetLastError());                   \
    } while (0)

int ensure_workspace(pfn_ctx* c, int64_t tokens, cudaStream_t st) {
    if (tokens <= c->cap_tok) return 0;
    PFN_CUDA_OK(cudaStreamSynchronize(st));
    if (c->ws) PFN_CUDA_OK(cudaFree(c->ws));
    c->ws = nullptr;
    c->cap_tok = 0;
    const int64_t cap = tokens + 1024;
    c->sub_tok = cap;
    const size_t per_tok = (size_t)kE * 4 + kE * 2 + 3 * kE * 2 + kE * 2;
    const size_t scratch = (size_t)(c->sub_tok + 128) * (3 * kE + kHid) * 2;
    PFN_CUDA_OK(cudaMalloc(&c->ws, per_tok * (size_t)cap + scratch + 4096));
    uint8_t* p = c->ws;
    c->xf = reinterpret_cast<float*>(p); p += (size_t)cap * kE * 4;
    c->xb = reinterpret_cast<bf16*>(p);  p += (size_t)cap * kE * 2;
    c->qkv = reinterpret_cast<bf16*>(p); p += (size_t)cap * 3 * kE * 2;
    c->ob = reinterpret_cast<bf16*>(p);  p += (size_t)cap * kE * 2;
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
    c->qkv_s = reinterpret_cast<bf16*>(p); p += (size_t)(c->sub_tok + 128) * 3 * kE * 2;
    c->hb_s = reinterpret_cast<bf16*>(p);
    c->cap_tok = cap;
    return 0;
}

int ensure_decoder_ws(pfn_ctx* c) {
    if (c->dech) return 0;
    PFN_CUDA_OK(cudaMalloc(&c->dech, (size_t)c->dec_rows * kHid * 2));
    PFN_CUDA_OK(cudaMalloc(&c->logits, (size_t)c->dec_rows * c->cfg.num_buckets * 4));
    return 0;
}

cudaEvent_t take_event(pfn_ctx* c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct TimeScope {  // records an event pair around the launches issued inside its lifetime
    pfn_ctx* c; cudaStream_t st; pfn_ctx::Timed t; bool on;
    TimeScope(pfn_ctx* c_, cudaStream_t st_, int cls, double flops, double bytes = 0.0) : c(c_), st(st_), on(c_->time_kernels != 0) {
        if (on) { t.a = take_event(c); t.b = take_event(c); t.cls = cls; t.flops = flops; t.bytes = bytes; cudaEventRecord(t.a, st); }
    }
    ~TimeScope() { if (on) { cudaEventRecord(t.b, st); c->timed.push_back(t); } }
};

template <int EPI>
int gemm(pfn_ctx* c, const GemmArgs& a, cudaStream_t st) {
    TimeScope ts(c, st, KC_GEMM, 2.0 * (double)a.M * a.N * a.K);
#ifdef PFN_WITH_ATTN_TC
    if (c->gemm_impl == 1) {
        PFN_CUDA_OK(launch_gemm_tc<EPI>(a, c->num_sms, st));
    } else
#endif
    PFN_CUDA_OK(launch_gemm_mma<EPI>(a, st));
    c->launches++;
    return 0;
}

int item_attention(pfn_ctx* c, const AttnArgs& a, int T, cudaStream_t st) {
    TimeScope ts(c, st, a.k_head ? KC_ATTN_CTX : KC_ATTN_TEST, 4.0 * (double)a.R * kHeads * T * (double)a.N * kDh);
#ifdef PFN_WITH_ATTN_TC
    if (c->attn_impl == 2 && a.k_head == 0 && !c->attn_persist) {
        // v6 (attn_tc6.cuh): test rows against the cached K/V, row sum on the tensor core, two query tiles per CTA
        PFN_CUDA_OK(launch_attn_tc6(a, kHeads, T, c->attn_poly % 100, (uint32_t)c->attn_wait_ticks, c->attn_debug ? c->attn_dbg : nullptr, st));
    } else if (c->attn_impl >= 1) {
        PFN_CUDA_OK(launch_attn_tc(a, kHeads, T, c->attn_poly + 1000 * (c->attn_persist ? 0 : c->attn_lean), c->num_sms, c->attn_persist, (uint32_t)c->attn_wait_ticks, (uint32_t)c->attn_stagger_ns,
                                       c->attn_debug ? c->attn_dbg : nullptr, st));
    } else
#endif
    {
        PFN_CUDA_OK(launch_attn_mma(a, kHeads, T, st));
    }
    c->launches++;
    return 0;
}

// rows through the encoder and all layers; final states in c->xf / c->xb.
// context (y != null): self-attention between items + cache fill; test rows: attend to the slot cache.
int forward_rows(pfn_ctx* c, Slot& s, const float* X, int64_t ldx, const float* y, int64_t R, cudaStream_t st) {
    const int T = s.T, G = s.G;
    const int64_t tok = R * T;
    const bool ctx_rows = y != nullptr;
    const int L = c->cfg.nlayers;
    if (int rc = ensure_workspace(c, tok, st)) return rc;
    const float* wf = c->wf;
    const bf16* wb = c->wb;
    const Offsets& o = c->off;

    {
        // algorithmic bytes: F raw features (+ y) in, T tokens of 192 (fp32 + bf16 copy) out per row
        TimeScope ts(c, st, KC_ENCODE, 0.0, (double)R * (4.0 * (s.F + (ctx_rows ? 1 : 0)) + 6.0 * T * kE));
        encode_kernel<<<(unsigned)ceil_div(R, ENC_ROWS), ENC_ROWS * ENC_TPR, 0, st>>>(X, ldx, s.F, G, y, R, s.enc, wf + o.enc_x_w,
                                                                     wf + o.enc_y_w, wf + o.enc_y_b, wf + o.pos_emb, c->xf,
                                                                     c->xb);
        PFN_LAUNCH_OK(c);
    }

    const size_t fa_smem = (size_t)FA_WARPS * 2 * T * 33 * sizeof(float);
    if (fa_smem > 48 * 1024)
        PFN_CUDA_OK(cudaFuncSetAttribute(feature_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa_smem));

    // The chain between two item-attention kernels runs sub-chunk by sub-chunk (one wave of 128-token tiles), so
    // its intermediates (feature-attention qkv, MLP hidden) live in small reused scratch buffers that stay in L2;
    // only the residual stream, the item queries and the attention output cross HBM once per layer.
    const int64_t sub_rows = c->sub_tok_opt > 0 ? std::max<int64_t>(1, std::min(c->sub_tok_opt, c->sub_tok) / T) : R;
    // Dead-code elimination in the LAST layer: the decoder reads only the y-token (column T-1) of test rows, and
    // nothing reads the final context states.  So for test rows the last layer's item attention, out-projection
    // and MLP run on the y-token column alone (strided rows: leading dimension T*E), and for context rows the last
    // layer stops after its K/V projection.  Identical results, ~1/12 * (T-1)/T less attention / MLP work.
    const int64_t ycol = (int64_t)(T - 1) * kE;  // element offset of the y-token inside a row's token block
    auto first_half = [&](int l, int64_t r0, int64_t nr) -> int {  // feature attention + item-attention projections
        const bool y_only = (l == L - 1) && !ctx_rows;
        const int64_t t0 = r0 * T, ntok = nr * T;
        GemmArgs g{};
        g.ln_eps = c->cfg.ln_eps;
#ifdef PFN_WITH_ATTN_TC
        if (c->gemm_impl == 1 && c->feat_fused && T <= 16) {
            // QKV projection + attention between the T tokens of each row in ONE kernel: qkv never reaches HBM
            TimeScope ts(c, st, KC_GEMM, 2.0 * (double)ntok * 3 * kE * kE + 4.0 * (double)nr * T * T * kE);
            PFN_CUDA_OK(launch_qkv_feature_attn_tc(c->xb + t0 * kE, c->wb_fqkv + (size_t)l * 3 * kE * kE, nr, T, c->ob + t0 * kE,
                                                   c->num_sms, st));
            c->launches++;
        } else
#endif
        {
            g.A = c->xb + t0 * kE; g.lda = kE; g.W = wb + o.feat_wqkv + (size_t)l * 3 * kE * kE; g.M = ntok; g.N = 3 * kE; g.K = kE;
            g.Cb = c->qkv_s; g.ldcb = 3 * kE;
            if (int rc = gemm<EPI_BF16>(c, g, st)) return rc;
            TimeScope ts(c, st, KC_OTHER, 4.0 * (double)nr * T * T * kE);
            if (T <= 16) {
                feature_attn_mma_kernel<<<(unsigned)ceil_div(nr * kHeads, FAM_WARPS), FAM_WARPS * 32, 0, st>>>(
                    c->qkv_s, nr, T, c->ob + t0 * kE);
            } else {
                const unsigned blocks = (unsigned)std::min<int64_t>(nr, 148 * 16);
                feature_attn_kernel<<<blocks, FA_WARPS * 32, fa_smem, st>>>(c->qkv_s, nr, T, c->ob + t0 * kE);
            }
            PFN_LAUNCH_OK(c);
        }
        const int64_t off = y_only ? t0 * kE + ycol : t0 * kE;  // y_only: one row per test row, stride T*E
        const int64_t ld = y_only ? (int64_t)T * kE : kE;
        const int64_t rows = y_only ? nr : ntok;
        g.A = c->ob + off; g.lda = ld; g.W = wb + o.feat_wo + (size_t)l * kE * kE; g.M = rows; g.N = kE; g.K = kE;
        g.Cb = c->xb + off; g.ldcb = ld; g.Cf = c->xf + off; g.ldcf = ld;
        if (int rc = gemm<EPI_RESID_LN>(c, g, st)) return rc;
        g.A = c->xb + off; g.lda = ld; g.W = wb + o.item_wqkv + (size_t)l * 3 * kE * kE; g.K = kE; g.Cf = nullptr;
        if (ctx_rows) {
            g.N = 3 * kE; g.Cb = c->qkv + t0 * 3 * kE; g.ldcb = 3 * kE;
            if (int rc = gemm<EPI_BF16>(c, g, st)) return rc;
            bf16* cache_l = s.kv + (size_t)l * T * s.N * kKvRow;
            {
                TimeScope ts(c, st, KC_KVCACHE, 0.0, (double)ntok * 2.0 * kKvRow * sizeof(bf16));  // 128 B read + 128 B written per token
                kv_cache_kernel<<<(unsigned)ceil_div(ntok * 8, 256), 256, 0, st>>>(c->qkv + t0 * 3 * kE, nr, T, r0, s.N, cache_l);
                PFN_LAUNCH_OK(c);
            }
        } else {
            g.N = kE; g.Cb = c->qkv + off; g.ldcb = ld;  // Q rows of the projection only
            if (int rc = gemm<EPI_BF16>(c, g, st)) return rc;
        }
        return 0;
    };
    auto second_half = [&](int l, int64_t r0, int64_t nr) -> int {  // item out-projection + MLP
        const bool y_only = (l == L - 1) && !ctx_rows;
        const int64_t t0 = r0 * T, ntok = nr * T;
        const int64_t off = y_only ? t0 * kE + ycol : t0 * kE;
        const int64_t ld = y_only ? (int64_t)T * kE : kE;
        const int64_t rows = y_only ? nr : ntok;
        GemmArgs g{};
        g.ln_eps = c->cfg.ln_eps;
        g.A = c->ob + off; g.lda = ld; g.W = wb + o.item_wo + (size_t)l * kE * kE; g.M = rows; g.N = kE; g.K = kE;
        g.Cb = c->xb + off; g.ldcb = ld; g.Cf = c->xf + off; g.ldcf = ld;
        if (int rc = gemm<EPI_RESID_LN>(c, g, st)) return rc;
#ifdef PFN_WITH_ATTN_TC
        if (c->gemm_impl == 1 && c->mlp_fused) {
            TimeScope ts(c, st, KC_MLP, 4.0 * (double)rows * kHid * kE);
            PFN_CUDA_OK(launch_mlp_tc(c->xb + off, ld, wb + o.mlp_w1 + (size_t)l * kHid * kE, wb + o.mlp_w2 + (size_t)l * kE * kHid,
                                      c->xb + off, ld, c->xf + off, ld, rows, c->cfg.ln_eps, c->num_sms, st));
            c->launches++;
            return 0;
        }
#endif
        g.A = c->xb + off; g.lda = ld; g.W = wb + o.mlp_w1 + (size_t)l * kHid * kE; g.N = kHid; g.K = kE;
        g.Cb = c->hb_s; g.ldcb = kHid; g.Cf = nullptr; g.bias = nullptr;
        if (int rc = gemm<EPI_BIAS_GELU_BF16>(c, g, st)) return rc;
        g.A = c->hb_s; g.lda = kHid; g.W = wb + o.mlp_w2 + (size_t)l * kE * kHid; g.N = kE; g.K = kHid;
        g.Cb = c->xb + off; g.ldcb = ld; g.Cf = c->xf + off; g.ldcf = ld;
        return gemm<EPI_RESID_LN>(c, g, st);
    };

    for (int64_t r0 = 0; r0 < R; r0 += sub_rows)
        if (int rc = first_half(0, r0, std::min(sub_rows, R - r0))) return rc;
    for (int l = 0; l < L; ++l) {
        const bool last = l == L - 1;
        if (last && ctx_rows) break;  // the context's last-layer K/V are cached; its final states are never read
        AttnArgs a{};
        a.R = R; a.N = s.N;
        a.O = c->ob; a.o_row = (int64_t)T * kE; a.o_tok = kE;
        int Tq = T;  // token columns the queries cover
        if (ctx_rows) {
            a.Q = c->qkv; a.q_row = (int64_t)T * 3 * kE; a.q_tok = 3 * kE;
            a.K = c->qkv + kE; a.k_tok = 3 * kE; a.k_row = (int64_t)T * 3 * kE; a.k_head = kDh; a.v_off = kE;
        } else {
            a.Q = c->qkv; a.q_row = (int64_t)T * kE; a.q_tok = kE;
            a.K = s.kv + (size_t)l * T * s.N * kKvRow; a.k_tok = s.N * kKvRow; a.k_row = kKvRow; a.k_head = 0; a.v_off = kDh;
            if (last) {  // y-token column only
                a.Q += ycol; a.O += ycol; a.K += (int64_t)(T - 1) * s.N * kKvRow;
                Tq = 1;
            }
        }
        if (int rc = item_attention(c, a, Tq, st)) return rc;
        for (int64_t r0 = 0; r0 < R; r0 += sub_rows) {
            const int64_t nr = std::min(sub_rows, R - r0);
            if (int rc = second_half(l, r0, nr)) return rc;
            if (l + 1 < L)
                if (int rc = first_half(l + 1, r0, nr)) return rc;
        }
    }
    c->last_rows = R;
    c->last_T = T;
    return 0;
}

// decoder on the y-token of rows [r0, r0+n) of the last forward chunk -> out[n, B] (row stride ld_out)
int decode_rows(pfn_ctx* c, const Slot& s, int64_t r0, int64_t n, float* out, int64_t ld_out, cudaStream_t st) {
    const Offsets& o = c->off;
    GemmArgs g{};
    g.A = c->xb + (r0 * s.T + (s.T - 1)) * kE; g.lda = (int64_t)s.T * kE;
    g.W = c->wb + o.dec_w1; g.M = n; g.N = kHid; g.K = kE;
    g.Cb = c->dech; g.ldcb = kHid; g.bias = c->wf + o.dec_b1;
    if (int rc = gemm<EPI_BIAS_GELU_BF16>(c, g, st)) return rc;
    g = GemmArgs{};
    g.A = c->dech; g.lda = kHid; g.W = c->wb + o.dec_w2; g.M = n; g.N = c->cfg.num_buckets; g.K = kHid;
    g.Cf = out; g.ldcf = ld_out; g.bias = c->wf + o.dec_b2; g.scale = 1.0f / c->cfg.softmax_temperature;
    return gemm<EPI_BIAS_SCALE_F32>(c, g, st);
}

int launch_head(pfn_ctx* c, const HeadArgs& h, bool sample, cudaStream_t st) {
    // algorithmic bytes: every distinct logits row once (20 000 B) + the row's scalar inputs / outputs
    const int64_t distinct = h.ld_logits == 0 ? 1 : ceil_div(h.M, h.group);
    TimeScope ts(c, st, KC_HEAD, 0.0, (double)distinct * h.B * 4.0 + (double)h.M * 12.0);
    if (c->head_impl >= 1 && h.M >= 2 * distinct && h.M >= 64) {
        // many draws / targets per logits row: CDF once per distinct row, then one thread per draw
        if (distinct > c->hd_rows) {
            PFN_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(c->hd_cdf); cudaFree(c->hd_max); cudaFree(c->hd_logZ);
            c->hd_cdf = nullptr; c->hd_max = nullptr; c->hd_logZ = nullptr; c->hd_rows = 0;
            const int64_t cap = distinct + 64;
            PFN_CUDA_OK(cudaMalloc(&c->hd_cdf, (size_t)cap * h.B * 8));
            PFN_CUDA_OK(cudaMalloc(&c->hd_max, (size_t)cap * 4));
            PFN_CUDA_OK(cudaMalloc(&c->hd_logZ, (size_t)cap * 8));
            c->hd_rows = cap;
        }
        HeadCdf o{c->hd_cdf, c->hd_max, c->hd_logZ};
        head_cdf_kernel<<<(unsigned)distinct, HC_THREADS, 0, st>>>(h.logits, h.ld_logits, h.B, o);
        PFN_LAUNCH_OK(c);
        const unsigned blocks = (unsigned)ceil_div(h.M, 256);
        if (sample) head_shared_kernel<true><<<blocks, 256, 0, st>>>(h, o);
        else head_shared_kernel<false><<<blocks, 256, 0, st>>>(h, o);
        PFN_LAUNCH_OK(c);
        return 0;
    }
    const bool vec_ok = h.B % 4 == 0 && h.B <= HR_MAX_B && h.ld_logits % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(h.logits) & 15) == 0;
    if (c->head_impl == 2 && vec_ok) {  // persistent CTAs, next row prefetched by a bulk copy (4 x 20.6 KB of shared memory per SM)
        if (c->head_threads == 64) {
            const unsigned blocks = (unsigned)std::min<int64_t>(h.M, (int64_t)c->num_sms * 6);
            if (sample) head_row2_kernel<true, 64, 6><<<blocks, 64, 0, st>>>(h);
            else head_row2_kernel<false, 64, 6><<<blocks, 64, 0, st>>>(h);
        } else if (c->head_threads == 128) {
            const unsigned blocks = (unsigned)std::min<int64_t>(h.M, (int64_t)c->num_sms * 5);
            if (sample) head_row2_kernel<true, 128, 5><<<blocks, 128, 0, st>>>(h);
            else head_row2_kernel<false, 128, 5><<<blocks, 128, 0, st>>>(h);
        } else {
            const unsigned blocks = (unsigned)std::min<int64_t>(h.M, (int64_t)c->num_sms * 4);
            if (sample) head_row2_kernel<true, 256, 4><<<blocks, 256, 0, st>>>(h);
            else head_row2_kernel<false, 256, 4><<<blocks, 256, 0, st>>>(h);
        }
        PFN_LAUNCH_OK(c);
        return 0;
    }
    if (c->head_impl >= 1 && vec_ok) {
        const unsigned blocks = (unsigned)std::min<int64_t>(h.M, (int64_t)c->num_sms * 16);
        if (sample) head_row_kernel<true><<<blocks, HR_THREADS, 0, st>>>(h);
        else head_row_kernel<false><<<blocks, HR_THREADS, 0, st>>>(h);
        PFN_LAUNCH_OK(c);
        return 0;
    }
    const size_t smem = (size_t)HEAD_WARPS * h.B * sizeof(float);
    if (smem > 48 * 1024) {  // per device and cheap: set it every time
        if (sample) PFN_CUDA_OK(cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else PFN_CUDA_OK(cudaFuncSetAttribute(head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(h.M, HEAD_WARPS), 148 * 8);
    if (sample) head_kernel<true><<<blocks, HEAD_WARPS * 32, smem, st>>>(h);
    else head_kernel<false><<<blocks, HEAD_WARPS * 32, smem, st>>>(h);
    PFN_LAUNCH_OK(c);
    return 0;
}

int check_slot(pfn_ctx* c, int slot, bool need_valid) {
    PFN_REQUIRE(c != nullptr, "null context");
    PFN_REQUIRE(slot >= 0 && slot < (int)c->slots.size(), "slot out of range");
    if (need_valid) PFN_REQUIRE(c->slots[slot].valid, "slot has not been prefilled (call pfn_prefill first)");
    return 0;
}

// default: two waves of 128-row query tiles per (head, column): 2 * 148 * 128 = 37 888 rows on a B200, which makes the
// item-attention grid (tiles * 6 * T CTAs against 3 * 148 resident) and the GEMM tile counts whole multiples of a wave
int chunk_rows_of(const pfn_ctx* c) { return c->cfg.chunk_rows > 0 ? c->cfg.chunk_rows : 2 * c->num_sms * 128; }

}  // namespace

// =================================================================================================
extern "C" {

int pfn_abi_version(void) { return PFN_ABI_VERSION; }
const char* pfn_last_error(void) { return last_error().c_str(); }

int pfn_ctx_create(const pfn_model_config* cfg, const float* weights, size_t n_floats, int device, void* stream,
                   pfn_ctx** out) {
    PFN_REQUIRE(cfg && weights && out, "null argument");
    PFN_REQUIRE(cfg->emsize == kE && cfg->nhead == kHeads && cfg->nhid == kHid,
                "this build is specialised for emsize 192, 6 heads, hidden 768");
    PFN_REQUIRE(cfg->nlayers >= 1 && cfg->num_buckets >= 2 && cfg->num_buckets % 2 == 0, "bad layer/bucket count");
    PFN_REQUIRE(cfg->max_groups >= 1 && cfg->max_groups * 2 <= kMaxFeat, "max_groups out of range");
    PFN_REQUIRE(cfg->max_slots >= 1, "max_slots must be >= 1");
    PFN_REQUIRE(cfg->softmax_temperature > 0.f, "softmax_temperature must be > 0");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    PFN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    PFN_REQUIRE(prop.major == 10, "npe_pfn_b200 kernels are built for sm_100a (Blackwell B200) only");
    pfn_ctx* c = new pfn_ctx();
    c->num_sms = prop.multiProcessorCount;
    c->cfg = *cfg;
    c->device = device;
    c->off = make_offsets(*cfg);
    if (c->off.total != n_floats) {
        last_error() = "weight blob has " + std::to_string(n_floats) + " floats, layout needs " + std::to_string(c->off.total);
        delete c;
        return 2;
    }
    if (const char* e = getenv("NPE_PFN_B200_ATTN")) c->attn_impl = (strcmp(e, "mma") == 0) ? 0 : (strcmp(e, "tc5") == 0) ? 1 : (strcmp(e, "tc6") == 0) ? 2 : c->attn_impl;
    if (const char* e = getenv("NPE_PFN_B200_GEMM")) c->gemm_impl = (strcmp(e, "mma") == 0) ? 0 : 1;
    if (const char* e = getenv("NPE_PFN_B200_ATTN_LEAN")) c->attn_lean = atoi(e);
    if (const char* e = getenv("NPE_PFN_B200_MLP_FUSED")) c->mlp_fused = atoi(e);
    if (const char* e = getenv("NPE_PFN_B200_ATTN_POLY")) {  // tuning / parity sweeps of the exponential split
        if (int rc = pfn_set_option(c, "attn_poly", atoll(e))) { delete c; return rc; }
    }
    PFN_CUDA_OK(cudaMalloc(&c->wf, n_floats * 4));
    PFN_CUDA_OK(cudaMalloc(&c->wb, n_floats * 2));
    PFN_CUDA_OK(cudaMemcpyAsync(c->wf, weights, n_floats * 4, cudaMemcpyDeviceToDevice, st));
    f32_to_bf16_kernel<<<(unsigned)ceil_div((int64_t)n_floats, 256), 256, 0, st>>>(c->wf, c->wb, (int64_t)n_floats);
    // fold the softmax scale into the item-attention query projection (bf16 copy only; see common.cuh)
    scale_item_q_kernel<<<(unsigned)ceil_div((int64_t)kE * kE * cfg->nlayers, 256), 256, 0, st>>>(
        c->wf, c->wb, (int64_t)c->off.item_wqkv, cfg->nlayers, kItemScaleLog2);
    PFN_LAUNCH_OK(c);
    {
        // head-pair-major copy of every layer's feature-attention projection: chunk hp = rows [q | k | v] of heads 2hp, 2hp + 1
        const size_t per_layer = (size_t)3 * kE * kE;
        PFN_CUDA_OK(cudaMalloc(&c->wb_fqkv, (size_t)cfg->nlayers * per_layer * sizeof(bf16)));
        for (int l = 0; l < cfg->nlayers; ++l)
            for (int hp = 0; hp < kHeads / 2; ++hp)
                for (int part = 0; part < 3; ++part)  // q, k, v
                    PFN_CUDA_OK(cudaMemcpyAsync(c->wb_fqkv + l * per_layer + ((size_t)hp * 3 + part) * 2 * kDh * kE,
                                                c->wb + c->off.feat_wqkv + l * per_layer + ((size_t)part * kE + (size_t)hp * 2 * kDh) * kE,
                                                (size_t)2 * kDh * kE * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    }
    c->slots.resize(cfg->max_slots);
    for (auto& s : c->slots) {
        PFN_CUDA_OK(cudaMalloc(&s.enc, kEncFloats * 4));
        PFN_CUDA_OK(cudaMalloc(&s.borders, (size_t)(cfg->num_buckets + 1) * 4));
    }
    PFN_CUDA_OK(cudaStreamSynchronize(st));
    *out = c;
    return 0;
}

int pfn_ctx_destroy(pfn_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& s : c->slots) {
        cudaFree(s.enc);
        cudaFree(s.borders);
        cudaFree(s.kv);
    }
    cudaFree(c->wf);
    cudaFree(c->wb);
    cudaFree(c->wb_fqkv);
    cudaFree(c->ws);
    cudaFree(c->dech);
    cudaFree(c->logits);
    cudaFree(c->cp_state);
    cudaFree(c->hd_cdf);
    cudaFree(c->hd_max);
    cudaFree(c->hd_logZ);
    cudaFree(c->rj_buf);
    cudaFree(c->rj_logp);
    cudaFree(c->flt_ws);
    cudaFree(c->attn_dbg);
    delete c;
    return 0;
}

int pfn_set_option(pfn_ctx* c, const char* key, int64_t value) {
    PFN_REQUIRE(c && key, "null argument");
    if (!strcmp(key, "attn_impl")) { c->attn_impl = (int)value; return 0; }
    if (!strcmp(key, "gemm_impl")) { c->gemm_impl = (int)value; return 0; }
    if (!strcmp(key, "standardize_y")) { c->standardize_y = (int)value; return 0; }
    if (!strcmp(key, "attn_persist")) { c->attn_persist = (int)value; return 0; }
    if (!strcmp(key, "attn_wait_ticks")) { c->attn_wait_ticks = (int)value; return 0; }
    if (!strcmp(key, "attn_stagger_ns")) { c->attn_stagger_ns = (int)value; return 0; }
    if (!strcmp(key, "attn_lean")) { c->attn_lean = (int)value; return 0; }
    if (!strcmp(key, "mlp_fused")) { c->mlp_fused = (int)value; return 0; }
    if (!strcmp(key, "head_impl")) { c->head_impl = (int)value; return 0; }
    if (!strcmp(key, "head_threads")) {
        PFN_REQUIRE(value == 64 || value == 128 || value == 256, "head_threads must be 64, 128 or 256");
        c->head_threads = (int)value;
        return 0;
    }
    if (!strcmp(key, "feat_fused")) { c->feat_fused = (int)value; return 0; }
    if (!strcmp(key, "dec_rows")) {  // rows per decoder + head pass (logits workspace = dec_rows x num_buckets fp32)
        PFN_REQUIRE(value >= 128 && value <= (1 << 20), "dec_rows out of range");
        PFN_CUDA_OK(cudaSetDevice(c->device));
        PFN_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(c->dech); cudaFree(c->logits);
        c->dech = nullptr; c->logits = nullptr;
        c->dec_rows = (int)value;
        return 0;
    }
    if (!strcmp(key, "attn_debug")) {
        if (value && !c->attn_dbg) PFN_CUDA_OK(cudaMalloc(&c->attn_dbg, 3 * sizeof(unsigned long long)));
        if (value) PFN_CUDA_OK(cudaMemset(c->attn_dbg, 0, 3 * sizeof(unsigned long long)));
        c->attn_debug = (int)value;
        return 0;
    }
    if (!strcmp(key, "attn_poly")) {
        const int64_t k = value % 100;
        PFN_REQUIRE(value == 106 || (value >= 0 && value < 100 && (k == 0 || (k >= 3 && k <= 8))),
                    "attn_poly must be 0 or k in 3..8 (106: k = 6 with the degree-2 polynomial)");
        c->attn_poly = (int)value;
        return 0;
    }
    if (!strcmp(key, "chunk_rows")) { c->cfg.chunk_rows = (int)value; return 0; }
    if (!strcmp(key, "time_kernels")) { c->time_kernels = (int)value; return 0; }
    if (!strcmp(key, "sub_tokens")) { c->sub_tok_opt = value; return 0; }
    last_error() = std::string("unknown option ") + key;
    return 2;
}

int pfn_prefill(pfn_ctx* c, int slot, const float* X, int64_t ldx, const float* y, int64_t N, int F, void* stream) {
    if (int rc = check_slot(c, slot, false)) return rc;
    PFN_REQUIRE(X && y, "null data pointer");
    PFN_REQUIRE(N >= 1, "context needs at least one row");
    PFN_REQUIRE(F >= 1 && F <= 2 * c->cfg.max_groups, "feature count out of range");
    PFN_REQUIRE(ldx >= F, "row stride smaller than feature count");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    Slot& s = c->slots[slot];
    s.valid = false;
    s.N = N; s.F = F; s.G = (F + 1) / 2; s.T = s.G + 1;
    const size_t need = (size_t)c->cfg.nlayers * s.T * N * kKvRow * sizeof(bf16);
    if (need > s.kv_cap) {
        PFN_CUDA_OK(cudaStreamSynchronize(st));
        if (s.kv) PFN_CUDA_OK(cudaFree(s.kv));
        s.kv = nullptr; s.kv_cap = 0;
        PFN_CUDA_OK(cudaMalloc(&s.kv, need));
        s.kv_cap = need;
    }
    fit_stats_kernel<<<2 * s.G + 1, 256, 0, st>>>(X, ldx, y, N, F, s.G, s.enc, c->standardize_y);
    PFN_LAUNCH_OK(c);
    const int nb1 = c->cfg.num_buckets + 1;
    fit_finalize_kernel<<<(unsigned)ceil_div(std::max(nb1, s.G), 256), 256, 0, st>>>(s.enc, s.G, c->wf + c->off.borders,
                                                                                    nb1, s.borders);
    PFN_LAUNCH_OK(c);
    if (int rc = forward_rows(c, s, X, ldx, y, N, st)) return rc;
    s.valid = true;
    return 0;
}

int pfn_forward_logits(pfn_ctx* c, int slot, const float* X, int64_t ldx, int64_t M, float* out, int64_t ld_out,
                       void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    PFN_REQUIRE(M >= 0, "negative row count");
    if (M == 0) return 0;
    PFN_REQUIRE(X && out, "null data pointer");
    Slot& s = c->slots[slot];
    PFN_REQUIRE(ldx >= s.F, "row stride smaller than the slot's feature count");
    PFN_REQUIRE(ld_out >= c->cfg.num_buckets && ld_out % 4 == 0, "logits row stride must be a multiple of 4 and >= num_buckets");
    PFN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "logits buffer must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    if (int rc = ensure_decoder_ws(c)) return rc;
    const int64_t chunk = chunk_rows_of(c);
    for (int64_t r0 = 0; r0 < M; r0 += chunk) {
        const int64_t R = std::min(chunk, M - r0);
        if (int rc = forward_rows(c, s, X + r0 * ldx, ldx, nullptr, R, st)) return rc;
        for (int64_t d0 = 0; d0 < R; d0 += c->dec_rows) {
            const int64_t n = std::min<int64_t>(c->dec_rows, R - d0);
            if (int rc = decode_rows(c, s, d0, n, out + (r0 + d0) * ld_out, ld_out, st)) return rc;
        }
    }
    return 0;
}

int pfn_head_sample(pfn_ctx* c, int slot, const float* logits, int64_t ld_logits, int64_t group, int64_t M,
                    const float* uniforms, uint64_t seed, uint64_t row0, uint64_t offset, float* out_theta, int64_t ld_theta, int32_t* out_bin,
                    float* out_u, float* out_logp, float eps, int accumulate, void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    if (M == 0) return 0;
    PFN_REQUIRE(logits && out_theta, "null data pointer");
    PFN_REQUIRE(group >= 1, "group must be >= 1");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    HeadArgs h{};
    h.logits = logits; h.ld_logits = ld_logits; h.group = group; h.M = M; h.B = c->cfg.num_buckets; h.borders = c->slots[slot].borders;
    h.uniforms = uniforms; h.seed = seed; h.row0 = row0; h.offset = offset;
    h.out_theta = out_theta; h.ld_theta = ld_theta; h.out_bin = out_bin; h.out_u = out_u; h.out_logp = out_logp;
    h.log_eps = std::log((float)eps); h.accumulate = accumulate;
    return launch_head(c, h, true, (cudaStream_t)stream);
}

int pfn_head_nll(pfn_ctx* c, int slot, const float* logits, int64_t ld_logits, int64_t group, int64_t M, const float* y, int64_t ld_y,
                 float* out_nll, float* out_logp, float eps, int accumulate, void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    if (M == 0) return 0;
    PFN_REQUIRE(logits && y, "null data pointer");
    PFN_REQUIRE(group >= 1, "group must be >= 1");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    HeadArgs h{};
    h.logits = logits; h.ld_logits = ld_logits; h.group = group; h.M = M; h.B = c->cfg.num_buckets; h.borders = c->slots[slot].borders;
    h.y = y; h.ld_y = ld_y; h.out_nll = out_nll; h.out_logp = out_logp;
    h.log_eps = std::log((float)eps); h.accumulate = accumulate;
    return launch_head(c, h, false, (cudaStream_t)stream);
}

static int fused_step(pfn_ctx* c, int slot, const float* X, int64_t ldx, int64_t M, bool sample, const float* uniforms,
                      uint64_t seed, uint64_t row0, uint64_t offset, float* out_theta, int64_t ld_theta,
                      int32_t* out_bin, const float* y, int64_t ld_y, float* out_logp, float eps, int accumulate,
                      void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    PFN_REQUIRE(M >= 0, "negative row count");
    if (M == 0) return 0;
    PFN_REQUIRE(X, "null data pointer");
    Slot& s = c->slots[slot];
    PFN_REQUIRE(ldx >= s.F, "row stride smaller than the slot's feature count");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    if (int rc = ensure_decoder_ws(c)) return rc;
    const int64_t chunk = chunk_rows_of(c);
    const int B = c->cfg.num_buckets;
    for (int64_t r0 = 0; r0 < M; r0 += chunk) {
        const int64_t R = std::min(chunk, M - r0);
        if (int rc = forward_rows(c, s, X + r0 * ldx, ldx, nullptr, R, st)) return rc;
        for (int64_t d0 = 0; d0 < R; d0 += c->dec_rows) {
            const int64_t n = std::min<int64_t>(c->dec_rows, R - d0);
            if (int rc = decode_rows(c, s, d0, n, c->logits, B, st)) return rc;
            const int64_t g0 = r0 + d0;
            HeadArgs h{};
            h.logits = c->logits; h.ld_logits = B; h.group = 1; h.M = n; h.B = B; h.borders = s.borders;
            h.log_eps = std::log((float)eps); h.accumulate = accumulate;
            h.out_logp = out_logp ? out_logp + g0 : nullptr;
            if (sample) {
                h.uniforms = uniforms ? uniforms + g0 : nullptr;
                h.seed = seed; h.row0 = row0 + (uint64_t)g0; h.offset = offset;
                h.out_theta = out_theta + g0 * ld_theta; h.ld_theta = ld_theta;
                h.out_bin = out_bin ? out_bin + g0 : nullptr;
            } else {
                h.y = y + g0 * ld_y; h.ld_y = ld_y;
            }
            if (int rc = launch_head(c, h, sample, st)) return rc;
        }
    }
    return 0;
}

int pfn_sample(pfn_ctx* c, int slot, const float* X, int64_t ldx, int64_t M, const float* uniforms, uint64_t seed,
               uint64_t row0, uint64_t offset, float* out_theta, int64_t ld_theta, int32_t* out_bin, float* out_logp,
               float eps, int accumulate, void* stream) {
    PFN_REQUIRE(out_theta || M == 0, "null output pointer");
    return fused_step(c, slot, X, ldx, M, true, uniforms, seed, row0, offset, out_theta, ld_theta, out_bin, nullptr, 0,
                      out_logp, eps, accumulate, stream);
}

int pfn_logprob(pfn_ctx* c, int slot, const float* X, int64_t ldx, int64_t M, const float* y, int64_t ld_y,
                float* out_logp, float eps, int accumulate, void* stream) {
    PFN_REQUIRE((y && out_logp) || M == 0, "null data pointer");
    return fused_step(c, slot, X, ldx, M, false, nullptr, 0, 0, 0, nullptr, 0, nullptr, y, ld_y, out_logp, eps,
                      accumulate, stream);
}

static int launch_compact(pfn_ctx* c, CompactArgs a, cudaStream_t st) {
    const int64_t nblocks = ceil_div(a.M, CP_TILE);
    if (nblocks > c->cp_cap) {
        PFN_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(c->cp_state);
        c->cp_state = nullptr; c->cp_cap = 0;
        PFN_CUDA_OK(cudaMalloc(&c->cp_state, (size_t)(nblocks + 1024 + 1) * 8));
        c->cp_cap = nblocks + 1024;
    }
    PFN_CUDA_OK(cudaMemsetAsync(c->cp_state, 0, (size_t)(nblocks + 1) * 8, st));
    a.tile_state = c->cp_state;
    a.ticket = reinterpret_cast<unsigned int*>(c->cp_state + nblocks);
    TimeScope ts(c, st, KC_COMPACT, 0.0, (double)a.M * a.dim * 4.0);
    compact_append_kernel<<<(unsigned)nblocks, CP_THREADS, 0, st>>>(a);
    PFN_LAUNCH_OK(c);
    return 0;
}

int pfn_accept_compact(pfn_ctx* c, const float* theta, int64_t ld, int64_t M, int dim, const float* lo, const float* hi,
                       const uint8_t* mask, int64_t* out_idx, float* out_rows, int64_t* out_count, void* stream) {
    PFN_REQUIRE(c && out_count, "null argument");
    PFN_REQUIRE(M >= 0 && dim >= 1 && ld >= dim, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    if (M == 0) {
        PFN_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st));
        return 0;
    }
    PFN_REQUIRE(theta, "null data pointer");
    CompactArgs a{};
    a.theta = theta; a.ld = ld; a.M = M; a.dim = dim; a.lo = lo; a.hi = hi; a.mask = mask;
    a.out_idx = out_idx; a.out_rows = out_rows; a.out_ld = dim; a.capacity = M; a.out_count = out_count;
    return launch_compact(c, a, st);
}

int pfn_accept_append(pfn_ctx* c, const float* theta, int64_t ld, int64_t M, int dim, const float* lo, const float* hi,
                      const uint8_t* mask, const float* score, const float* thr, const float* logp, float* out_rows,
                      int64_t out_ld, float* out_logp, int64_t capacity, int64_t* cursor, void* stream) {
    PFN_REQUIRE(c && cursor, "null argument");
    PFN_REQUIRE(M >= 0 && dim >= 1 && ld >= dim && out_ld >= dim && capacity >= 0, "bad shape");
    PFN_REQUIRE((score == nullptr) == (thr == nullptr), "score and thr come together");
    if (M == 0) return 0;
    PFN_REQUIRE(theta && out_rows, "null data pointer");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    CompactArgs a{};
    a.theta = theta; a.ld = ld; a.M = M; a.dim = dim; a.lo = lo; a.hi = hi; a.mask = mask; a.score = score; a.thr = thr;
    a.logp = logp; a.out_rows = out_rows; a.out_ld = out_ld; a.out_logp = out_logp; a.capacity = capacity; a.cursor = cursor;
    return launch_compact(c, a, (cudaStream_t)stream);
}

int pfn_uniform_box(pfn_ctx* c, const float* lo, const float* hi, int64_t M, int dim, uint64_t seed, uint64_t row0,
                    float* out, int64_t ld, void* stream) {
    PFN_REQUIRE(c && lo && hi && (out || M == 0), "null argument");
    PFN_REQUIRE(M >= 0 && dim >= 1 && ld >= dim, "bad shape");
    if (M == 0) return 0;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    uniform_box_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(lo, hi, M, dim, seed, row0, out, ld);
    PFN_LAUNCH_OK(c);
    return 0;
}

int pfn_sample_rejection(pfn_ctx* c, const int32_t* slots, const float* x_obs, int dx, int dtheta, int n_rounds,
                         int64_t round_rows, const float* lo, const float* hi, uint64_t seed, uint64_t row0, float eps,
                         float* out_theta, int64_t out_ld, float* out_logp, int64_t capacity, int64_t* cursor,
                         void* stream) {
    PFN_REQUIRE(c && slots && x_obs && out_theta && cursor, "null argument");
    PFN_REQUIRE(dx >= 1 && dtheta >= 1 && n_rounds >= 0 && round_rows >= 1 && out_ld >= dtheta && capacity >= 0, "bad shape");
    for (int d = 0; d < dtheta; ++d) {
        if (int rc = check_slot(c, slots[d], true)) return rc;
        PFN_REQUIRE(c->slots[slots[d]].F == dx + d, "slot d must hold the context of dimension d (F = dx + d features)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    if (int rc = ensure_decoder_ws(c)) return rc;
    const int64_t ld = dx + dtheta;
    if (round_rows > c->rj_rows || ld > c->rj_ld) {
        PFN_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(c->rj_buf); cudaFree(c->rj_logp);
        c->rj_buf = nullptr; c->rj_logp = nullptr; c->rj_rows = 0; c->rj_ld = 0;
        PFN_CUDA_OK(cudaMalloc(&c->rj_buf, (size_t)round_rows * ld * 4));
        PFN_CUDA_OK(cudaMalloc(&c->rj_logp, (size_t)round_rows * 4));
        c->rj_rows = round_rows; c->rj_ld = ld;
    }
    const int B = c->cfg.num_buckets;
    const float log_eps = std::log(eps);
    for (int k = 0; k < n_rounds; ++k) {
        const uint64_t r0 = row0 + (uint64_t)k * (uint64_t)round_rows;
        // x_o into every row of the joint test matrix (npe_pfn.py:121-122), log-prob accumulator to zero
        broadcast_rows_kernel<<<(unsigned)ceil_div(round_rows * dx, 256), 256, 0, st>>>(x_obs, dx, round_rows, c->rj_buf, ld);
        PFN_LAUNCH_OK(c);
        if (out_logp) PFN_CUDA_OK(cudaMemsetAsync(c->rj_logp, 0, (size_t)round_rows * 4, st));
        float* lp = out_logp ? c->rj_logp : nullptr;
        for (int d = 0; d < dtheta; ++d) {
            Slot& s = c->slots[slots[d]];
            if (d == 0) {
                // all rows are the same observation: one forward row, round_rows inverse-CDF draws from its logits
                if (int rc = forward_rows(c, s, c->rj_buf, ld, nullptr, 1, st)) return rc;
                if (int rc = decode_rows(c, s, 0, 1, c->logits, B, st)) return rc;
                HeadArgs h{};
                h.logits = c->logits; h.ld_logits = 0; h.group = 1; h.M = round_rows; h.B = B; h.borders = s.borders;
                h.seed = seed; h.row0 = r0; h.offset = 0; h.out_theta = c->rj_buf + dx; h.ld_theta = ld;
                h.out_logp = lp; h.log_eps = log_eps; h.accumulate = 1;
                if (int rc = launch_head(c, h, true, st)) return rc;
            } else {
                if (int rc = fused_step(c, slots[d], c->rj_buf, ld, round_rows, true, nullptr, seed, r0, (uint64_t)d,
                                        c->rj_buf + dx + d, ld, nullptr, nullptr, 0, lp, eps, 1, stream))
                    return rc;
            }
        }
        CompactArgs a{};
        a.theta = c->rj_buf + dx; a.ld = ld; a.M = round_rows; a.dim = dtheta; a.lo = lo; a.hi = hi; a.logp = lp;
        a.out_rows = out_theta; a.out_ld = out_ld; a.out_logp = out_logp; a.capacity = capacity; a.cursor = cursor;
        if (int rc = launch_compact(c, a, st)) return rc;
    }
    return 0;
}

int pfn_filter_context(pfn_ctx* c, const float* x_train, int64_t ld, int64_t Ntot, int dx, const float* obs, int64_t k,
                       int64_t* out_idx, float* out_dist, void* stream) {
    PFN_REQUIRE(c && x_train && obs && out_idx, "null argument");
    PFN_REQUIRE(Ntot >= 1 && dx >= 1 && ld >= dx, "bad shape");
    PFN_REQUIRE(k >= 1 && k <= Ntot, "k must be in [1, Ntot]");
    PFN_REQUIRE(k <= FLT_MAX_K, "k above 16384 is not supported by the single-CTA sort");
    PFN_REQUIRE(Ntot < (1ll << 32), "at most 2^32 - 1 simulations");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    if (Ntot > c->flt_cap) {
        PFN_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(c->flt_ws);
        c->flt_ws = nullptr; c->flt_cap = 0;
        // dist f32 | mask_lt u8 | mask_eq u8 | idx_lt i64 | idx_eq i64 | stats | hist | state | counts
        const size_t bytes = (size_t)Ntot * (4 + 1 + 1 + 8 + 8) + 64 * 1024;
        PFN_CUDA_OK(cudaMalloc(&c->flt_ws, bytes));
        c->flt_cap = Ntot;
    }
    uint8_t* p = c->flt_ws;
    int64_t* idx_lt = reinterpret_cast<int64_t*>(p); p += (size_t)c->flt_cap * 8;
    int64_t* idx_eq = reinterpret_cast<int64_t*>(p); p += (size_t)c->flt_cap * 8;
    float* dist = reinterpret_cast<float*>(p); p += (size_t)c->flt_cap * 4;
    uint8_t* mask_lt = p; p += (size_t)c->flt_cap;
    uint8_t* mask_eq = p; p += (size_t)c->flt_cap;
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
    float* stats = reinterpret_cast<float*>(p); p += 4096;
    unsigned int* hist = reinterpret_cast<unsigned int*>(p); p += 1024;
    unsigned long long* state = reinterpret_cast<unsigned long long*>(p); p += 64;
    int64_t* counts = reinterpret_cast<int64_t*>(p);
    PFN_REQUIRE(dx * 2 * 4 <= 4096, "too many x columns");

    filter_stats_kernel<<<dx, 256, 0, st>>>(x_train, ld, Ntot, dx, stats);
    PFN_LAUNCH_OK(c);
    filter_dist_kernel<<<(unsigned)ceil_div(Ntot, 256), 256, 0, st>>>(x_train, ld, Ntot, dx, obs, stats, dist);
    PFN_LAUNCH_OK(c);
    const unsigned long long init[3] = {0ull, (unsigned long long)(k - 1), 0ull};
    PFN_CUDA_OK(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    PFN_CUDA_OK(cudaMemsetAsync(hist, 0, 1024, st));
    const unsigned hb = (unsigned)std::min<int64_t>(ceil_div(Ntot, 256), 148 * 8);
    for (int shift = 24; shift >= 0; shift -= 8) {
        filter_hist_kernel<<<hb, 256, 0, st>>>(dist, Ntot, shift, state, hist);
        PFN_LAUNCH_OK(c);
        filter_pick_kernel<<<1, 32, 0, st>>>(hist, shift, state);
        PFN_LAUNCH_OK(c);
    }
    filter_mask_kernel<<<(unsigned)ceil_div(Ntot, 256), 256, 0, st>>>(dist, Ntot, state, mask_lt, mask_eq);
    PFN_LAUNCH_OK(c);
    // ordered compaction of the two masks (row "theta" is unused: dim 1 view of dist keeps the finite check cheap)
    if (int rc = pfn_accept_compact(c, dist, 1, Ntot, 1, nullptr, nullptr, mask_lt, idx_lt, nullptr, counts, stream)) return rc;
    if (int rc = pfn_accept_compact(c, dist, 1, Ntot, 1, nullptr, nullptr, mask_eq, idx_eq, nullptr, counts + 1, stream)) return rc;
    int npow2 = 1;
    while (npow2 < k) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 8;
    if (smem > 48 * 1024)
        PFN_CUDA_OK(cudaFuncSetAttribute(filter_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_sort_kernel<<<1, 1024, smem, st>>>(dist, idx_lt, counts, idx_eq, k, npow2, out_idx, out_dist);
    PFN_LAUNCH_OK(c);
    return 0;
}

int pfn_slot_info(pfn_ctx* c, int slot, int64_t* N, int32_t* F, int32_t* T, int64_t* kv_bytes) {
    if (int rc = check_slot(c, slot, true)) return rc;
    const Slot& s = c->slots[slot];
    if (N) *N = s.N;
    if (F) *F = s.F;
    if (T) *T = s.T;
    if (kv_bytes) *kv_bytes = (int64_t)c->cfg.nlayers * s.T * s.N * kKvRow * (int64_t)sizeof(bf16);
    return 0;
}

int64_t pfn_launch_count(pfn_ctx* c) { return c ? c->launches : 0; }

int pfn_kernel_times(pfn_ctx* c, double* ms, int64_t* counts, double* flops, double* bytes, int n_classes, int reset) {
    PFN_REQUIRE(c && ms && counts && flops && bytes, "null argument");
    PFN_REQUIRE(n_classes >= KC_COUNT, "output arrays are shorter than the number of kernel classes");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    PFN_CUDA_OK(cudaDeviceSynchronize());
    for (int k = 0; k < n_classes; ++k) { ms[k] = 0.0; counts[k] = 0; flops[k] = 0.0; bytes[k] = 0.0; }
    for (auto& t : c->timed) {
        float e = 0.f;
        PFN_CUDA_OK(cudaEventElapsedTime(&e, t.a, t.b));
        ms[t.cls] += e; counts[t.cls] += 1; flops[t.cls] += t.flops; bytes[t.cls] += t.bytes;
    }
    if (reset) {
        for (auto& t : c->timed) { c->ev_pool.push_back(t.a); c->ev_pool.push_back(t.b); }
        c->timed.clear();
    }
    return 0;
}

int pfn_slot_export(pfn_ctx* c, int slot, float* stats, float* y_stats, float* borders, void* kv, void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    const Slot& s = c->slots[slot];
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    const int Fp = 2 * s.G;
    if (stats) {
        PFN_CUDA_OK(cudaMemcpyAsync(stats, s.enc + kEncMean, Fp * 4, cudaMemcpyDeviceToDevice, st));
        PFN_CUDA_OK(cudaMemcpyAsync(stats + Fp, s.enc + kEncStd, Fp * 4, cudaMemcpyDeviceToDevice, st));
        PFN_CUDA_OK(cudaMemcpyAsync(stats + 2 * Fp, s.enc + kEncScale, s.G * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (y_stats) PFN_CUDA_OK(cudaMemcpyAsync(y_stats, s.enc + kEncY, 3 * 4, cudaMemcpyDeviceToDevice, st));
    if (borders)
        PFN_CUDA_OK(cudaMemcpyAsync(borders, s.borders, (size_t)(c->cfg.num_buckets + 1) * 4, cudaMemcpyDeviceToDevice, st));
    if (kv) {
        const size_t bytes = (size_t)c->cfg.nlayers * s.T * s.N * kKvRow * sizeof(bf16);
        PFN_CUDA_OK(cudaMemcpyAsync(kv, s.kv, bytes, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int pfn_slot_import(pfn_ctx* c, int slot, int64_t N, int F, const float* enc_state, const float* borders, const void* kv,
                    void* stream) {
    if (int rc = check_slot(c, slot, false)) return rc;
    PFN_REQUIRE(enc_state && borders && kv, "null data pointer");
    PFN_REQUIRE(N >= 1 && F >= 1 && F <= 2 * c->cfg.max_groups, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    Slot& s = c->slots[slot];
    s.valid = false;
    s.N = N; s.F = F; s.G = (F + 1) / 2; s.T = s.G + 1;
    const size_t need = (size_t)c->cfg.nlayers * s.T * N * kKvRow * sizeof(bf16);
    if (need > s.kv_cap) {
        PFN_CUDA_OK(cudaStreamSynchronize(st));
        if (s.kv) PFN_CUDA_OK(cudaFree(s.kv));
        s.kv = nullptr; s.kv_cap = 0;
        PFN_CUDA_OK(cudaMalloc(&s.kv, need));
        s.kv_cap = need;
    }
    PFN_CUDA_OK(cudaMemcpyAsync(s.enc, enc_state, kEncFloats * 4, cudaMemcpyDeviceToDevice, st));
    PFN_CUDA_OK(cudaMemcpyAsync(s.borders, borders, (size_t)(c->cfg.num_buckets + 1) * 4, cudaMemcpyDeviceToDevice, st));
    PFN_CUDA_OK(cudaMemcpyAsync(s.kv, kv, need, cudaMemcpyDeviceToDevice, st));
    s.valid = true;
    return 0;
}

int pfn_slot_state(pfn_ctx* c, int slot, float* enc_state, void* stream) {
    if (int rc = check_slot(c, slot, true)) return rc;
    PFN_REQUIRE(enc_state, "null data pointer");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    PFN_CUDA_OK(cudaMemcpyAsync(enc_state, c->slots[slot].enc, kEncFloats * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int pfn_debug_last_states(pfn_ctx* c, float* out, int64_t max_floats, void* stream) {
    PFN_REQUIRE(c && out, "null argument");
    const int64_t n = c->last_rows * c->last_T * kE;
    PFN_REQUIRE(n > 0 && n <= max_floats, "no forward has run or output buffer too small");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    PFN_CUDA_OK(cudaMemcpyAsync(out, c->xf, (size_t)n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int pfn_attn_debug_counts(pfn_ctx* c, uint64_t* out3) {
    PFN_REQUIRE(c && out3, "null argument");
    PFN_REQUIRE(c->attn_dbg, "set option attn_debug first");
    PFN_CUDA_OK(cudaSetDevice(c->device));
    PFN_CUDA_OK(cudaDeviceSynchronize());
    PFN_CUDA_OK(cudaMemcpy(out3, c->attn_dbg, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}

int pfn_member_transform(pfn_ctx* c, const pfn_member_desc* d, const float* X, int64_t ldx, int64_t M, float* out,
                         int64_t ld_out, void* stream) {
    PFN_REQUIRE(c && d && (M == 0 || (X && out)), "null argument");
    PFN_REQUIRE(d->kind >= 0 && d->kind <= 2, "member kind must be 0 (quantile), 1 (safepower) or 2 (none)");
    PFN_REQUIRE(d->n_features_in >= 1 && d->n_keep >= 1 && d->n_keep <= d->n_features_in && d->keep, "bad kept-feature list");
    const int n_el = d->kind == 0 ? 2 * d->n_keep : d->n_keep;
    const int svd_k = d->kind == 0 ? d->svd_k : 0;
    const int n_base = n_el + svd_k + (d->fingerprint ? 1 : 0);
    PFN_REQUIRE(n_base <= kMaxBase && d->n_out == n_base && d->perm, "member feature count out of range");
    PFN_REQUIRE(d->kind == 0 ? (d->n_quantiles >= 2 && d->quantiles) : (d->kind == 2 || d->safepower != nullptr), "missing member tables");
    PFN_REQUIRE(svd_k == 0 || (d->svd_inv_scale && d->svd_vt), "missing SVD tables");
    PFN_REQUIRE(ldx >= d->n_features_in && ld_out >= d->n_out, "row stride smaller than feature count");
    if (M == 0) return 0;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    MemberXformArgs a{};
    a.X = X; a.ldx = ldx; a.M = M; a.F_in = d->n_features_in; a.out = out; a.ld_out = ld_out;
    a.n_keep = d->n_keep; a.keep = d->keep; a.kind = d->kind; a.nq = d->n_quantiles; a.quantiles = d->quantiles;
    a.sp = d->safepower; a.svd_k = svd_k; a.svd_inv_scale = d->svd_inv_scale; a.svd_vt = d->svd_vt;
    a.fingerprint = d->fingerprint; a.n_out = d->n_out; a.perm = d->perm;
    TimeScope ts(c, st, KC_OTHER, 0.0);
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(M, XF_WARPS), (int64_t)c->num_sms * 8);
    member_transform_kernel<<<grid, XF_WARPS * 32, 0, st>>>(a);
    PFN_CUDA_OK(cudaGetLastError());
    c->launches++;
    return 0;
}

int pfn_ensemble_combine(pfn_ctx* c, const float* logits, int64_t ld_logits, int64_t member_stride, int n_members, int64_t M,
                         const int32_t* idx, const float* frac, const uint8_t* valid, float* out, int64_t ld_out,
                         void* stream) {
    PFN_REQUIRE(c && (M == 0 || (logits && out)) && idx && frac && valid, "null argument");
    PFN_REQUIRE(n_members >= 1, "need at least one member");
    const int B = c->cfg.num_buckets;
    PFN_REQUIRE(ld_logits >= B && ld_out >= B, "row stride smaller than the bucket count");
    if (M == 0) return 0;
    PFN_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = combine_smem_bytes(B);
    static bool configured_dev[64] = {};
    if (!configured_dev[c->device & 63]) {
        PFN_CUDA_OK(cudaFuncSetAttribute(ensemble_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev[c->device & 63] = true;
    }
    CombineArgs a{};
    a.logits = logits; a.ld_logits = ld_logits; a.member_stride = member_stride; a.E = n_members; a.B = B; a.M = M;
    a.idx = idx; a.frac = frac; a.valid = valid; a.out = out; a.ld_out = ld_out;
    TimeScope ts(c, st, KC_OTHER, 0.0);
    const unsigned grid = (unsigned)std::min<int64_t>(M, (int64_t)c->num_sms * 2);
    ensemble_combine_kernel<<<grid, CB_THREADS, smem, st>>>(a);
    PFN_CUDA_OK(cudaGetLastError());
    c->launches++;
    return 0;
}

}  // extern "C"
