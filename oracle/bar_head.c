/* CPU ORACLE — test infrastructure only (never linked into the product).
 *
 * C restatement of the bar-distribution head the reference reaches through
 *   pred_dist["criterion"].sample(pred_dist["logits"])      /root/reference/npe_pfn/npe_pfn.py:146, 220
 *   pred_dist["criterion"](pred_dist["logits"], y)          /root/reference/npe_pfn/npe_pfn.py:149-151, 226-228, 510-512
 * i.e. upstream tabpfn 2.2.1 `FullSupportBarDistribution` (absent offline; algorithm
 * as published):
 *   sample: p = softmax(logits); c = cumsum(p); idx = searchsorted(c, u) clamped;
 *           theta = border[idx] + width[idx] * (u - c[idx-1]) / p[idx]
 *   nll:    idx = searchsorted(borders, y) - 1 clamped; logp = log_softmax[idx] - log width[idx];
 *           half-normal tails on the first / last bucket; returns -logp.
 * PARITY UNPINNED against tabpfn itself (no golden values exist in /root/reference);
 * pinned against the textbook fp32 torch formulation in tests/test_oracle.py.
 *
 * Arithmetic is specified so that ANY implementation following the spec is bit-identical
 * for bucket indices and samples (DESIGN.md "head arithmetic"):
 *   e_i  = exp_det(logit_i - max)          fp32, fixed fmaf sequence below
 *   q_i  = trunc(e_i * 2^40)               uint64 (exact; order-independent integer prefix sums)
 *   Z    = sum q_i ; target = (double)u * (double)Z
 *   idx  = #{k : C_k < target}, C_k = q_0+..+q_k   (searchsorted side='left'); frac = 0 if q_idx = 0
 *   theta = lo + (hi - lo) * ((target - C_{idx-1}) / q_idx)       in double, borders already
 *           renormalised to original units in fp32: b' = fmaf(b, y_std, y_mean)
 * Uniforms are either injected or Philox4x32-10(seed; counter = row, offset), u = ((w0>>9) + 0.5) * 2^-23, i.e. strictly inside (0, 1).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define PFN_Q_SHIFT 40

/* ---- deterministic exp (t <= 0) ---------------------------------------------------- */
static const float kLog2e = 0x1.715476p+0f;
static const float kC0 = 0x1.000000p+0f, kC1 = 0x1.62e428p-1f, kC2 = 0x1.ebfdf2p-3f,
                   kC3 = 0x1.c67f50p-5f, kC4 = 0x1.3d54d8p-7f, kC5 = 0x1.44d4d2p-10f,
                   kC6 = 0x1.ca8f0ap-13f;

float pfn_oracle_exp_det(float t) {
    if (!(t > -64.0f)) t = -64.0f; /* also maps NaN to the floor */
    float x = t * kLog2e;
    float n = floorf(x);
    float f = x - n;
    float p = kC6;
    p = fmaf(p, f, kC5);
    p = fmaf(p, f, kC4);
    p = fmaf(p, f, kC3);
    p = fmaf(p, f, kC2);
    p = fmaf(p, f, kC1);
    p = fmaf(p, f, kC0);
    uint32_t sb = (uint32_t)((int)n + 127) << 23; /* 2^n, n in [-93, 0] */
    float s;
    memcpy(&s, &sb, 4);
    return p * s;
}

/* trunc(e * 2^40) by integer shifts on the fp32 encoding */
uint64_t pfn_oracle_quantize(float e) {
    uint32_t b;
    memcpy(&b, &e, 4);
    int ex = (int)((b >> 23) & 0xff);
    if (ex == 0) return 0;
    uint64_t m = (uint64_t)((b & 0x7fffffu) | 0x800000u);
    int sh = ex - 127 - 23 + PFN_Q_SHIFT;
    if (sh >= 0) return m << sh;
    if (sh <= -64) return 0;
    return m >> (-sh);
}

/* ---- Philox4x32-10 --------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void pfn_oracle_philox4x32(uint64_t seed, uint64_t row, uint64_t offset, uint32_t out[4]) {
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, 16);
}

float pfn_oracle_philox_uniform(uint64_t seed, uint64_t row, uint64_t offset) {
    uint32_t w[4];
    pfn_oracle_philox4x32(seed, row, offset, w);
    return ((float)(w[0] >> 9) + 0.5f) * 0x1.0p-23f; /* in (0, 1): never 0, never 1 */
}

/* ---- borders --------------------------------------------------------------------- */
void pfn_oracle_renorm_borders(const float* borders, int nb1, float y_mean, float y_std, float* out) {
    for (int i = 0; i < nb1; ++i) out[i] = fmaf(borders[i], y_std, y_mean);
}

/* ---- inverse-CDF sample for one row ------------------------------------------------ */
void pfn_oracle_icdf_row(const float* logits, int B, const float* borders, float u,
                         float* theta_out, int32_t* idx_out) {
    float m = logits[0];
    for (int i = 1; i < B; ++i) m = logits[i] > m ? logits[i] : m;
    uint64_t Z = 0;
    for (int i = 0; i < B; ++i) Z += pfn_oracle_quantize(pfn_oracle_exp_det(logits[i] - m));
    double target = (double)u * (double)Z;
    uint64_t C = 0, Cprev = 0, q = 0;
    int idx = -1;
    for (int i = 0; i < B; ++i) {
        uint64_t qi = pfn_oracle_quantize(pfn_oracle_exp_det(logits[i] - m));
        uint64_t Cn = C + qi;
        if (idx < 0 && !((double)Cn < target)) { idx = i; Cprev = C; q = qi; }
        C = Cn;
    }
    if (idx < 0) { idx = B - 1; Cprev = Z; q = 0; } /* unreachable for u < 1 */
    double frac = q ? (target - (double)Cprev) / (double)q : 0.0;
    if (frac < 0.0) frac = 0.0;
    if (frac > 1.0) frac = 1.0;
    double lo = (double)borders[idx], hi = (double)borders[idx + 1];
    double th = lo + (hi - lo) * frac;
    *theta_out = (float)th;
    *idx_out = idx;
}

/* ---- negative log density for one row ---------------------------------------------- */
/* searchsorted(borders, y) - 1 (side='left'): #{k : borders[k] < y} - 1, clamped */
int32_t pfn_oracle_bucket_of(const float* borders, int B, float y) {
    int lo = 0, hi = B + 1; /* first k with borders[k] >= y */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (borders[mid] < y) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    if (idx < 0) idx = 0;
    if (idx > B - 1) idx = B - 1;
    return idx;
}

#define PFN_PI 3.14159265358979323846
#define PFN_HALFNORMAL_ICDF_HALF 0.6744897501960817 /* Phi^-1(0.75) */

static double halfnormal_logpdf(double v, double scale) {
    return log(2.0) - log(scale) - 0.5 * log(2.0 * PFN_PI) - 0.5 * (v / scale) * (v / scale);
}

double pfn_oracle_nll_row(const float* logits, int B, const float* borders, float y) {
    float m = logits[0];
    for (int i = 1; i < B; ++i) m = logits[i] > m ? logits[i] : m;
    double Z = 0.0;
    for (int i = 0; i < B; ++i) Z += exp((double)logits[i] - (double)m);
    int idx = pfn_oracle_bucket_of(borders, B, y);
    double width = (double)borders[idx + 1] - (double)borders[idx];
    double logp = ((double)logits[idx] - (double)m) - log(Z) - log(width);
    if (idx == 0) {
        double w0 = (double)borders[1] - (double)borders[0];
        double v = (double)borders[1] - (double)y;
        if (v < 1e-8) v = 1e-8;
        logp += halfnormal_logpdf(v, w0 / PFN_HALFNORMAL_ICDF_HALF) + log(w0);
    } else if (idx == B - 1) {
        double wl = (double)borders[B] - (double)borders[B - 1];
        double v = (double)y - (double)borders[B - 1];
        if (v < 1e-8) v = 1e-8;
        logp += halfnormal_logpdf(v, wl / PFN_HALFNORMAL_ICDF_HALF) + log(wl);
    }
    return -logp;
}

/* ---- batched entry points (what tests/ and bench.py call through ctypes) ------------ */
void pfn_oracle_sample(const float* logits, int64_t M, int B, const float* borders,
                       const float* uniforms /* or NULL */, uint64_t seed, uint64_t row0, uint64_t offset,
                       float* theta_out, int32_t* idx_out, float* u_out /* or NULL */) {
    for (int64_t r = 0; r < M; ++r) {
        float u = uniforms ? uniforms[r] : pfn_oracle_philox_uniform(seed, row0 + (uint64_t)r, offset);
        if (u_out) u_out[r] = u;
        pfn_oracle_icdf_row(logits + r * (int64_t)B, B, borders, u, theta_out + r, idx_out + r);
    }
}

void pfn_oracle_nll(const float* logits, int64_t M, int B, const float* borders, const float* y,
                    float* nll_out) {
    for (int64_t r = 0; r < M; ++r)
        nll_out[r] = (float)pfn_oracle_nll_row(logits + r * (int64_t)B, B, borders, y[r]);
}
