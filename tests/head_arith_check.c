/* TEST INFRASTRUCTURE.  CPU emulation of the 15-instruction bucket mass of head_row2_kernel
 * (npe_pfn_b200/csrc/head_kernels.cuh::mass_q40) against the specification in oracle/bar_head.c
 * (pfn_oracle_quantize(pfn_oracle_exp_det(t))), argument by argument:
 *     head_arith_check <stride>      walks every stride-th fp32 value t in [-64, -0] (stride 1 = all 1.1e9 of them), the
 *                                    values below -64, -inf and NaN, and +0
 * prints the number of arguments checked and of mismatches; exit code 1 on any mismatch.
 * FADD.RM is emulated with fesetround(FE_DOWNWARD) around the one addition (-frounding-math, volatile operands);
 * F2I.U64.TRUNC with the C conversion (truncation), fmaxf(NaN, c) = c like FMNMX.
 * Build: gcc -O2 -ffp-contract=off -frounding-math tests/head_arith_check.c oracle/bar_head.c -lm */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

float pfn_oracle_exp_det(float t);
uint64_t pfn_oracle_quantize(float e);

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#define BLK 65536
static float xs[BLK], trs[BLK];
static const float kMagic = 12582912.0f + 40.0f;

static uint64_t finish(float x, float tr) { /* everything after the round-down addition, round-to-nearest mode */
    float n = tr - kMagic;
    float f = x - n;
    float p = 0x1.ca8f0ap-13f;
    p = fmaf(p, f, 0x1.44d4d2p-10f);
    p = fmaf(p, f, 0x1.3d54d8p-7f);
    p = fmaf(p, f, 0x1.c67f50p-5f);
    p = fmaf(p, f, 0x1.ebfdf2p-3f);
    p = fmaf(p, f, 0x1.62e428p-1f);
    p = fmaf(p, f, 0x1.000000p+0f);
    float e40 = u2f(f2u(p) + (f2u(tr) << 23));
    return (uint64_t)e40;
}

static uint64_t checked = 0, bad = 0;
static float ts[BLK];
static int nblk = 0;

static void flush(void) {
    for (int i = 0; i < nblk; ++i) {
        float t = ts[i];
        if (!(t > -64.0f)) t = -64.0f; /* FMNMX(t, -64): NaN -> -64 */
        xs[i] = t * 0x1.715476p+0f;
    }
    fesetround(FE_DOWNWARD);
    for (int i = 0; i < nblk; ++i) {
        volatile float a = xs[i], b = kMagic;
        trs[i] = a + b;
    }
    fesetround(FE_TONEAREST);
    for (int i = 0; i < nblk; ++i) {
        uint64_t got = finish(xs[i], trs[i]);
        uint64_t want = pfn_oracle_quantize(pfn_oracle_exp_det(ts[i]));
        ++checked;
        if (got != want) {
            if (bad < 10) fprintf(stderr, "mismatch t=%a (0x%08x): got %llu want %llu\n", ts[i], f2u(ts[i]),
                                  (unsigned long long)got, (unsigned long long)want);
            ++bad;
        }
    }
    nblk = 0;
}
static void push(float t) {
    ts[nblk++] = t;
    if (nblk == BLK) flush();
}

int main(int argc, char** argv) {
    uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;
    if (!stride) stride = 1;
    const uint32_t lo = 0x80000000u, hi = 0xC2800000u; /* -0 .. -64 */
    for (uint64_t b = lo; b <= hi; b += stride) push(u2f((uint32_t)b));
    /* both ends of every binade and of every integer step of x = t log2(e), whatever the stride */
    for (uint32_t e = 1; e <= 0x85; ++e)
        for (int d = -3; d <= 3; ++d) push(u2f((uint32_t)(0x80000000u | (e << 23)) + (uint32_t)d));
    for (int k = 0; k <= 93; ++k) {
        float t = -(float)k * 0x1.62e430p-1f; /* near x = -k */
        for (int d = -40; d <= 40; ++d) { uint32_t u = f2u(t) + (uint32_t)d; if (u >= lo && u <= hi + 64) push(u2f(u)); }
    }
    push(0.0f); push(-64.0f); push(-65.0f); push(-1e30f); push(-INFINITY); push(NAN); push(u2f(0xFFC00001u));
    for (uint32_t b = hi; b < hi + 5000; ++b) push(u2f(b));
    flush();
    printf("checked %llu mismatches %llu\n", (unsigned long long)checked, (unsigned long long)bad);
    return bad ? 1 : 0;
}
