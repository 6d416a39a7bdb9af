/* q4_lut_equiv.c -- the table form of the 4-bit decision procedures in the product's fast quantize kernel
 * (bitsandbytes_b200/csrc/blockwise.cu: build_q4_lut / quantize4_lut -- one look-up in a 33- / 17-entry table of
 * {pivot inside the cell, code below | (below ^ above) << 8} plus ONE comparison) returns the same code as the
 * reference's decision trees (restated in blockwise.cu: quantize_nf4 / quantize_fp4, pinned to the oracle by
 * q4_tree_equiv.c) for EVERY fp32 value the kernel can feed it: x = a * rcp.approx(absmax) with |a| <= absmax,
 * i.e. |x| <= 1 + 2^-20 (generous), and NaN (absmax = 0 or inf).  The table build and the look-up are restated
 * verbatim, with fmaf / truncation as the GPU executes them (fma.rn.f32, cvt.rzi.s32.f32: NaN -> 0).
 * build & run:  gcc -O2 -fopenmp -o q4_lut_equiv q4_lut_equiv.c -lm && ./q4_lut_equiv   (~15 s on 8 cores) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static const float PIV_NF4[15] = {
    -0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f, -0.23460740596055984f,
    -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f, 0.1202552504837513f, 0.2035212516784668f,
    0.2920137718319893f, 0.3893125355243683f, 0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f,
};
static const float PIV_FP4[7] = {0.00260417f, 0.0859375f, 0.20833333f, 0.29166667f, 0.4166667f, 0.583333f, 0.8333333f};
static const unsigned CODE_FP4[8] = {0u, 1u, 6u, 7u, 4u, 5u, 2u, 3u};

/* ---- reference trees (as in blockwise.cu / q4_tree_equiv.c) */
static unsigned tree_nf4(float x) {
    unsigned c = 0;
    for (int i = 0; i < 15; ++i) c += (x > PIV_NF4[i]) ? 1u : 0u; /* == the nested tree: q4_tree_equiv.c */
    return c;
}
static unsigned tree_fp4(float x) {
    unsigned sign = (x < 0.0f) ? 8u : 0u;
    float a = fabsf(x);
    unsigned c;
    if (a > 0.29166667f) {
        if (a > 0.583333f) c = (a > 0.8333333f) ? 3u : 2u;
        else c = (a > 0.4166667f) ? 5u : 4u;
    } else {
        if (a > 0.0859375f) c = (a > 0.20833333f) ? 7u : 6u;
        else c = (a > 0.00260417f) ? 1u : 0u;
    }
    return c + sign;
}

/* ---- product: table build (build_q4_lut) */
typedef struct { float pivot; unsigned w; } entry;
static entry LUT_NF4[33], LUT_FP4[17];
static void build(void) {
    for (int t = 0; t < 33; ++t) {
        float lo = (float)t * 0.0625f - 1.0f, hi = lo + 0.0625f;
        int r = 0;
        for (int i = 0; i < 15; ++i) r += PIV_NF4[i] < lo ? 1 : 0;
        float pivot = INFINITY;
        unsigned below = (unsigned)r, above = below;
        if (r < 15 && PIV_NF4[r] < hi) { pivot = PIV_NF4[r]; above = below + 1u; }
        LUT_NF4[t].pivot = pivot;
        LUT_NF4[t].w = below | ((below ^ above) << 8);
    }
    for (int t = 0; t < 17; ++t) {
        float lo = (float)t * 0.0625f, hi = lo + 0.0625f;
        int r = 0;
        for (int i = 0; i < 7; ++i) r += PIV_FP4[i] < lo ? 1 : 0;
        float pivot = INFINITY;
        unsigned below = CODE_FP4[r], above = below;
        if (r < 7 && PIV_FP4[r] < hi) { pivot = PIV_FP4[r]; above = CODE_FP4[r + 1]; }
        LUT_FP4[t].pivot = pivot;
        LUT_FP4[t].w = below | ((below ^ above) << 8);
    }
}
static int cvt_rzi(float f) { return isnan(f) ? 0 : (int)f; } /* cvt.rzi.s32.f32 (in range here) */

/* ---- product: look-up (quantize4_lut); returns 0xff if the index would leave the table */
static unsigned lut_nf4(float x) {
    int t = cvt_rzi(fmaf(x, 16.0f, 16.0f));
    if (t < 0 || t > 32) return 0xffu;
    entry e = LUT_NF4[t];
    return (e.w & 0xffu) ^ ((x > e.pivot) ? (e.w >> 8) : 0u);
}
static unsigned lut_fp4(float x) {
    float a = fabsf(x);
    int t = cvt_rzi(a * 16.0f);
    if (t < 0 || t > 16) return 0xffu;
    entry e = LUT_FP4[t];
    return ((e.w & 0xffu) ^ ((a > e.pivot) ? (e.w >> 8) : 0u)) | ((x < 0.0f) ? 8u : 0u);
}

int main(void) {
    build();
    const float bound = 1.0f + 9.5367431640625e-07f; /* 1 + 2^-20 */
    unsigned long long bad_nf4 = 0, bad_fp4 = 0, checked = 0;
#pragma omp parallel for reduction(+ : bad_nf4, bad_fp4, checked) schedule(static)
    for (long long b = 0; b < (1LL << 32); ++b) {
        uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        if (!isnan(x) && fabsf(x) > bound) continue; /* cannot reach the kernel: |a * rcp(absmax)| <= 1 + 2^-22 */
        ++checked;
        bad_nf4 += lut_nf4(x) != tree_nf4(x);
        bad_fp4 += lut_fp4(x) != tree_fp4(x);
    }
    printf("checked %llu fp32 values (|x| <= 1 + 2^-20, and every NaN): NF4 mismatches %llu, FP4 mismatches %llu\n", checked,
           bad_nf4, bad_fp4);
    return (bad_nf4 || bad_fp4) ? 1 : 0;
}
