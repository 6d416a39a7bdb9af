/* q8_lut_equiv.c -- exhaustive proof, on the CPU, that the BRACKET-TABLE form of the 8-bit code search in the
 * product's fast quantize kernel (bitsandbytes_b200/csrc/blockwise.cu: build_q8_bracket / quantize_8bit_fast)
 * returns the SAME code as the reference's dQuantize<0> walk (reference csrc/kernels.cu:160-219, restated in
 * blockwise.cu::quantize_8bit) for every fp32 input the kernel can produce -- x = a * rcp.approx(absmax) with
 * |a| <= absmax, i.e. |x| <= 1 + 2^-20 (generous), and NaN (all-zero block) -- and any sorted code book.
 *
 * The value axis is cut into 772 cells by the float's own bits (sign, exponent, 4 mantissa bits; everything below
 * 2^-24 shares one cell per sign); T[t] = #{j : code[j] < low(t)}; for x in cell t, T[t] <= c <= T[t+1] with
 * c = #{j : code[j] < x}; a linear scan finishes the count; the structural table + the reference's midpoint rule
 * (proved in q8_search_equiv.c) turn c into the code.  Everything below restates the CUDA code statement for
 * statement.
 * build & run:  gcc -O2 -fopenmp -o q8_lut_equiv q8_lut_equiv.c -lm && ./q8_lut_equiv   (about a minute on 8 cores) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static unsigned ref_search(const float* code, float x) {
    int pivot = 127, upper_pivot = 255, lower_pivot = 0;
    float lower = -1.0f, upper = 1.0f;
    float val = code[pivot];
    for (int i = 64; i > 0; i >>= 1) {
        if (x > val) {
            lower_pivot = pivot;
            lower = val;
            pivot += i;
        } else {
            upper_pivot = pivot;
            upper = val;
            pivot -= i;
        }
        val = code[pivot];
    }
    if (upper_pivot == 255) upper = code[255];
    if (lower_pivot == 0) lower = code[0];
    if (x > val) {
        float midpoint = (upper + val) * 0.5f;
        return x > midpoint ? (unsigned)upper_pivot : (unsigned)pivot;
    } else {
        float midpoint = (lower + val) * 0.5f;
        return x < midpoint ? (unsigned)lower_pivot : (unsigned)pivot;
    }
}

static uint32_t g_po[257];
static void build_structure(void) {
    for (int c = 0; c <= 256; ++c) {
        int pivot = 127, up = 255, lp = 0;
        for (int i = 64; i > 0; i >>= 1) {
            const int gt = pivot < c;
            lp = gt ? pivot : lp;
            up = gt ? up : pivot;
            pivot += gt ? i : -i;
        }
        g_po[c] = (uint32_t)(4 * pivot) | ((uint32_t)(4 * (pivot < c ? up : lp)) << 16);
    }
}

#define Q8_MIN_KEY (((127 - 24) << 4) - 1)
#define Q8_MAG_CELLS ((127 << 4) - Q8_MIN_KEY + 1)
#define Q8_CELLS (2 * Q8_MAG_CELLS)

static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float mag_edge(int cm) { return cm == 0 ? 0.0f : from_bits((uint32_t)(cm + Q8_MIN_KEY) << 19); }
static uint32_t count_below(const float* code, float v) {
    unsigned c = 0;
    for (int s = 128; s > 0; s >>= 1) c += (code[c + s - 1] < v) ? (unsigned)s : 0u;
    c += (c == 255u && code[255] < v) ? 1u : 0u;
    return c;
}
static float cell_low(int t) {
    if (t < Q8_MAG_CELLS) return -mag_edge(Q8_MAG_CELLS - t);
    return mag_edge(t - Q8_MAG_CELLS);
}
static void build_bracket(const float* code, uint32_t* br) {
    for (int t = 0; t < Q8_CELLS; ++t) br[t] = count_below(code, cell_low(t)) | (count_below(code, cell_low(t + 1)) << 16);
}
/* build_q8_final: entry c = { midpoint of the two candidates, p | o << 8 | (p < c) << 16 } */
typedef struct { float mid; uint32_t w; } fin_t;
static void build_final(const float* code, fin_t* fin) {
    for (int c = 0; c <= 256; ++c) {
        const uint32_t po = g_po[c];
        const unsigned p = (po & 0xffffu) >> 2, o = (po >> 16) >> 2;
        fin[c].mid = (code[o] + code[p]) * 0.5f;
        fin[c].w = p | (o << 8) | ((p < (unsigned)c ? 1u : 0u) << 16);
    }
}
static fin_t g_fin[8][257];
static int g_book = 0;
#pragma omp threadprivate(g_book)

static unsigned lut_search(const float* code, const uint32_t* br, float x) {
    unsigned cm = ((to_bits(x) & 0x7fffffffu) >> 19) - (unsigned)Q8_MIN_KEY;
    cm = cm < (unsigned)Q8_MAG_CELLS ? cm : 0u;
    const unsigned t = (unsigned)Q8_MAG_CELLS + ((x < 0.0f) ? ~cm : cm);
    const uint32_t b = br[t];
    unsigned c = b & 0xffffu;
    const unsigned hi = b >> 16;
    while (c < hi && code[c] < x) ++c;
    c = (x == x) ? c : 0u;
    const fin_t f = g_fin[g_book][c];
    const int move = (f.w >> 16) ? (x > f.mid) : (x < f.mid);
    return move ? ((f.w >> 8) & 0xffu) : (f.w & 0xffu);
}

static int cmpf(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

int main(void) {
    build_structure();
    static float books[8][256];
    static uint32_t br[8][Q8_CELLS];
    static const float dynamic_map[256] = {
#include "q8_dynamic_map.inc"
    };
    static const float udynamic_map[256] = {
#include "q8_udynamic_map.inc"
    };
    const char* names[8] = {"dynamic-like (log spaced, signed)", "linear signed", "unsigned with zero padding",
                            "4-bit values zero-padded to 256", "random sorted with duplicates",
                            "create_dynamic_map() (the default 8-bit code)", "tiny magnitudes + a negative zero",
                            "create_dynamic_map(signed=False) (second optimizer state)"};
    memcpy(books[5], dynamic_map, sizeof(dynamic_map));
    memcpy(books[7], udynamic_map, sizeof(udynamic_map));
    for (int i = 0; i < 127; ++i) {
        float v = powf(10.0f, -7.0f * (float)(126 - i) / 126.0f);
        books[0][129 + i] = v;
        books[0][126 - i] = -v;
    }
    books[0][127] = -1e-9f; books[0][128] = 0.0f;
    qsort(books[0], 256, sizeof(float), cmpf);
    for (int i = 0; i < 256; ++i) books[1][i] = -1.0f + 2.0f * (float)i / 255.0f;
    for (int i = 0; i < 256; ++i) books[2][i] = i < 128 ? 0.0f : (float)(i - 127) / 128.0f;
    for (int i = 0; i < 256; ++i) books[3][i] = 0.0f;
    for (int i = 0; i < 8; ++i) { books[3][i] = -1.0f + (float)i / 8.0f; books[3][255 - i] = 1.0f - (float)i / 8.0f; }
    qsort(books[3], 256, sizeof(float), cmpf);
    srand(7);
    for (int i = 0; i < 256; ++i) books[4][i] = (float)(rand() % 97) / 48.0f - 1.0f;
    qsort(books[4], 256, sizeof(float), cmpf);
    /* many entries inside the two "tiny" cells, denormals, and -0.0 next to +0.0 */
    for (int i = 0; i < 256; ++i) books[6][i] = ((float)i - 127.5f) * 1e-10f;
    books[6][0] = -1.0f; books[6][1] = -0.5f; books[6][254] = 0.5f; books[6][255] = 1.0f;
    books[6][100] = -1e-42f; books[6][101] = -0.0f; books[6][102] = 0.0f; books[6][103] = 1e-42f;
    qsort(books[6], 256, sizeof(float), cmpf);
    for (int b = 0; b < 8; ++b) { build_bracket(books[b], br[b]); build_final(books[b], g_fin[b]); }

    const float bound = 1.0f + 9.5367431640625e-07f; /* 1 + 2^-20 */
    int bad_total = 0;
    for (int b = 0; b < 8; ++b) {
        long long bad = 0, checked = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad, checked)
        for (long long bits = 0; bits < (1LL << 32); ++bits) {
            g_book = b;
            float x = from_bits((uint32_t)bits);
            if (!isnan(x) && fabsf(x) > bound) continue;
            ++checked;
            if (ref_search(books[b], x) != lut_search(books[b], br[b], x)) ++bad;
        }
        printf("%-45s mismatches over %lld inputs (|x| <= 1 + 2^-20, NaN): %lld\n", names[b], checked, bad);
        bad_total += bad != 0;
    }
    return bad_total;
}
