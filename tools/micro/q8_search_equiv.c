/* q8_search_equiv.c -- exhaustive proof, on the CPU, that a cheaper 8-bit code search returns the SAME
 * code as the reference's dQuantize<0> (reference csrc/kernels.cu:160-219, restated in
 * bitsandbytes_b200/csrc/blockwise.cu::quantize_8bit) for EVERY fp32 input and any sorted code book.
 *
 * Idea: for a sorted table every comparison `x > code[j]` of the reference's 7-step walk is decided by
 * c = #{j : code[j] < x}, so the walk's end state (pivot, the neighbour it may still move to) is a
 * function of c alone -- a 257-entry structural table that does not depend on the code values.  The
 * fast version finds c with a plain 9-probe lower-bound search (3 instructions per probe instead of
 * 8) and finishes with the reference's own midpoint rule.
 *
 * build & run:  gcc -O2 -fopenmp -o q8_search_equiv q8_search_equiv.c -lm && ./q8_search_equiv
 * (checks all 2^32 bit patterns against several code books; a few minutes on 8 cores) */
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

/* structural table, exactly as the CUDA kernel builds it (blockwise.cu::q8_structure): for c = number of
 * entries below x, (4 * last pivot) | (4 * the neighbour it may still move to) << 16 */
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

/* blockwise.cu::quantize_8bit_fast, statement for statement (byte offsets c4 = 4 c) */
static unsigned fast_search(const float* code, float x) {
    const char* cb = (const char*)code;
    unsigned c4 = 0;
    for (int s = 128; s > 0; s >>= 1) c4 += (*(const float*)(cb + c4 + 4 * (s - 1)) < x) ? 4u * s : 0u;
    c4 += (c4 == 1020u && code[255] < x) ? 4u : 0u;
    const uint32_t po = *(const uint32_t*)((const char*)g_po + c4);
    const unsigned p4 = po & 0xffffu, o4 = po >> 16;
    const float midpoint = (*(const float*)(cb + o4) + *(const float*)(cb + p4)) * 0.5f;
    const int move = (p4 < c4) ? (x > midpoint) : (x < midpoint);
    return (move ? o4 : p4) >> 2;
}

static int cmpf(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

int main(void) {
    build_structure();
    static float books[6][256];
    static const float dynamic_map[256] = {
#include "q8_dynamic_map.inc"
    };
    const char* names[6] = {"dynamic-like (log spaced, signed)", "linear signed", "unsigned with zero padding",
                            "4-bit values zero-padded to 256", "random sorted with duplicates",
                            "create_dynamic_map() (the default 8-bit code)"};
    memcpy(books[5], dynamic_map, sizeof(dynamic_map));
    /* log-spaced signed */
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

    int bad_total = 0;
    for (int b = 0; b < 6; ++b) {
        long long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (long long bits = 0; bits < (1LL << 32); ++bits) {
            uint32_t u = (uint32_t)bits;
            float x;
            memcpy(&x, &u, 4);
            if (ref_search(books[b], x) != fast_search(books[b], x)) ++bad;
        }
        printf("%-45s mismatches over all 2^32 inputs: %lld\n", names[b], bad);
        bad_total += bad != 0;
    }
    return bad_total;
}
