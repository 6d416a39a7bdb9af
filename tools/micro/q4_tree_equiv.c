/* q4_tree_equiv.c -- the 4-bit decision procedures of the product (bitsandbytes_b200/csrc/blockwise.cu:
 * quantize_nf4 / quantize_fp4, 4- and 3-level nested comparisons, restated here verbatim) return the
 * same code as the oracle's forms (oracle/oracle_c.c: NF4 = number of pivots below x, FP4 = the
 * reference's tree) for EVERY fp32 input, NaNs included.  So a GPU-vs-oracle code mismatch can only come
 * from the normalisation x = a * rcp(absmax), never from the tree.
 * build & run:  gcc -O2 -fopenmp -o q4_tree_equiv q4_tree_equiv.c -lm && ./q4_tree_equiv   (~20 s on 8 cores) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* ---- product forms (blockwise.cu) */
static unsigned product_nf4(float x) {
    unsigned c;
    if (x > 0.03979014977812767f) {
        if (x > 0.3893125355243683f) {
            if (x > 0.6427869200706482f)
                c = (x > 0.8614784181118011f) ? 15u : 14u;
            else
                c = (x > 0.5016634166240692f) ? 13u : 12u;
        } else {
            if (x > 0.2035212516784668f)
                c = (x > 0.2920137718319893f) ? 11u : 10u;
            else
                c = (x > 0.1202552504837513f) ? 9u : 8u;
        }
    } else {
        if (x > -0.33967943489551544f) {
            if (x > -0.13791173323988914f)
                c = (x > -0.045525018125772476f) ? 7u : 6u;
            else
                c = (x > -0.23460740596055984f) ? 5u : 4u;
        } else {
            if (x > -0.6106329262256622f)
                c = (x > -0.4599952697753906f) ? 3u : 2u;
            else
                c = (x > -0.8480964004993439f) ? 1u : 0u;
        }
    }
    return c;
}

static unsigned product_fp4(float x) {
    unsigned sign = (x < 0.0f) ? 8u : 0u;
    float a = fabsf(x);
    unsigned c;
    if (a > 0.29166667f) {
        if (a > 0.583333f)
            c = (a > 0.8333333f) ? 3u : 2u;
        else
            c = (a > 0.4166667f) ? 5u : 4u;
    } else {
        if (a > 0.0859375f)
            c = (a > 0.20833333f) ? 7u : 6u;
        else
            c = (a > 0.00260417f) ? 1u : 0u;
    }
    return c + sign;
}

/* ---- oracle forms (oracle_c.c) */
static const float NF4_PIVOT[15] = {
    -0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f, -0.23460740596055984f,
    -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f, 0.1202552504837513f, 0.2035212516784668f,
    0.2920137718319893f, 0.3893125355243683f, 0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f,
};
static unsigned oracle_nf4(float x) {
    unsigned c = 0;
    for (int i = 0; i < 15; ++i) c += (x > NF4_PIVOT[i]) ? 1u : 0u;
    return c;
}
static unsigned oracle_fp4(float x) {
    unsigned sign = (x < 0.0f) ? 8u : 0u;
    x = fabsf(x);
    if (x > 0.29166667f) {
        if (x > 0.583333f) {
            if (x > 0.8333333f) return 3u + sign;
            return 2u + sign;
        }
        if (x > 0.4166667f) return 5u + sign;
        return 4u + sign;
    }
    if (x > 0.0859375f) {
        if (x > 0.20833333f) return 7u + sign;
        return 6u + sign;
    }
    if (x > 0.00260417f) return 1u + sign;
    return 0u + sign;
}

int main(void) {
    long long bad_nf4 = 0, bad_fp4 = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_nf4, bad_fp4)
    for (long long bits = 0; bits < (1LL << 32); ++bits) {
        uint32_t u = (uint32_t)bits;
        float x;
        memcpy(&x, &u, 4);
        bad_nf4 += product_nf4(x) != oracle_nf4(x);
        bad_fp4 += product_fp4(x) != oracle_fp4(x);
    }
    printf("NF4: mismatches over all 2^32 inputs: %lld\nFP4: mismatches over all 2^32 inputs: %lld\n", bad_nf4, bad_fp4);
    return (bad_nf4 || bad_fp4) ? 1 : 0;
}
