/* prmt_decode_equiv.c -- the register-table decode of the fused GEMM (bitsandbytes_b200/csrc/decode4.cuh:
 * build_table's byte planes + decode_word's PRMT network, restated with an exact emulation of the PRMT
 * instruction) returns, for EVERY packed 32-bit word and any 16-entry table of 16-bit values, exactly
 * table[code] for each of the eight codes, in k order (element 2b = high nibble of byte b).
 * build & run:  gcc -O2 -fopenmp -o prmt_decode_equiv prmt_decode_equiv.c && ./prmt_decode_equiv */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

/* PTX prmt.b32 (default mode) == CUDA __byte_perm: selector nibble n picks byte (n & 7) of {y:x};
 * bit 3 of the nibble replicates that byte's sign bit instead */
static uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xf;
        uint32_t b = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
        if (n & 8) b = (b & 0x80) ? 0xff : 0x00;
        r |= b << (8 * i);
    }
    return r;
}

typedef struct { uint32_t lo[4], hi[4]; } DecodeTable;

/* build_table after the roundings: pr[j] = entry(2j) | entry(2j+1) << 16 */
static void build_planes(const uint16_t* entry, DecodeTable* t) {
    uint32_t pr[8];
    for (int j = 0; j < 8; ++j) pr[j] = (uint32_t)entry[2 * j] | ((uint32_t)entry[2 * j + 1] << 16);
    for (int j = 0; j < 4; ++j) {
        t->lo[j] = byte_perm(pr[2 * j], pr[2 * j + 1], 0x6420);
        t->hi[j] = byte_perm(pr[2 * j], pr[2 * j + 1], 0x7531);
    }
}

static void decode_word(uint32_t w, const DecodeTable* t, uint32_t* o) {
    const uint32_t c7 = w & 0x77777777u;
    const uint32_t w1 = w >> 1;
    for (int g = 0; g < 2; ++g) {
        const uint32_t c = g ? (c7 >> 16) : c7;
        const uint32_t m = g ? (w1 >> 16) : w1;
        const uint32_t selm = (m & 0x4444u) | 0x3210u;
        const uint32_t lo = byte_perm(byte_perm(t->lo[0], t->lo[1], c), byte_perm(t->lo[2], t->lo[3], c), selm);
        const uint32_t hi = byte_perm(byte_perm(t->hi[0], t->hi[1], c), byte_perm(t->hi[2], t->hi[3], c), selm);
        o[2 * g] = byte_perm(lo, hi, 0x4051);
        o[2 * g + 1] = byte_perm(lo, hi, 0x6273);
    }
}

int main(void) {
    long long bad = 0;
    srand(11);
    for (int trial = 0; trial < 3; ++trial) {
        uint16_t entry[16];
        for (int i = 0; i < 16; ++i) entry[i] = (uint16_t)(rand() ^ (rand() << 9));  /* arbitrary bit patterns, sign bits set too */
        if (trial == 2) for (int i = 0; i < 16; ++i) entry[i] = (uint16_t)(0xff00 | i);  /* all high bytes negative */
        DecodeTable t;
        build_planes(entry, &t);
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (long long bits = 0; bits < (1LL << 32); ++bits) {
            const uint32_t w = (uint32_t)bits;
            uint32_t o[4];
            decode_word(w, &t, o);
            for (int b = 0; b < 4; ++b) {
                const uint32_t byte = (w >> (8 * b)) & 0xff;
                const uint32_t want = (uint32_t)entry[byte >> 4] | ((uint32_t)entry[byte & 15] << 16);
                bad += o[b] != want;
            }
        }
    }
    printf("PRMT decode: mismatches over 3 tables x all 2^32 packed words: %lld\n", bad);
    return bad != 0;
}
