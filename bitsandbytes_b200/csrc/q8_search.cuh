// q8_search.cuh -- nearest entry of a sorted 256-entry code book: the bracket-table form of the reference's 7-step
// search (reference csrc/kernels.cu:160-267), shared by the blockwise quantizer (blockwise.cu) and the 8-bit optimizer
// updates (optim.cu).
#pragma once

#include "common.cuh"

namespace bnb200 {

// The same codes with fewer instructions (the fast kernel's form).  For a sorted code book every comparison of the
// walk above is decided by c = #{j : code[j] < x}, so the walk's end state -- its last pivot and the neighbour it
// may still move to -- is a function of c alone: a 257-entry structural table (q8_structure) that does not depend on
// the code values; the final decision is the reference's own midpoint rule.  (Proved exhaustively on the CPU for all
// 2^32 inputs and six code books, duplicates included: tools/micro/q8_search_equiv.c.)
// c itself comes from a BRACKET table instead of a search: the value axis is cut into 772 cells by the float's own
// bits (sign, exponent, 4 mantissa bits; everything below 2^-24 in one cell per sign), T[t] = #{j : code[j] < low(t)}
// is built once per CTA (773 nine-probe searches), and for x in cell t   T[t] <= c <= T[t+1],   so a short linear
// scan (0..3 steps for the default dynamic map) finishes the count.  ~35 instructions per element instead of ~70.
// Equality with quantize_8bit for every fp32 input the kernel can produce: tools/micro/q8_lut_equiv.c.
// entry c of the structural table, pre-scaled to byte offsets into the code book:
// (4 * last pivot) | (4 * neighbour it may still move to) << 16
__device__ __forceinline__ uint32_t q8_structure(int c) {
    int pivot = 127, up = 255, lp = 0;
#pragma unroll
    for (int i = 64; i > 0; i >>= 1) {
        const bool gt = pivot < c;
        lp = gt ? pivot : lp;
        up = gt ? up : pivot;
        pivot += gt ? i : -i;
    }
    return (uint32_t)(4 * pivot) | ((uint32_t)(4 * (pivot < c ? up : lp)) << 16);
}

constexpr int kQ8MinKey = ((127 - 24) << 4) - 1;         // magnitude keys <= this share cell 0: |x| < 2^-24
constexpr int kQ8MagCells = (127 << 4) - kQ8MinKey + 1;   // 386: the last one is [1, 1.0625)
constexpr int kQ8Cells = 2 * kQ8MagCells;                 // negative cells (most negative first), then positive

// lower edge of magnitude cell cm (cm = 0: zero)
__device__ __forceinline__ float q8_mag_edge(int cm) {
    return cm == 0 ? 0.0f : __uint_as_float((uint32_t)(cm + kQ8MinKey) << 19);
}

// bracket entry t: T[t] | T[t + 1] << 16 with T[t] = #{j : code[j] < low(t)}; low(t) = the (exclusive, for negative
// cells: values (-edge(cm+1), -edge(cm)]) lower end of value cell t; T[kQ8Cells] counts against 1.0625
__device__ __forceinline__ uint32_t q8_count_below(const float* __restrict__ scode, float v) {
    unsigned c = 0;
#pragma unroll
    for (int s = 128; s > 0; s >>= 1) c += (scode[c + s - 1] < v) ? (unsigned)s : 0u;
    c += (c == 255u && scode[255] < v) ? 1u : 0u;
    return c;
}
__device__ __forceinline__ float q8_cell_low(int t) {
    if (t < kQ8MagCells) return -q8_mag_edge(kQ8MagCells - t);  // cm = kQ8MagCells - 1 - t, low = -edge(cm + 1)
    return q8_mag_edge(t - kQ8MagCells);
}
__device__ __forceinline__ void build_q8_bracket(const float* __restrict__ scode, uint32_t* __restrict__ sbr) {
    for (int t = threadIdx.x; t < kQ8Cells; t += blockDim.x)
        sbr[t] = q8_count_below(scode, q8_cell_low(t)) | (q8_count_below(scode, q8_cell_low(t + 1)) << 16);
}

// final decision table, entry c: { midpoint of the two candidate entries, candidate p | other o << 8 | (p < c) << 16 }
// (the reference's midpoint rule with the midpoint precomputed once per CTA instead of two code look-ups per element)
__device__ __forceinline__ void build_q8_final(const float* __restrict__ scode, float2* __restrict__ sfin) {
    for (int c = threadIdx.x; c <= 256; c += blockDim.x) {
        const uint32_t po = q8_structure(c);
        const unsigned p = (po & 0xffffu) >> 2, o = (po >> 16) >> 2;
        const float midpoint = mul_ftz(scode[o] + scode[p], 0.5f);
        sfin[c] = make_float2(midpoint, __uint_as_float(p | (o << 8) | ((p < (unsigned)c ? 1u : 0u) << 16)));
    }
}

__device__ __forceinline__ unsigned quantize_8bit_fast(const float* __restrict__ scode,
                                                       const float2* __restrict__ sfin,
                                                       const uint32_t* __restrict__ sbr, float x) {
    // value cell: magnitude key, everything tiny (and NaN, whose key is out of range) -> 0; zero counts as positive
    unsigned cm = ((__float_as_uint(x) & 0x7fffffffu) >> 19) - (unsigned)kQ8MinKey;
    cm = cm < (unsigned)kQ8MagCells ? cm : 0u;
    const unsigned t = (unsigned)kQ8MagCells + ((x < 0.0f) ? ~cm : cm);
    const uint32_t br = sbr[t];
    unsigned c = br & 0xffffu;
    const unsigned hi = br >> 16;
    while (c < hi && scode[c] < x) ++c;
    c = (x == x) ? c : 0u;  // NaN (0 * rcp(0)): no entry is below it
    const float2 f = sfin[c];
    const unsigned w = __float_as_uint(f.y);
    const bool move = (w >> 16) ? (x > f.x) : (x < f.x);
    return move ? ((w >> 8) & 0xffu) : (w & 0xffu);
}

} // namespace bnb200
