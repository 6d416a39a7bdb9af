/*
 * oracle_c.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU-only restatement of the reference's (bitsandbytes @ 95f9af3)
 * algorithms for the hot path named in BASELINE.json: blockwise quantize /
 * dequantize (8-bit dynamic map, NF4, FP4), the 4-bit dequant-fused GEMM
 * contract, and the LLM.int8() row quantisation / int8 GEMM / dequant epilogue.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's shared object.  The product
 * (bitsandbytes_b200/) never links, imports or falls back to it.
 *
 * Every function cites the reference file:line it restates.  The restatement
 * follows the reference **CUDA** kernels (not the CPU backend's 64K-LUT
 * approximation, csrc/cpu_ops.cpp:496-572), because GPU parity is defined
 * against the CUDA semantics (SURVEY.md key fact 5).
 *
 * Parity pinning (see tests/test_oracle_pinned.py and tests/golden/):
 *   - absmax, dequantize (8-bit, NF4, FP4 -> fp32/bf16/fp16): bit-exact against
 *     the reference CPU library built from the reference sources
 *     (oracle/_ref/libbitsandbytes_cpu.so) and against golden vectors generated
 *     by importing the reference Python package (tests/golden/make_golden.py).
 *   - quantize codes: the reference CUDA kernels normalise with a *fast-math*
 *     reciprocal (MUFU.RCP, `1.0f / absmax` under --use_fast_math,
 *     CMakeLists.txt:190, kernels.cu:332) that cannot be restated bit-exactly in
 *     C.  This oracle uses the IEEE reciprocal, so a code may differ from the GPU
 *     on inputs within ~2 ulp of a decision threshold.  The GPU tests therefore
 *     (a) demand bit-exactness against the reference CUDA library itself
 *     (oracle/_ref/libbitsandbytes_cuda_ref.so, same C ABI, same device buffers)
 *     and (b) against this oracle allow only such threshold-adjacent mismatches.
 *   - Denormal flushing (-ftz=true under fast-math) is not modelled; tests keep
 *     absmax in the normal range.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Code books.  csrc/kernels.cu:15-43, csrc/gemm_4bit_common.cuh:17-54  */
/* ------------------------------------------------------------------ */
static const float NF4_LUT[16] = {
    -1.0f,
    -0.6961928009986877f,
    -0.5250730514526367f,
    -0.39491748809814453f,
    -0.28444138169288635f,
    -0.18477343022823334f,
    -0.09105003625154495f,
    0.0f,
    0.07958029955625534f,
    0.16093020141124725f,
    0.24611230194568634f,
    0.33791524171829224f,
    0.44070982933044434f,
    0.5626170039176941f,
    0.7229568362236023f,
    1.0f,
};

/* magnitude table indexed by the low 3 bits; bit 3 is the sign (kernels.cu:15-24, 59-62) */
static const float FP4_MAG[8] = {
    0.0f, 0.005208333333f, 0.66666667f, 1.0f, 0.33333333f, 0.5f, 0.16666667f, 0.25f,
};

static inline float fp4_value(unsigned q) {
    /* kernels.cu:59-62: lut[val & 7] * (1 - 2*sign) */
    float sign = 1.0f - 2.0f * (float)((q >> 3) & 1u);
    return FP4_MAG[q & 7u] * sign;
}

static inline float code4_value(unsigned q, int quant_type) {
    return quant_type == 2 ? NF4_LUT[q & 15u] : fp4_value(q & 15u);
}

void oracle_get_4bit_lut(float* out16, int quant_type) {
    for (unsigned i = 0; i < 16; ++i) out16[i] = code4_value(i, quant_type);
}

/* ------------------------------------------------------------------ */
/* value -> code decision procedures                                   */
/* ------------------------------------------------------------------ */

/* kernels.cu:110-153 (pivots = midpoints of NF4_LUT, strict '>', ties -> lower code) */
static const float NF4_PIVOT[15] = {
    -0.8480964004993439f,   /* 0|1  */
    -0.6106329262256622f,   /* 1|2  */
    -0.4599952697753906f,   /* 2|3  */
    -0.33967943489551544f,  /* 3|4  */
    -0.23460740596055984f,  /* 4|5  */
    -0.13791173323988914f,  /* 5|6  */
    -0.045525018125772476f, /* 6|7  */
    0.03979014977812767f,   /* 7|8  */
    0.1202552504837513f,    /* 8|9  */
    0.2035212516784668f,    /* 9|10 */
    0.2920137718319893f,    /* 10|11 */
    0.3893125355243683f,    /* 11|12 */
    0.5016634166240692f,    /* 12|13 */
    0.6427869200706482f,    /* 13|14 */
    0.8614784181118011f,    /* 14|15 */
};

static inline unsigned quantize_nf4(float x) {
    /* The reference tree is a balanced binary search over the 15 pivots with strict '>'.
     * Equivalent: code = number of pivots p with x > p.  NaN compares false -> code 0,
     * which is what the tree returns as well (every branch falls to the 'else'). */
    unsigned c = 0;
    for (int i = 0; i < 15; ++i) c += (x > NF4_PIVOT[i]) ? 1u : 0u;
    return c;
}

/* kernels.cu:64-106 */
static inline unsigned quantize_fp4(float x) {
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

/* kernels.cu:160-219 (dQuantize<0>): 7-step binary search from pivot 127 over the
 * 256-entry code, then a midpoint comparison against the bracketing neighbour. */
static inline unsigned quantize_8bit(const float* code, float x) {
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
    if (upper_pivot == 255) upper = code[upper_pivot];
    if (lower_pivot == 0) lower = code[lower_pivot];
    if (x > val) {
        float midpoint = (upper + val) * 0.5f;
        return (x > midpoint) ? (unsigned)upper_pivot : (unsigned)pivot;
    } else {
        float midpoint = (lower + val) * 0.5f;
        return (x < midpoint) ? (unsigned)lower_pivot : (unsigned)pivot;
    }
}

/* ------------------------------------------------------------------ */
/* quantize_blockwise.  kernels.cu:269-375 (+ Small variant :388-463),  */
/* launch ops.cu:36-75.  A is given as fp32 (bf16/fp16 inputs are       */
/* widened exactly by the caller, as `(float)vals[j]` does).            */
/* quant_type: 0 = General8bit, 1 = FP4, 2 = NF4 (csrc/common.h:3-7).   */
/* 4-bit: byte b holds element 2b in the HIGH nibble, 2b+1 in the LOW   */
/* nibble; elements past n read as 0.0f (kernels.cu:357-358, :317).     */
/* ------------------------------------------------------------------ */
void oracle_quantize_blockwise(
    const float* code, const float* A, float* absmax, uint8_t* out, long blocksize, long n, int quant_type
) {
    long nblocks = (n + blocksize - 1) / blocksize;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nblocks; ++b) {
        long lo = b * blocksize;
        long hi = lo + blocksize < n ? lo + blocksize : n;
        float m = -3.402823466e+38f; /* -FLT_MAX, kernels.cu:316 */
        for (long i = lo; i < hi; ++i) m = fmaxf(m, fabsf(A[i]));
        absmax[b] = m;
        /* GPU: rcp.approx.ftz (fast-math); here IEEE.  See header. */
        float inv = 1.0f / m;
        if (quant_type == 0) {
            for (long i = lo; i < hi; ++i) out[i] = (uint8_t)quantize_8bit(code, A[i] * inv);
        } else {
            long padded_hi = lo + ((hi - lo + 1) / 2) * 2;
            for (long i = lo; i < padded_hi; i += 2) {
                float x0 = A[i] * inv;
                float x1 = (i + 1 < hi ? A[i + 1] : 0.0f) * inv;
                unsigned q0 = quant_type == 2 ? quantize_nf4(x0) : quantize_fp4(x0);
                unsigned q1 = quant_type == 2 ? quantize_nf4(x1) : quantize_fp4(x1);
                out[i / 2] = (uint8_t)((q0 << 4) | q1);
            }
        }
    }
}

/* For the threshold-adjacency check in tests: returns the normalised value the
 * decision procedure saw and the distance (in the normalised domain) to the
 * nearest decision threshold of the 4-bit trees. */
float oracle_nf4_threshold_distance(float x) {
    float best = 1e30f;
    for (int i = 0; i < 15; ++i) {
        float d = fabsf(x - NF4_PIVOT[i]);
        if (d < best) best = d;
    }
    return best;
}

float oracle_fp4_threshold_distance(float x) {
    static const float P[7] = {0.29166667f, 0.583333f, 0.8333333f, 0.4166667f, 0.0859375f, 0.20833333f, 0.00260417f};
    float ax = fabsf(x), best = fabsf(x); /* sign threshold at 0 */
    for (int i = 0; i < 7; ++i) {
        float d = fabsf(ax - P[i]);
        if (d < best) best = d;
    }
    return best;
}

/* ------------------------------------------------------------------ */
/* dequantize_blockwise.  kernels.cu:465-529, launch ops.cu:77-94.      */
/* out = T(value(code) * absmax[block]); product in fp32, ONE rounding. */
/* This function returns the fp32 products; oracle_round_* apply T().   */
/* n = number of OUTPUT elements.                                       */
/* ------------------------------------------------------------------ */
void oracle_dequantize_blockwise(
    const float* code, const uint8_t* A, const float* absmax, float* out, long blocksize, long n, int quant_type
) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        float s = absmax[i / blocksize];
        float v;
        if (quant_type == 0) {
            v = code[A[i]];
        } else {
            uint8_t byte = A[i >> 1];
            unsigned q = (i & 1) ? (byte & 0x0Fu) : (byte >> 4);
            v = code4_value(q, quant_type);
        }
        out[i] = v * s;
    }
}

/* fp32 -> bf16, round-to-nearest-even (what __float2bfloat16_rn / cvt.rn.bf16.f32 do). */
static inline uint16_t f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)0x7fffu; /* NaN */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}

static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline uint16_t f32_to_f16(float f) {
    _Float16 h = (_Float16)f; /* IEEE RNE */
    uint16_t r;
    memcpy(&r, &h, 2);
    return r;
}

static inline float f16_to_f32(uint16_t h) {
    _Float16 x;
    memcpy(&x, &h, 2);
    return (float)x;
}

void oracle_round_bf16(const float* in, uint16_t* out, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) out[i] = f32_to_bf16(in[i]);
}

void oracle_round_fp16(const float* in, uint16_t* out, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) out[i] = f32_to_f16(in[i]);
}

void oracle_widen_bf16(const uint16_t* in, float* out, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) out[i] = bf16_to_f32(in[i]);
}

void oracle_widen_fp16(const uint16_t* in, float* out, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) out[i] = f16_to_f32(in[i]);
}

/* ------------------------------------------------------------------ */
/* gemm_4bit contract.  _ops.py:239-295; backends/cuda/ops.py:904-916   */
/* (dequant + F.linear) and gemm_4bit_sm80.cu:220,292-307 /             */
/* gemm_4bit_mma.cuh:99-101 (fused MMA):                                */
/*   scale[i] = nested ? code2[absmax8[i]] * absmax[i >> 8] + offset    */
/*                     : absmax[i]                       (fp32)         */
/*   W_T[n,k] = rn_T(value(q[n,k]) * scale[(n*K+k)/blocksize])          */
/*   out[m,n] = T( sum_k A[m,k] * W_T[n,k]  + bias[n] )                 */
/* The reference accumulates in fp32 in an unspecified order (tensor    */
/* cores / cuBLAS); this oracle accumulates in double and returns the   */
/* unrounded sum as fp32 + as double so tests can bound the difference. */
/* wdtype: 0 = fp32 (no rounding of W), 1 = bf16, 2 = fp16.             */
/* A must already be widened to fp32 (exact).                           */
/* ------------------------------------------------------------------ */
void oracle_gemm_4bit(
    const float* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit, const float* absmax_code,
    const float* absmax_offset, double* out, const float* bias, long M, long N, long K, long blocksize, int quant_type,
    int wdtype
) {
    float offset = absmax_offset ? absmax_offset[0] : 0.0f;
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (long n = 0; n < N; ++n) {
            /* dequantise row n once */
            float* wrow = (float*)__builtin_malloc((size_t)K * sizeof(float));
            for (long k = 0; k < K; ++k) {
                long e = n * K + k;
                long bi = e / blocksize;
                float s;
                if (absmax_8bit) {
                    s = absmax_code[absmax_8bit[bi]] * absmax[bi >> 8] + offset;
                } else {
                    s = absmax[bi];
                }
                uint8_t byte = B[e >> 1];
                unsigned q = (e & 1) ? (byte & 0x0Fu) : (byte >> 4);
                float v = code4_value(q, quant_type) * s;
                if (wdtype == 1) v = bf16_to_f32(f32_to_bf16(v));
                if (wdtype == 2) v = f16_to_f32(f32_to_f16(v));
                wrow[k] = v;
            }
            for (long m = 0; m < M; ++m) {
                const float* a = A + m * K;
                double acc = 0.0;
                for (long k = 0; k < K; ++k) acc += (double)a[k] * (double)wrow[k];
                if (bias) acc += (double)bias[n];
                out[m * N + n] = acc;
            }
            __builtin_free(wrow);
        }
    }
}

/* Nested absmax reconstruction alone (functional.py:746-750; default/ops.py:336-341). */
void oracle_nested_absmax(
    const float* absmax2, const uint8_t* absmax_8bit, const float* code2, float offset, float* out, long nblocks
) {
    for (long i = 0; i < nblocks; ++i) out[i] = code2[absmax_8bit[i]] * absmax2[i >> 8] + offset;
}

/* ------------------------------------------------------------------ */
/* LLM.int8() pieces                                                    */
/* ------------------------------------------------------------------ */

/* kernels.cu:1331-1385 (kInt8VectorQuant<half,1024,SPARSE>), launch ops.cu:424-433.
 * A is fp16 [rows, cols].  Row absmax is computed on fp16 values (exact max),
 * excluding |a| >= threshold when threshold > 0; q = rint(a * (127 / absmax)),
 * outliers -> 0.  GPU uses __fdividef (approximate); here IEEE division: a code
 * may differ by +-1 only when a*scale is within ~2 ulp of a .5 boundary. */
void oracle_int8_vector_quant(
    const uint16_t* A, int8_t* out, float* rowStats, float threshold, long rows, long cols
) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        const uint16_t* row = A + r * cols;
        /* T(-FLT_MIN) with T = half is -0.0 */
        float m = -0.0f;
        float thr_h = f16_to_f32(f32_to_f16(threshold)); /* T(threshold) */
        for (long c = 0; c < cols; ++c) {
            float a = fabsf(f16_to_f32(row[c]));
            if (threshold > 0.0f) {
                if (a < thr_h) m = fmaxf(m, a);
            } else {
                m = fmaxf(m, a);
            }
        }
        rowStats[r] = m;
        float scale = 127.0f / m;
        for (long c = 0; c < cols; ++c) {
            float v = f16_to_f32(row[c]);
            int q;
            if (threshold > 0.0f && !(fabsf(v) < threshold)) {
                q = 0;
            } else {
                q = (int)nearbyintf(v * scale);
            }
            out[r * cols + c] = (int8_t)q;
        }
    }
}

/* ops.cu:282-404 (igemmlt<32,0> -> cublasLtMatmul): exact int32 = sum int8*int8.
 * A: activations [M,K], B: weights [N,K] (both row-major), C [M,N]. */
void oracle_int8_gemm(const int8_t* A, const int8_t* B, int32_t* C, long M, long N, long K) {
#pragma omp parallel for schedule(static)
    for (long m = 0; m < M; ++m) {
        for (long n = 0; n < N; ++n) {
            int32_t acc = 0;
            const int8_t* a = A + m * K;
            const int8_t* b = B + n * K;
            for (long k = 0; k < K; ++k) acc += (int32_t)a[k] * (int32_t)b[k];
            C[m * N + n] = acc;
        }
    }
}

/* kernels.cu:1394-1448 (kdequant_mm_int32_fp16):
 *   out = __float2half( fmaf( float(i32) * rowStats[r] * colStats[c], 6.200012e-05f, bias[c] ) ) */
void oracle_int8_mm_dequant(
    const int32_t* A, const float* rowStats, const float* colStats, uint16_t* out, const uint16_t* bias, long rows,
    long cols
) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        for (long c = 0; c < cols; ++c) {
            float b = bias ? f16_to_f32(bias[c]) : 0.0f;
            float t = (float)A[r * cols + c] * rowStats[r] * colStats[c];
            out[r * cols + c] = f32_to_f16(fmaf(t, 6.200012e-05f, b));
        }
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
