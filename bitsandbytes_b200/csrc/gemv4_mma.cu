// gemv4_mma.cu -- decode-time 4-bit GEMV / skinny GEMM (M <= 16 tokens) for sm_100a.
//
// The regime the reference serves with gemm_4bit_simt (reference csrc/gemm_4bit_simt.cu:109-480,
// dispatch bitsandbytes/backends/cuda/ops.py:583-623): one to a few tokens against the whole packed
// weight.  The roofline is the HBM stream of the codes (N*K/2 bytes); what actually limits a
// CUDA-core kernel is the instruction count per weight (decode + widen + FFMA).  Here the multiply-
// accumulate goes to the tensor cores through the warp-level mma.sync.m16n8k16 (a 128 x MT tcgen05
// tile would be 87 % padding at M = 8), so the CUDA cores only decode:
//
//   * a CTA owns 16 output features (the MMA's M); its 4 or 8 warps split K in 256-wide chunks and
//     meet in shared memory at the end;
//   * lane (g = lane / 4, t = lane % 4) owns k in [64 t, 64 t + 64) of the chunk for rows g and
//     g + 8: two 16-byte code loads per row (a warp reads 16 rows x 128 contiguous bytes), one
//     quantisation block per row, so one register table (decode4.cuh) per 64 weights;
//   * the sum over k does not care about order, so the four codes a lane feeds to one MMA are four
//     CONSECUTIVE k of its own range (k-slots 2t, 2t+1, 2t+8, 2t+9 of the instruction), and the
//     activation fragment is loaded from the same addresses: no shuffles, no shared-memory staging;
//   * tokens are the MMA's N (8 per instruction); columns beyond M are zero fragments.
//
// Numerics: W_T = rn_T(value * scale) exactly as everywhere else (decode4.cuh), products exact,
// fp32 accumulation in the tensor core, bias added in fp32, one rounding to T.
#include "common.cuh"
#include "decode4.cuh"

#include <cstdlib>

namespace bnb200 {

namespace {

constexpr int kGRows = 16;    // output features per CTA
constexpr int kGChunk = 256;  // k per warp iteration

template <typename T> struct WarpMma;
template <> struct WarpMma<__nv_bfloat16> {
    static __device__ __forceinline__ void run(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                     "{%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};
template <> struct WarpMma<__half> {
    static __device__ __forceinline__ void run(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                     "{%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

// W = warps per CTA (they split K).  Codes and scales of the next chunk are prefetched into registers;
// activations come through L1 (every CTA on an SM reads the same M x K slice).  Measured and dropped:
// staging the activations with cp.async (slower, it bypasses L1), issuing their loads before the
// decode, 3 or 4 CTAs per SM through a register cap -- all within noise of this version.
// NT = groups of 8 tokens (1: M <= 8, 2: M <= 16): the decoded weight fragments feed NT MMAs each.
template <typename T, int QT, int W, int NT>
__global__ void __launch_bounds__(W * 32, 2)
    gemv4_mma_kernel(const T* __restrict__ A, const uint8_t* __restrict__ B, const float* absmax,
                     const uint8_t* absmax_8bit, const float* absmax_code, const float* absmax_offset,
                     T* __restrict__ out, const T* __restrict__ bias, int M, int N, int K, int ldc, int log2_bs) {
    __shared__ float red[W][kGRows * 8 * NT];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2;
    const int t = lane & 3;
    const int n0 = blockIdx.x * kGRows;
    ScaleSrc sc{absmax, absmax_8bit, absmax_code,
                (absmax_8bit != nullptr && absmax_offset != nullptr) ? __ldg(absmax_offset) : 0.f};
    const bool two_scales = log2_bs == 5;  // blocksize 32: two quantisation blocks per 64 codes
    bool tok_ok[NT];                       // this lane's tokens (MMA column g of token group u)
    const T* arow[NT];
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        tok_ok[u] = g + 8 * u < M;
        arow[u] = A + (long long)(tok_ok[u] ? g + 8 * u : 0) * K;
    }

    // this lane's two weight rows (rows past N contribute zero fragments)
    long long e_row[2];
    bool row_ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n = n0 + g + 8 * h;
        row_ok[h] = n < N;
        e_row[h] = (long long)(row_ok[h] ? n : 0) * K;
    }

    float c[NT][4];
#pragma unroll
    for (int u = 0; u < NT; ++u) c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0.f;
    const int nchunks = (K + kGChunk - 1) / kGChunk;

    uint4 q[2][2];
    float s[2][2];
    auto fetch = [&](int ch) {
        const int kb = ch * kGChunk + 64 * t;
        const bool live = ch < nchunks && kb < K;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            q[h][0] = q[h][1] = make_uint4(0, 0, 0, 0);
            s[h][0] = s[h][1] = 0.f;
            if (live && row_ok[h]) {
                const long long e = e_row[h] + kb;
                const uint8_t* src = B + (e >> 1);
                q[h][0] = ldg_stream_v4(src);
                q[h][1] = ldg_stream_v4(src + 16);
                s[h][0] = sc.load(e >> log2_bs);
                s[h][1] = two_scales ? sc.load((e + 32) >> log2_bs) : s[h][0];
            }
        }
    };
    fetch(warp);

    for (int ch = warp; ch < nchunks; ch += W) {
        const int kb = ch * kGChunk + 64 * t;
        const bool k_ok = kb < K;
        const uint4 q00 = q[0][0], q01 = q[0][1], q10 = q[1][0], q11 = q[1][1];
        const float s00 = s[0][0], s01 = s[0][1], s10 = s[1][0], s11 = s[1][1];
        fetch(ch + W);

        DecodeTable tab0, tab1;
        build_table<T, QT>(s00, tab0);
        build_table<T, QT>(s10, tab1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            if (hh == 1 && two_scales) {
                build_table<T, QT>(s01, tab0);
                build_table<T, QT>(s11, tab1);
            }
            const uint4 qa = hh ? q01 : q00;  // row g:     32 codes = k [kb + 32 hh, +32)
            const uint4 qb = hh ? q11 : q10;  // row g + 8
            uint32_t ra[16], rb[16];
            decode_word(qa.x, tab0, ra + 0);
            decode_word(qa.y, tab0, ra + 4);
            decode_word(qa.z, tab0, ra + 8);
            decode_word(qa.w, tab0, ra + 12);
            decode_word(qb.x, tab1, rb + 0);
            decode_word(qb.y, tab1, rb + 4);
            decode_word(qb.z, tab1, rb + 8);
            decode_word(qb.w, tab1, rb + 12);
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                // activations of token g + 8u at the same 32 k: 16 pairs
                const uint4* xp = reinterpret_cast<const uint4*>(arow[u] + kb);
                uint32_t xw[16];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    uint4 av = make_uint4(0, 0, 0, 0);
                    if (tok_ok[u] && k_ok) av = __ldg(xp + 4 * hh + v);
                    xw[4 * v + 0] = av.x;
                    xw[4 * v + 1] = av.y;
                    xw[4 * v + 2] = av.z;
                    xw[4 * v + 3] = av.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    WarpMma<T>::run(c[u], ra[2 * j], rb[2 * j], ra[2 * j + 1], rb[2 * j + 1], xw[2 * j], xw[2 * j + 1]);
            }
        }
    }

    // accumulator fragment: c0/c1 = (row g, tokens 8u + 2t, 8u + 2t+1), c2/c3 = (row g + 8, same tokens)
#pragma unroll
    for (int u = 0; u < NT; ++u) {
        float* r = red[warp] + u * 8 * kGRows;
        r[(2 * t) * kGRows + g] = c[u][0];
        r[(2 * t + 1) * kGRows + g] = c[u][1];
        r[(2 * t) * kGRows + g + 8] = c[u][2];
        r[(2 * t + 1) * kGRows + g + 8] = c[u][3];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kGRows * 8 * NT; idx += W * 32) {
        const int tok = idx / kGRows;  // idx = token * 16 + row
        const int n = n0 + (idx % kGRows);
        if (tok < M && n < N) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < W; ++w) acc += red[w][idx];
            const float b = bias != nullptr ? DT<T>::to_f32(bias[n]) : 0.f;
            out[(long long)tok * ldc + n] = DT<T>::from_f32(acc + b);
        }
    }
}

} // namespace

// M <= 16, 16-bit activations, K % 64 == 0, power-of-two blocksize >= 32, 16-byte aligned A and B.
template <typename T>
bool launch_gemv4_mma(const T* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit,
                      const float* absmax_code, const float* absmax_offset, T* out, const T* bias, int M, int N, int K,
                      int ldc, int blocksize, int quant_type, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return true;
    if (M > 16 || K < 64 || (K % 64) != 0) return false;
    if (blocksize < 32 || (blocksize & (blocksize - 1)) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0) return false;
    if (quant_type != kNF4 && quant_type != kFP4) return false;
    const dim3 grid((N + kGRows - 1) / kGRows);
    const int l2 = ilog2_pow2(blocksize);
    // Few row tiles (small N): 8 warps per CTA split K so that the machine is covered; otherwise 4
    // warps, which run more k-chunks each and keep the prefetch pipeline busy.
    static const int forced_w = [] {
        const char* e = getenv("BNB_B200_GEMV_WARPS");
        return e ? atoi(e) : 0;
    }();
    // Narrow layers (N <= 2048: fewer than one CTA per SM at 8 warps): 16 warps per CTA split K into single
    // 256-wide chunks so that the whole packed weight is in flight at once (measured, 1024 x 4096, M = 1..8:
    // 5.9-6.6 us against 6.7-7.2; wider layers lose 5-15 % with 16 warps, profiles/r02_decode_regime.md).
    int warps = ((long long)grid.x * 4 >= 12LL * device_sm_count()) ? 4 : 8;
    if (K >= 16 * kGChunk && (int)grid.x <= 128) warps = 16;
    if (forced_w == 4 || forced_w == 8 || forced_w == 16) warps = forced_w;
#define BNB200_GEMV_MMA(QT, WV)                                                                                        \
    do {                                                                                                               \
        if (M <= 8)                                                                                                    \
            gemv4_mma_kernel<T, QT, WV, 1><<<grid, WV * 32, 0, stream>>>(A, B, absmax, absmax_8bit, absmax_code,      \
                                                                         absmax_offset, out, bias, M, N, K, ldc, l2); \
        else                                                                                                           \
            gemv4_mma_kernel<T, QT, WV, 2><<<grid, WV * 32, 0, stream>>>(A, B, absmax, absmax_8bit, absmax_code,      \
                                                                         absmax_offset, out, bias, M, N, K, ldc, l2); \
    } while (0)
    if (quant_type == kNF4) {
        if (warps == 4) BNB200_GEMV_MMA(kNF4, 4);
        else if (warps == 8) BNB200_GEMV_MMA(kNF4, 8);
        else BNB200_GEMV_MMA(kNF4, 16);
    } else {
        if (warps == 4) BNB200_GEMV_MMA(kFP4, 4);
        else if (warps == 8) BNB200_GEMV_MMA(kFP4, 8);
        else BNB200_GEMV_MMA(kFP4, 16);
    }
#undef BNB200_GEMV_MMA
    BNB200_CHECK_LAUNCH("gemv4_mma");
    return true;
}

template bool launch_gemv4_mma<__nv_bfloat16>(const __nv_bfloat16*, const uint8_t*, const float*, const uint8_t*,
                                              const float*, const float*, __nv_bfloat16*, const __nv_bfloat16*, int,
                                              int, int, int, int, int, cudaStream_t);
template bool launch_gemv4_mma<__half>(const __half*, const uint8_t*, const float*, const uint8_t*, const float*,
                                       const float*, __half*, const __half*, int, int, int, int, int, int,
                                       cudaStream_t);

} // namespace bnb200
