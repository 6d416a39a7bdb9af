// int8.cu -- the LLM.int8() hot path for sm_100a.
//
//   int8_vector_quant   replaces reference kInt8VectorQuant (csrc/kernels.cu:1331-1385):
//                       one read of A instead of two (the row is held in registers between
//                       the absmax pass and the quantise pass) and, optionally, outlier-column
//                       flags in the same pass (the reference finds them with 3-4 torch kernels
//                       and a host sync, backends/cuda/ops.py:230-236).
//   int8 GEMM           replaces reference igemmlt<32,0> -> cublasLtMatmul (csrc/ops.cu:282-404)
//                       with a tcgen05 kind::i8 kernel: both operands TMA-staged (128-byte
//                       swizzle), int32 accumulators in TMEM, exact.  Epilogue either stores
//                       int32 (cigemmlt_32 ABI) or applies the dequantisation
//                       fp16/bf16( fma(acc * SCA[m] * SCB[n], 1/127^2, bias[n]) ) in-kernel,
//                       which removes the 2 x M x N x 4-byte int32 round trip through HBM.
//   dequant_mm_int32    replaces reference kdequant_mm_int32_fp16 (csrc/kernels.cu:1396-1448).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace bnb200 {

namespace {

// ======================================================================================
// row-wise int8 quantisation
// ======================================================================================
constexpr int kVqThreads = 256;
constexpr int kVqVecs = 4;  // 16-byte vectors cached per thread -> rows up to 256*4*8 = 8192 columns

template <typename T> __device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__device__ __forceinline__ float block_max(float m, float* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = sred[0];
#pragma unroll
    for (int w = 1; w < kVqThreads / 32; ++w) r = fmaxf(r, sred[w]);
    return r;
}

__device__ __forceinline__ int8_t quant_one(float v, float scale) {
    // __float2int_rn(val * scale) truncated to int8, reference kernels.cu:1376-1381
    return (int8_t)__float2int_rn(mul_ftz(v, scale));
}

// One CTA per row.  thr_stat = threshold as the reference's first pass sees it (rounded to T),
// thr = raw fp32 threshold used by the second pass (reference kernels.cu:1358 vs :1378).
template <typename T, bool kVec>
__global__ void __launch_bounds__(kVqThreads)
    int8_vector_quant_kernel(const T* __restrict__ A, int8_t* __restrict__ out, float* __restrict__ rowStats,
                             int* __restrict__ col_flags, float thr, float thr_stat, int rows, int cols) {
    __shared__ float sred[kVqThreads / 32];
    const int row = blockIdx.x;
    const T* a = A + (long long)row * cols;
    int8_t* o = out + (long long)row * cols;
    const bool sparse = thr > 0.0f;

    // T(-FLT_MIN) is -0.0 for 16-bit T (reference kernels.cu:1353); max with -0.0 keeps it for an
    // all-outlier row.
    float m = -0.0f;

    if constexpr (kVec) {
        uint4 cache[kVqVecs];
        const int nvec = cols >> 3;
#pragma unroll
        for (int j = 0; j < kVqVecs; ++j) {
            const int vi = threadIdx.x + j * kVqThreads;
            cache[j] = make_uint4(0, 0, 0, 0);
            if (vi < nvec) cache[j] = ldg_stream_v4(a + 8 * vi);
        }
#pragma unroll
        for (int j = 0; j < kVqVecs; ++j) {
            const int vi = threadIdx.x + j * kVqThreads;
            if (vi < nvec) {
                float v[8];
                unpack8<T>(cache[j], v);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float av = fabsf(v[t]);
                    if (!sparse || av < thr_stat) m = fmaxf(m, av);
                    if (sparse && col_flags != nullptr && av >= thr) col_flags[8 * vi + t] = 1;
                }
            }
        }
        const float row_absmax = block_max(m, sred);
        if (threadIdx.x == 0) rowStats[row] = row_absmax;
        const float scale = div_approx_ftz(127.0f, row_absmax);  // __fdividef
#pragma unroll
        for (int j = 0; j < kVqVecs; ++j) {
            const int vi = threadIdx.x + j * kVqThreads;
            if (vi < nvec) {
                float v[8];
                unpack8<T>(cache[j], v);
                uint32_t w[2] = {0u, 0u};
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    int8_t q = (!sparse || fabsf(v[t]) < thr) ? quant_one(v[t], scale) : (int8_t)0;
                    w[t >> 2] |= (uint32_t)(uint8_t)q << (8 * (t & 3));
                }
                stg_stream_v2(o + 8 * vi, make_uint2(w[0], w[1]));
            }
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += kVqThreads) {
            const float av = fabsf(DT<T>::to_f32(a[c]));
            if (!sparse || av < thr_stat) m = fmaxf(m, av);
            if (sparse && col_flags != nullptr && av >= thr) col_flags[c] = 1;
        }
        const float row_absmax = block_max(m, sred);
        if (threadIdx.x == 0) rowStats[row] = row_absmax;
        const float scale = div_approx_ftz(127.0f, row_absmax);
        for (int c = threadIdx.x; c < cols; c += kVqThreads) {
            const float v = DT<T>::to_f32(a[c]);
            o[c] = (!sparse || fabsf(v) < thr) ? quant_one(v, scale) : (int8_t)0;
        }
    }
}

// ======================================================================================
// int32 -> fp16 dequantisation epilogue as a stand-alone kernel (cdequant_mm_int32_fp16 ABI)
// ======================================================================================
constexpr float kMmDequantConst = 6.200012e-05f;  // reference kernels.cu:1394 ("1/(127*127)")

__device__ __forceinline__ float dequant_value(int acc, float rs, float cs, float bias) {
    // reference kernels.cu:1436-1438: fmaf(int * rowStats * colStats, C, bias), all ftz
    float t = mul_ftz(mul_ftz((float)acc, rs), cs);
    float r;
    asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(t), "f"(kMmDequantConst), "f"(bias));
    return r;
}

__global__ void __launch_bounds__(256)
    dequant_mm_int32_fp16_kernel(const int* __restrict__ A, const float* __restrict__ rowStats,
                                 const float* __restrict__ colStats, __half* __restrict__ out,
                                 const __half* __restrict__ bias, int numRows, int numCols, int vec_ok) {
    const int row = blockIdx.y;
    const float rs = __ldg(rowStats + row);
    const int* a = A + (long long)row * numCols;
    __half* o = out + (long long)row * numCols;
    if (vec_ok) {
        const int nvec = numCols >> 2;
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
            const uint4 x = ldg_stream_v4(a + 4 * v);
            const float4 cs = __ldg(reinterpret_cast<const float4*>(colStats) + v);
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            if (bias != nullptr) {
                const uint2 bb = __ldg(reinterpret_cast<const uint2*>(bias) + v);
                const float2 b01 = __half22float2(*reinterpret_cast<const __half2*>(&bb.x));
                const float2 b23 = __half22float2(*reinterpret_cast<const __half2*>(&bb.y));
                b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
            }
            const uint32_t lo = pack2<__half>(dequant_value((int)x.x, rs, cs.x, b[0]),
                                              dequant_value((int)x.y, rs, cs.y, b[1]));
            const uint32_t hi = pack2<__half>(dequant_value((int)x.z, rs, cs.z, b[2]),
                                              dequant_value((int)x.w, rs, cs.w, b[3]));
            stg_stream_v2(o + 4 * v, make_uint2(lo, hi));
        }
    } else {
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < numCols; c += gridDim.x * blockDim.x) {
            const float b = bias != nullptr ? __half2float(bias[c]) : 0.f;
            o[c] = __float2half_rn(dequant_value(a[c], rs, __ldg(colStats + c), b));
        }
    }
}

// ======================================================================================
// tcgen05 int8 GEMM:  C[M,N] = A[M,K] . B[N,K]^T   (A = activations, B = weights)
// ======================================================================================
constexpr int kI8Stages = 4;
constexpr int kI8BK = 128;        // int8 elements per k-block = one 128-byte swizzled row
constexpr int kI8TileM = 128;     // tokens per CTA (TMEM lanes)
constexpr int kI8TileN = 256;     // output features per CTA (TMEM columns)
constexpr int kI8Threads = 6 * 32;
constexpr int kI8StageBytes = (kI8TileM + kI8TileN) * 128;

// EPI: 0 = int32 out, 1 = fp16 out, 2 = bf16 out (fused dequant)
struct I8Params {
    void* out;
    const float* SCA;   // [M]  row stats of the activations
    const float* SCB;   // [N]  row stats of the weights
    const void* bias;   // T[N] or NULL
    int M, N, K, ldc;
    int kblocks;
};

template <int EPI>
__global__ void __launch_bounds__(kI8Threads, 1)
    int8_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const I8Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem;
    float* s_scb = reinterpret_cast<float*>(smem + kI8Stages * kI8StageBytes);          // [256]
    float* s_bias = s_scb + kI8TileN;                                                     // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kI8Stages * kI8StageBytes + 2048);
    uint64_t* full = bars;
    uint64_t* empty = bars + kI8Stages;
    uint64_t* acc_full = bars + 2 * kI8Stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kI8Stages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kI8TileN;
    const int m0 = blockIdx.y * kI8TileM;
    constexpr uint32_t kTmemCols = 256;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < kI8Stages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc<kTmemCols>(tmem_slot);
        ptx::tmem_relinquish();
    }
    if (EPI != 0 && warp >= 2) {
        for (int c = threadIdx.x - 64; c < kI8TileN; c += 128) {
            const int n = n0 + c;
            s_scb[c] = (n < p.N) ? __ldg(p.SCB + n) : 0.f;
            float b = 0.f;
            if (p.bias != nullptr && n < p.N) {
                if (EPI == 1)
                    b = __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
                else
                    b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n]);
            }
            s_bias[c] = b;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < p.kblocks; ++i) {
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_arrive_expect_tx(&full[s], kI8StageBytes);
                uint8_t* sa = stages + s * kI8StageBytes;
                ptx::tma_load_2d(sa, &tmap_a, &full[s], i * kI8BK, m0);
                ptx::tma_load_2d(sa + kI8TileM * 128, &tmap_b, &full[s], i * kI8BK, n0);
                if (++s == kI8Stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // kind::i8: D = S32 (2), A/B = signed int8 (1); UMMA K = 32 bytes
        constexpr uint32_t idesc = ptx::make_idesc(2, 1, 1, kI8TileM, kI8TileN);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < p.kblocks; ++i) {
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = ptx::smem_u32(stages + s * kI8StageBytes);
                const uint64_t adesc = ptx::make_sw128_kmajor_desc(sa);
                const uint64_t bdesc = ptx::make_sw128_kmajor_desc(sa + kI8TileM * 128);
#pragma unroll
                for (int j = 0; j < kI8BK / 32; ++j) {
                    ptx::mma_i8_ss(tmem_base, adesc + 2 * j, bdesc + 2 * j, idesc, (i | j) != 0 ? 1u : 0u);
                }
                ptx::tc_commit(&empty[s]);
                if (i == p.kblocks - 1) ptx::tc_commit(acc_full);
            }
            __syncwarp();
            if (++s == kI8Stages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        const int m = m0 + quarter * 32 + lane;
        const bool m_ok = m < p.M;
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        float sca = 0.f;
        if (EPI != 0 && m_ok) sca = __ldg(p.SCA + m);
#pragma unroll 1
        for (int c = 0; c < kI8TileN; c += 32) {
            uint32_t v[32];
            ptx::tmem_ld_x32(lane_addr + c, v);
            ptx::tmem_wait_ld();
            if (!m_ok) continue;
            const int n = n0 + c;
            if (EPI == 0) {
                int* dst = reinterpret_cast<int*>(p.out) + (long long)m * p.ldc + n;
                if (n + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int t = 0; t < 32; t += 4)
                        *reinterpret_cast<uint4*>(dst + t) = make_uint4(v[t], v[t + 1], v[t + 2], v[t + 3]);
                } else {
                    for (int t = 0; t < 32; ++t)
                        if (n + t < p.N) dst[t] = (int)v[t];
                }
            } else {
                uint32_t w[16];
#pragma unroll
                for (int t = 0; t < 32; t += 2) {
                    if (EPI == 1) {
                        const float f0 = dequant_value((int)v[t], sca, s_scb[c + t], s_bias[c + t]);
                        const float f1 = dequant_value((int)v[t + 1], sca, s_scb[c + t + 1], s_bias[c + t + 1]);
                        w[t >> 1] = pack2<__half>(f0, f1);
                    } else {
                        // bf16 output, bit-identical to the reference chain (backends/cuda/ops.py:186-210):
                        // the kernel result is fp16, a non-fp16 bias is added by `out.add_(bias)` on the
                        // fp16 tensor (fp32 add, one rounding to fp16), then `.to(bfloat16)`.
                        float f0 = __half2float(__float2half_rn(dequant_value((int)v[t], sca, s_scb[c + t], 0.f)));
                        float f1 =
                            __half2float(__float2half_rn(dequant_value((int)v[t + 1], sca, s_scb[c + t + 1], 0.f)));
                        if (p.bias != nullptr) {
                            f0 = __half2float(__float2half_rn(f0 + s_bias[c + t]));
                            f1 = __half2float(__float2half_rn(f1 + s_bias[c + t + 1]));
                        }
                        w[t >> 1] = pack2<__nv_bfloat16>(f0, f1);
                    }
                }
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + (long long)m * p.ldc + n;
                if (n + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                    for (int t = 0; t < 16; t += 4)
                        *reinterpret_cast<uint4*>(dst + 2 * t) = make_uint4(w[t], w[t + 1], w[t + 2], w[t + 3]);
                } else {
                    for (int t = 0; t < 32; ++t)
                        if (n + t < p.N) dst[t] = (uint16_t)(w[t >> 1] >> (16 * (t & 1)));
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_dyn(tmem_base, kTmemCols);
    }
}

template <int EPI> int launch_i8(const CUtensorMap& ta, const CUtensorMap& tb, const I8Params& p, cudaStream_t stream) {
    constexpr size_t smem_bytes = 1024 + size_t(kI8Stages) * kI8StageBytes + 2048 + 256;
    static bool attr_set = false;
    auto kern = int8_gemm_tc_kernel<EPI>;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            set_last_error("int8_gemm_tc smem attr", cudaGetLastError());
            return 1;
        }
        attr_set = true;
    }
    dim3 grid((p.N + kI8TileN - 1) / kI8TileN, (p.M + kI8TileM - 1) / kI8TileM);
    kern<<<grid, kI8Threads, smem_bytes, stream>>>(ta, tb, p);
    BNB200_CHECK_LAUNCH("int8_gemm_tc");
    return 0;
}

} // namespace

// ---------------------------------------------------------------- launch wrappers
void launch_int8_vector_quant(const void* A, int8_t* out, float* rowStats, int* col_flags, float threshold, int rows,
                              int cols, int dtype /*1 fp16, 2 bf16*/, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    const bool vec = (cols % 8 == 0) && (cols <= kVqThreads * kVqVecs * 8) &&
                     ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    // threshold as the statistic pass of the reference compares it: rounded to the input type
    float thr_stat = threshold;
    if (dtype == 1) thr_stat = __half2float(__float2half_rn(threshold));
    if (dtype == 2) thr_stat = __bfloat162float(__float2bfloat16_rn(threshold));
    if (dtype == 1) {
        const __half* a = reinterpret_cast<const __half*>(A);
        if (vec)
            int8_vector_quant_kernel<__half, true>
                <<<rows, kVqThreads, 0, stream>>>(a, out, rowStats, col_flags, threshold, thr_stat, rows, cols);
        else
            int8_vector_quant_kernel<__half, false>
                <<<rows, kVqThreads, 0, stream>>>(a, out, rowStats, col_flags, threshold, thr_stat, rows, cols);
    } else {
        const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(A);
        if (vec)
            int8_vector_quant_kernel<__nv_bfloat16, true>
                <<<rows, kVqThreads, 0, stream>>>(a, out, rowStats, col_flags, threshold, thr_stat, rows, cols);
        else
            int8_vector_quant_kernel<__nv_bfloat16, false>
                <<<rows, kVqThreads, 0, stream>>>(a, out, rowStats, col_flags, threshold, thr_stat, rows, cols);
    }
    BNB200_CHECK_LAUNCH("int8_vector_quant");
}

void launch_dequant_mm_int32_fp16(const int* A, const float* rowStats, const float* colStats, __half* out,
                                  const __half* bias, int numRows, int numCols, cudaStream_t stream) {
    if (numRows <= 0 || numCols <= 0) return;
    const bool vec = (numCols % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(colStats) & 15) == 0) &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 7) == 0);
    int per_row = vec ? (numCols / 4 + 255) / 256 : (numCols + 255) / 256;
    if (per_row > 64) per_row = 64;
    // blockIdx.y carries the row; 65535 rows per launch
    for (int r0 = 0; r0 < numRows; r0 += 65535) {
        const int nr = (numRows - r0 < 65535) ? numRows - r0 : 65535;
        dim3 grid(per_row, nr);
        dequant_mm_int32_fp16_kernel<<<grid, 256, 0, stream>>>(A + (long long)r0 * numCols, rowStats + r0, colStats,
                                                               out + (long long)r0 * numCols, bias, nr, numCols,
                                                               vec ? 1 : 0);
    }
    BNB200_CHECK_LAUNCH("dequant_mm_int32_fp16");
}

// int8_persist.cu (experimental persistent / multicast variant)
int launch_int8_gemm_persistent(const int8_t* acts, const int8_t* weights, void* out, const float* SCA,
                                const float* SCB, const void* bias, int M, int N, int K, int ldc, int epi,
                                cudaStream_t stream);

// epi: 0 int32, 1 fp16, 2 bf16.  Returns 0 ok, 100 "not implemented for this shape".
int launch_int8_gemm(const int8_t* acts, const int8_t* weights, void* out, const float* SCA, const float* SCB,
                     const void* bias, int M, int N, int K, int ldc, int epi, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return 0;
    if (K <= 0 || (K % 16) != 0) return 100;
    if ((reinterpret_cast<uintptr_t>(acts) & 15) != 0 || (reinterpret_cast<uintptr_t>(weights) & 15) != 0) return 100;
    static const bool persistent = [] {
        const char* e = getenv("BNB_B200_I8_PERSISTENT");
        return e != nullptr && e[0] == '1';
    }();
    if (persistent) return launch_int8_gemm_persistent(acts, weights, out, SCA, SCB, bias, M, N, K, ldc, epi, stream);
    CUtensorMap ta, tb;
    if (!encode_tmap_2d(&ta, acts, 1, 128, (uint64_t)M, (uint64_t)K, (uint64_t)K, kI8TileM, kI8BK)) return 100;
    if (!encode_tmap_2d(&tb, weights, 1, 128, (uint64_t)N, (uint64_t)K, (uint64_t)K, kI8TileN, kI8BK))
        return 100;
    I8Params p{};
    p.out = out;
    p.SCA = SCA;
    p.SCB = SCB;
    p.bias = bias;
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = ldc;
    p.kblocks = (K + kI8BK - 1) / kI8BK;
    switch (epi) {
    case 0: return launch_i8<0>(ta, tb, p, stream);
    case 1: return launch_i8<1>(ta, tb, p, stream);
    default: return launch_i8<2>(ta, tb, p, stream);
    }
}

} // namespace bnb200
