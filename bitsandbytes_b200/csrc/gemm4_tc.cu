// gemm4_tc.cu -- NF4/FP4 dequant-fused GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces the reference's mma.sync kernel gemm_4bit_sm80_m16n8k16 (reference
// csrc/gemm_4bit_sm80.cu:127-457) and, on B200, the dequantize + cuBLAS fallback the
// reference actually takes for M > 4 (reference bitsandbytes/backends/cuda/ops.py:617-623,
// 904-916).  Contract (reference _ops.py:239-295, gemm_4bit_mma.cuh:99-101):
//
//     out[m, n] = T( sum_k X[m, k] * W_T[n, k]  (fp32 accumulate)  + bias[n] )
//     W_T[n, k] = rn_T( value(code[n, k]) * scale[(n*K + k) / blocksize] )      (one rounding)
//     scale[i]  = absmax[i]                                       (plain)
//               = absmax_code[absmax_8bit[i]] * absmax[i >> 8] + offset   (double quant)
//
// B200-first design ("swap-AB, weights through TMEM"):
//   * The tensor-core M dimension (128 TMEM lanes) carries the OUTPUT FEATURES n; the
//     tokens m are the UMMA N dimension (16..256).  A CTA owns out[m0:m0+MT, n0:n0+128].
//   * The packed 4-bit weights never touch shared memory.  Eight decode warps read them
//     from global (16 B = 32 codes per thread per k-block, prefetched two k-blocks ahead),
//     expand them warp-locally to T with the exact reference rounding, and write the tile
//     straight into TENSOR MEMORY with tcgen05.st; tcgen05.mma consumes it as the A operand
//     ([tmem] form).  Shared memory therefore only carries the activation tile, which is
//     what limits a Blackwell SM in SS mode.
//   * The activation tile X[m0:m0+MT, k0:k0+64] arrives by TMA (128-byte swizzle) and is
//     the B operand (K-major smem descriptor).
//   * One elected thread issues tcgen05.mma (128 x MT x 16, four per 64-wide k-block) and
//     frees each pipeline stage with tcgen05.commit -> mbarrier.
//   * Accumulators (128 lanes x MT fp32 columns) live in TMEM; the decode warps become the
//     epilogue warps: tcgen05.ld -> +bias -> rn_T -> global, or -- for split-K, which fills
//     the 148 SMs when M is small -- fp32 partials to an L2-resident workspace with a
//     last-arriver reduction in deterministic split order.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
// warps 2..9 decode / epilogue (TMEM lane quarter = warp_id % 4, k-half = (warp_id-2)/4).
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <type_traits>

namespace bnb200 {

namespace {

constexpr int kStages = 6;       // pipeline depth (X tiles in smem, W tiles in TMEM): covers the TMA latency
constexpr int kBK = 64;          // k-block: 64 elements = 128 B of 16-bit X per row
constexpr int kTileN = 128;      // output features per CTA (TMEM lanes)
constexpr int kDecodeWarps = 8;
constexpr int kThreads = 32 * (2 + kDecodeWarps);
constexpr int kPrefetch = 4;     // decode-side global prefetch distance (k-blocks of one warp set)

struct Gemm4Params {
    const uint8_t* B;            // packed codes [N, K/2]
    const float* absmax;         // fp32 per block, or level-2 absmax when nested
    const uint8_t* absmax_8bit;  // NULL unless double quant
    const float* absmax_code;    // 256-entry code for absmax_8bit
    const float* absmax_offset;  // scalar
    const void* bias;            // T[N] or NULL
    void* out;                   // T[M, ldc]
    float* ws_partial;           // split-K partials [tiles][splits][128][MT]
    int* ws_counter;             // one per output tile, zero on entry, reset on exit
    int M, N, K, ldc;
    int log2_bs;
    int kblocks_total;           // K / 64
    int kblocks_per_split;
    int splits;
};

template <typename T> struct TcFmt;
template <> struct TcFmt<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };
template <> struct TcFmt<__half> { static constexpr uint32_t kFmt = 0; };

struct ScaleSrc {
    const float* absmax;
    const uint8_t* absmax_8bit;
    const float* absmax_code;
    float offset;
    __device__ __forceinline__ float load(long long idx) const {
        if (absmax_8bit != nullptr) {
            const float c = __ldg(absmax_code + __ldg(absmax_8bit + idx));
            return __fadd_rn(mul_ftz(c, __ldg(absmax + (idx >> 8))), offset);
        }
        return __ldg(absmax + idx);
    }
};

// NOTE on the nested (double-quant) scale: the reference has two behaviours.  Its fused
// kernels write `code[q] * absmax2 + offset`, which nvcc contracts to one fma
// (gemm_4bit_sm80.cu:292-297); its dequantize + F.linear path -- the one B200 takes for
// M > 4 (backends/cuda/ops.py:617-623, 904-916) and the one F.dequantize_4bit exposes --
// rounds the product and the sum separately.  We follow the second (mul, then add), so the
// fused GEMM sees exactly the weights F.dequantize_4bit returns.


// ---------------------------------------------------------------- register-resident decode
// W_T = rn_T(value(code) * scale) takes only 16 distinct values per quantisation block, so a
// decode thread first builds that 16-entry table (16 fp32 multiplies by immediates, 8 packed
// roundings -- bit-identical to rounding every element) and keeps it in 8 registers as two
// byte planes (low bytes / high bytes of the 16-bit entries).  Codes are then translated with
// PRMT (byte permute) only: no shared-memory look-up table, hence no bank conflicts and no
// competition with the tensor core for shared-memory bandwidth.
struct DecodeTable {
    uint32_t lo[4];  // lo[j] = low bytes of entries 4j .. 4j+3
    uint32_t hi[4];  // hi[j] = high bytes
};

template <typename T, int QT> __device__ __forceinline__ void build_table(float scale, DecodeTable& t) {
    uint32_t pr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        pr[j] = pack2<T>(mul_ftz(code4_value<QT>(2 * j), scale), mul_ftz(code4_value<QT>(2 * j + 1), scale));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t.lo[j] = __byte_perm(pr[2 * j], pr[2 * j + 1], 0x6420);
        t.hi[j] = __byte_perm(pr[2 * j], pr[2 * j + 1], 0x7531);
    }
}

// One packed word = 4 bytes = 8 codes (byte b: element 2b in the high nibble) -> 4 registers of
// T pairs, element 2b in the low half.  Selector nibbles must stay < 8 (bit 3 is PRMT's
// sign-replicate flag): `c` carries code & 7, `selm` picks between the idx<8 / idx>=8 halves.
__device__ __forceinline__ void decode_word(uint32_t w, const DecodeTable& t, uint32_t* o) {
    const uint32_t c7 = w & 0x77777777u;
    const uint32_t w1 = w >> 1;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const uint32_t c = g ? (c7 >> 16) : c7;
        const uint32_t m = g ? (w1 >> 16) : w1;
        const uint32_t selm = (m & 0x4444u) | 0x3210u;
        const uint32_t lo = __byte_perm(__byte_perm(t.lo[0], t.lo[1], c), __byte_perm(t.lo[2], t.lo[3], c), selm);
        const uint32_t hi = __byte_perm(__byte_perm(t.hi[0], t.hi[1], c), __byte_perm(t.hi[2], t.hi[3], c), selm);
        o[2 * g] = __byte_perm(lo, hi, 0x4051);      // (T[hi nibble of byte 0], T[lo nibble of byte 0])
        o[2 * g + 1] = __byte_perm(lo, hi, 0x6273);  // byte 1
    }
}

template <typename T, int QT, int MT>
__global__ void __launch_bounds__(kThreads, 1)
    gemm4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const Gemm4Params p) {
    // ------------------------------------------------------------------ shared memory
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // X stages: MT rows x 128 B each, 1024-B aligned (MT >= 16 -> multiple of 2048 B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kXStageBytes = MT * 128;
    uint8_t* sx = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kXStageBytes);
    uint64_t* full_x = bars;                  // [kStages] TMA -> MMA
    uint64_t* full_w = bars + kStages;        // [kStages] decode -> MMA   (count = kDecodeWarps / 2)
    uint64_t* empty = bars + 2 * kStages;     // [kStages] MMA -> TMA + decode (tcgen05.commit)
    uint64_t* acc_full = bars + 3 * kStages;  // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    constexpr uint32_t kTmemCols = 512u;  // D: columns [0, MT); W stages: kWCol0 + s*32, 32 columns each
    constexpr uint32_t kWCol0 = 256u;
    static_assert(kWCol0 + kStages * 32 <= kTmemCols, "TMEM budget");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n0 = blockIdx.x * kTileN;
    const int m0 = blockIdx.y * MT;
    const int split = blockIdx.z;
    const int kb_begin = split * p.kblocks_per_split;
    int kb_end = kb_begin + p.kblocks_per_split;
    if (kb_end > p.kblocks_total) kb_end = p.kblocks_total;
    const int nkb = kb_end - kb_begin;  // >= 1 by construction

    // ------------------------------------------------------------------ setup
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_x[s], 1);
            ptx::mbar_init(&full_w[s], kDecodeWarps / 2);
            ptx::mbar_init(&empty[s], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc<kTmemCols>(tmem_slot);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_arrive_expect_tx(&full_x[s], kXStageBytes);
                ptx::tma_load_2d(sx + s * kXStageBytes, &tmap_x, &full_x[s], (kb_begin + i) * kBK, m0);
                if (++s == kStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        constexpr uint32_t idesc = ptx::make_idesc(/*D=F32*/ 1, TcFmt<T>::kFmt, TcFmt<T>::kFmt, /*M=*/128, /*N=*/MT);
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nkb; ++i) {
            ptx::mbar_wait(&full_x[s], ph);
            ptx::mbar_wait(&full_w[s], ph);
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sx + s * kXStageBytes));
                const uint32_t a_tmem = tmem_base + kWCol0 + s * 32;
#pragma unroll
                for (int j = 0; j < kBK / 16; ++j) {
                    // K advances by 16 elements: +8 TMEM columns for A, +32 B (2 x 16 B) for B
                    ptx::mma_f16_ts(tmem_base, a_tmem + 8 * j, bdesc + 2 * j, idesc, (i | j) != 0 ? 1u : 0u);
                }
                ptx::tc_commit(&empty[s]);
                if (i == nkb - 1) ptx::tc_commit(acc_full);
            }
            __syncwarp();
            if (++s == kStages) {
                s = 0;
                ph ^= 1u;
            }
        }
    } else {
        // ================================================================== decode warps
        const int dw = warp - 2;          // 0..7
        const int quarter = warp & 3;     // TMEM lane quarter this warp may touch
        const int par = dw >> 2;          // warp set: handles k-blocks i with (i & 1) == par
        const int khalf = par;            // epilogue: which half of the accumulator columns
        const int row = quarter * 32 + lane;
        const int n = n0 + row;
        const bool n_ok = n < p.N;
        const uint8_t* wrow = p.B + ((long long)(n_ok ? n : 0) * p.K >> 1);
        const long long e_row = (long long)(n_ok ? n : 0) * p.K;
        ScaleSrc sc{p.absmax, p.absmax_8bit, p.absmax_code, p.absmax_offset ? __ldg(p.absmax_offset) : 0.0f};
        const bool two_scales = p.log2_bs == 5;  // blocksize 32: two quantisation blocks per 64-wide k-block

        // This warp set's k-blocks: i = 2 t + par, t = 0 .. cnt-1.
        const int cnt = (nkb - par + 1) >> 1;

        // Register prefetch ring, kPrefetch of this warp set's k-blocks deep.  The loop is unrolled
        // by kPrefetch so that slot j is a fixed set of registers: no register rotation (a move out
        // of a load's destination would wait for the load and collapse the prefetch).
        uint4 wq[kPrefetch][2];
        float wsc[kPrefetch][2];
#pragma unroll
        for (int j = 0; j < kPrefetch; ++j) {
            wq[j][0] = wq[j][1] = make_uint4(0, 0, 0, 0);
            wsc[j][0] = wsc[j][1] = 0.f;
            if (j < cnt && n_ok) {
                const int kb = kb_begin + 2 * j + par;
                const uint8_t* src = wrow + (long long)kb * (kBK / 2);
                wq[j][0] = ldg_stream_v4(src);
                wq[j][1] = ldg_stream_v4(src + 16);
                const long long e = e_row + (long long)kb * kBK;
                wsc[j][0] = sc.load(e >> p.log2_bs);
                if (two_scales) wsc[j][1] = sc.load((e + 32) >> p.log2_bs);
            }
        }

        for (int t0 = 0; t0 < cnt; t0 += kPrefetch) {
#pragma unroll
            for (int j = 0; j < kPrefetch; ++j) {
                const int t = t0 + j;
                if (t < cnt) {
                    const int i = 2 * t + par;
                    const int s = i % kStages;
                    const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                    const uint4 q0 = wq[j][0], q1 = wq[j][1];
                    const float sc0 = wsc[j][0], sc1 = wsc[j][1];
                    if (t + kPrefetch < cnt && n_ok) {
                        const int kb = kb_begin + 2 * (t + kPrefetch) + par;
                        const uint8_t* src = wrow + (long long)kb * (kBK / 2);
                        wq[j][0] = ldg_stream_v4(src);
                        wq[j][1] = ldg_stream_v4(src + 16);
                        const long long e = e_row + (long long)kb * kBK;
                        wsc[j][0] = sc.load(e >> p.log2_bs);
                        if (two_scales) wsc[j][1] = sc.load((e + 32) >> p.log2_bs);
                    }

                    // 64 codes of row n -> 32 registers of T pairs
                    uint32_t r[32];
                    DecodeTable tab;
                    build_table<T, QT>(sc0, tab);
                    decode_word(q0.x, tab, r + 0);
                    decode_word(q0.y, tab, r + 4);
                    decode_word(q0.z, tab, r + 8);
                    decode_word(q0.w, tab, r + 12);
                    if (two_scales) build_table<T, QT>(sc1, tab);
                    decode_word(q1.x, tab, r + 16);
                    decode_word(q1.y, tab, r + 20);
                    decode_word(q1.z, tab, r + 24);
                    decode_word(q1.w, tab, r + 28);

                    ptx::mbar_wait(&empty[s], ph ^ 1u);  // the MMAs that read this TMEM stage have retired
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + kWCol0 + s * 32;
                    ptx::tmem_st_x32(taddr, r);
                    ptx::tmem_wait_st();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&full_w[s]);
                }
            }
        }

        // ================================================================== epilogue
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();

        // this warp: lanes [quarter*32, +32) (= output features), columns [khalf*MT/2, +MT/2)
        constexpr int kColsPerWarp = MT / 2;
        constexpr int kChunk = (kColsPerWarp >= 32) ? 32 : kColsPerWarp;  // 8 (MT=16), 16, 32
        const int col0 = khalf * kColsPerWarp;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        T* outp = reinterpret_cast<T*>(p.out);
        float bias_v = 0.f;
        if (p.bias != nullptr && n_ok) bias_v = DT<T>::to_f32(reinterpret_cast<const T*>(p.bias)[n]);

        if (p.splits == 1) {
#pragma unroll 1
            for (int c = 0; c < kColsPerWarp; c += kChunk) {
                uint32_t v[32];
                if constexpr (kChunk == 32) {
                    ptx::tmem_ld_x32(lane_addr + col0 + c, v);
                } else {
                    uint32_t v16[16];
                    ptx::tmem_ld_x16(lane_addr + col0 + c, v16);  // MT=16 reads 8 columns past its half: unused
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = v16[t];
                }
                ptx::tmem_wait_ld();
#pragma unroll
                for (int t = 0; t < kChunk; ++t) {
                    const int m = m0 + col0 + c + t;
                    if (n_ok && m < p.M)
                        outp[(long long)m * p.ldc + n] = DT<T>::from_f32(__uint_as_float(v[t]) + bias_v);
                }
            }
        } else {
            // ---- split-K: publish the fp32 partial, last arriver reduces in split order
            const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
            float* my = p.ws_partial + ((long long)(tile_id * p.splits + split) * kTileN + row) * MT;
#pragma unroll 1
            for (int c = 0; c < kColsPerWarp; c += kChunk) {
                uint32_t v[32];
                if constexpr (kChunk == 32) {
                    ptx::tmem_ld_x32(lane_addr + col0 + c, v);
                } else {
                    uint32_t v16[16];
                    ptx::tmem_ld_x16(lane_addr + col0 + c, v16);
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = v16[t];
                }
                ptx::tmem_wait_ld();
#pragma unroll
                for (int t = 0; t < kChunk; t += 4) {
                    *reinterpret_cast<uint4*>(my + col0 + c + t) = make_uint4(v[t], v[t + 1], v[t + 2], v[t + 3]);
                }
            }
            __threadfence();
            // named barrier over the 8 epilogue warps (256 threads); barrier 0 is __syncthreads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                int prev = atomicAdd(p.ws_counter + tile_id, 1);
                *s_flag = (prev == p.splits - 1) ? 1 : 0;
                if (prev == p.splits - 1) p.ws_counter[tile_id] = 0;  // self-reset for the next launch
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (*s_flag) {
                __threadfence();
                const float* base = p.ws_partial + ((long long)(tile_id * p.splits) * kTileN + row) * MT;
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp; c += 4) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sp = 0; sp < p.splits; ++sp) {
                        const float4 x = __ldcg(reinterpret_cast<const float4*>(
                            base + (long long)sp * kTileN * MT + col0 + c));
                        acc.x += x.x;
                        acc.y += x.y;
                        acc.z += x.z;
                        acc.w += x.w;
                    }
                    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int m = m0 + col0 + c + t;
                        if (n_ok && m < p.M) outp[(long long)m * p.ldc + n] = DT<T>::from_f32(a4[t] + bias_v);
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------ teardown
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_dyn(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
struct Workspace {
    void* ptr = nullptr;
    size_t bytes = 0;
    int* counters = nullptr;
    size_t n_counters = 0;
};

constexpr int kMaxWs = 64;
struct WsEntry {
    int device;
    cudaStream_t stream;
    Workspace ws;
    bool used;
};
WsEntry g_ws[kMaxWs];

// Split-K scratch, one per (device, stream) so that launches on different streams never
// share partials or counters.  Grown with plain cudaMalloc on first use / growth only.
Workspace* get_workspace(cudaStream_t stream, size_t partial_bytes, size_t n_counters) {
    int dev = 0;
    cudaGetDevice(&dev);
    WsEntry* e = nullptr;
    for (int i = 0; i < kMaxWs; ++i) {
        if (g_ws[i].used && g_ws[i].device == dev && g_ws[i].stream == stream) {
            e = &g_ws[i];
            break;
        }
    }
    if (e == nullptr) {
        for (int i = 0; i < kMaxWs; ++i) {
            if (!g_ws[i].used) {
                e = &g_ws[i];
                e->used = true;
                e->device = dev;
                e->stream = stream;
                e->ws = Workspace{};
                break;
            }
        }
    }
    if (e == nullptr) return nullptr;
    if (e->ws.bytes < partial_bytes) {
        if (e->ws.ptr) {
            cudaStreamSynchronize(stream);
            cudaFree(e->ws.ptr);
        }
        size_t want = partial_bytes < (size_t(8) << 20) ? (size_t(8) << 20) : partial_bytes;
        if (cudaMalloc(&e->ws.ptr, want) != cudaSuccess) {
            e->ws.ptr = nullptr;
            e->ws.bytes = 0;
            return nullptr;
        }
        e->ws.bytes = want;
    }
    if (e->ws.n_counters < n_counters) {
        if (e->ws.counters) {
            cudaStreamSynchronize(stream);
            cudaFree(e->ws.counters);
        }
        size_t want = n_counters < 4096 ? 4096 : n_counters;
        if (cudaMalloc(&e->ws.counters, want * sizeof(int)) != cudaSuccess) {
            e->ws.counters = nullptr;
            e->ws.n_counters = 0;
            return nullptr;
        }
        cudaMemsetAsync(e->ws.counters, 0, want * sizeof(int), stream);
        e->ws.n_counters = want;
    }
    return &e->ws;
}

template <typename T, int QT, int MT>
bool launch_mt(const CUtensorMap& tmap, Gemm4Params& p, cudaStream_t stream) {
    constexpr size_t smem_bytes = 1024 /*align slack*/ + size_t(kStages) * MT * 128 + 256 /*barriers*/;
    static bool attr_set = false;
    auto kern = gemm4_tc_kernel<T, QT, MT>;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            set_last_error("gemm4_tc smem attr", cudaGetLastError());
            return false;
        }
        attr_set = true;
    }
    const int n_tiles = (p.N + kTileN - 1) / kTileN;
    const int m_tiles = (p.M + MT - 1) / MT;

    // split-K so that small problems still cover the machine: target >= ~1 CTA per SM,
    // at least 4 k-blocks per split.
    const int sms = device_sm_count();
    int splits = 1;
    const int tiles = n_tiles * m_tiles;
    if (tiles * 2 <= sms) {
        // one wave: the largest split count whose grid still fits the machine
        splits = sms / tiles;
        int max_by_k = p.kblocks_total / 4;
        if (max_by_k < 1) max_by_k = 1;
        if (splits > max_by_k) splits = max_by_k;
        if (splits > 16) splits = 16;
    }
    int per = (p.kblocks_total + splits - 1) / splits;
    splits = (p.kblocks_total + per - 1) / per;  // no empty split
    p.splits = splits;
    p.kblocks_per_split = per;
    p.ws_partial = nullptr;
    p.ws_counter = nullptr;
    if (splits > 1) {
        size_t bytes = size_t(tiles) * splits * kTileN * MT * sizeof(float);
        Workspace* ws = get_workspace(stream, bytes, tiles);
        if (ws == nullptr) {
            set_last_error_msg("gemm4_tc: could not allocate the split-K workspace");
            return false;
        }
        p.ws_partial = reinterpret_cast<float*>(ws->ptr);
        p.ws_counter = ws->counters;
    }
    dim3 grid(n_tiles, m_tiles, splits);
    kern<<<grid, kThreads, smem_bytes, stream>>>(tmap, p);
    BNB200_CHECK_LAUNCH("gemm4_tc");
    return true;
}

} // namespace

// Returns true if the tensor-core path handled the call.
template <typename T>
bool launch_gemm4_tc(const T* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit,
                     const float* absmax_code, const float* absmax_offset, T* out, const T* bias, int M, int N, int K,
                     int ldc, int blocksize, int quant_type, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return true;
    if (K < kBK || (K % kBK) != 0) return false;
    if (blocksize < 32 || (blocksize & (blocksize - 1)) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0) return false;
    if (quant_type != kNF4 && quant_type != kFP4) return false;

    int MT = 256;
    if (M <= 16) MT = 16;
    else if (M <= 32) MT = 32;
    else if (M <= 64) MT = 64;
    else if (M <= 128) MT = 128;

    CUtensorMap tmap;
    if (!encode_tmap_2d(&tmap, A, 2, false, std::is_same<T, __half>::value, (uint64_t)M, (uint64_t)K,
                        (uint64_t)K * 2, (uint32_t)MT, (uint32_t)kBK)) {
        return false;
    }
    Gemm4Params p{};
    p.B = B;
    p.absmax = absmax;
    p.absmax_8bit = absmax_8bit;
    p.absmax_code = absmax_code;
    p.absmax_offset = absmax_offset;
    p.bias = bias;
    p.out = out;
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = ldc;
    p.log2_bs = ilog2_pow2(blocksize);
    p.kblocks_total = K / kBK;

#define BNB200_DISPATCH_MT(QT)                                                                                         \
    switch (MT) {                                                                                                      \
    case 16: return launch_mt<T, QT, 16>(tmap, p, stream);                                                             \
    case 32: return launch_mt<T, QT, 32>(tmap, p, stream);                                                             \
    case 64: return launch_mt<T, QT, 64>(tmap, p, stream);                                                             \
    case 128: return launch_mt<T, QT, 128>(tmap, p, stream);                                                           \
    default: return launch_mt<T, QT, 256>(tmap, p, stream);                                                            \
    }
    if (quant_type == kNF4) {
        BNB200_DISPATCH_MT(kNF4)
    } else {
        BNB200_DISPATCH_MT(kFP4)
    }
#undef BNB200_DISPATCH_MT
}

template bool launch_gemm4_tc<__nv_bfloat16>(const __nv_bfloat16*, const uint8_t*, const float*, const uint8_t*,
                                             const float*, const float*, __nv_bfloat16*, const __nv_bfloat16*, int,
                                             int, int, int, int, int, cudaStream_t);
template bool launch_gemm4_tc<__half>(const __half*, const uint8_t*, const float*, const uint8_t*, const float*,
                                      const float*, __half*, const __half*, int, int, int, int, int, int,
                                      cudaStream_t);

} // namespace bnb200
