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
//   * A pipeline stage is 128 k-elements.  The TMA producer stages, per stage, the packed
//     codes of the CTA's 128 rows (128 x 64 B, 64-byte swizzle, 8 KB) and the activation tile
//     X[m0:m0+MT, k0:k0+128] (two 128-byte-swizzled sub-tiles).
//   * This one-CTA kernel serves M < 512 (with split-K when the grid would not fill the machine); larger
//     token counts go to the CTA-pair kernel of gemm4_pair.cu.  The cluster-multicast, cta_group::2, all-16-warp
//     and persistent variants that round 1 kept here behind environment switches were measured slower
//      and are gone.
//   * 16 decode warps (two groups that alternate stages) expand the codes in REGISTERS with
//     the exact reference rounding -- a per-block 16-entry table built with 16 FMUL + 8
//     cvt.rn.bf16x2 and looked up with PRMT only -- and write the 16-bit tile straight into
//     TENSOR MEMORY with tcgen05.st.  tcgen05.mma consumes it as the A operand ([tmem] form):
//     the decoded weights never pass through shared memory, whose bandwidth is what limits an
//     SS-mode Blackwell GEMM.  The activation tile is the K-major B operand (smem descriptor).
//   * One elected thread issues eight tcgen05.mma (128 x MT x 16) per stage behind ONE
//     mbarrier wait and releases the stage with ONE tcgen05.commit (multicast to the
//     cluster): synchronisation on the issuing thread was the first bottleneck found.
//   * Accumulators (128 lanes x MT fp32 columns) live in TMEM; the decode warps become the
//     epilogue warps: tcgen05.ld -> +bias -> rn_T -> global, or -- for split-K, which fills
//     the 148 SMs when M is small -- fp32 partials to an L2-resident workspace with a
//     last-arriver reduction in deterministic split order.
//
// Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
// warps 2..17 decode / epilogue (TMEM lane quarter = warp_id % 4).
#include "common.cuh"
#include "decode4.cuh"
#include "sm100_ptx.cuh"

#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace bnb200 {

namespace {

constexpr int kBK = 128;         // pipeline stage: 128 k-elements = two 128-byte swizzle atoms of X per row
constexpr int kTileN = 128;      // output features per CTA (TMEM lanes)
constexpr int kDecodeWarps = 16;  // two groups of 8: group g decodes the stages i with i % 2 == g
constexpr int kThreads = 32 * (2 + kDecodeWarps);
constexpr int kScaleDepth = 4;   // decode-side register ring for the per-block scales (stages of one group)

struct Gemm4Params {
    const uint8_t* B;            // packed codes [N, K/2]
    const float* absmax;         // fp32 per block, or level-2 absmax when nested
    const uint8_t* absmax_8bit;  // NULL unless double quant
    const float* absmax_code;    // 256-entry code for absmax_8bit
    const float* absmax_offset;  // scalar
    const void* bias;            // T[N] or NULL
    void* out;                   // T[M, ldc]
    void* peer_out[7];           // further copies of the output tile (peer GPUs' gather buffers, same ldc): the
    int n_peers;                 //   epilogue stores every element to all of them (fused all-gather)
    float* ws_partial;           // split-K partials [tiles][splits][128][MT]
    int* ws_counter;             // one per output tile, zero on entry, reset on exit
    int M, N, K, ldc;
    int log2_bs;
    int kblocks_total;           // number of 128-wide stages = ceil(K / 128)
    int n_tiles;                 // N tiles of 128 (tile = m_tile * n_tiles + n_tile)
    int tiles_total;
    int splits;                  // K splits per tile (1 = none)
};

template <typename T> struct TcFmt;
template <> struct TcFmt<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };
template <> struct TcFmt<__half> { static constexpr uint32_t kFmt = 0; };



// Pipeline stage = 128 k-elements: two 64-wide (128-byte, swizzle-atom) activation sub-tiles in
// shared memory and 64 TMEM columns of decoded weights.  One `full` and one `empty` mbarrier per
// stage keep the synchronisation cost on the single MMA-issuing thread at one wait + one commit
// per eight tcgen05.mma (it was the bottleneck with 64-wide stages and separate barriers).
//
template <int MT> struct StageCfg {
    static constexpr int kXRows = MT;                            // token rows staged by this CTA
    static constexpr int kStages = MT == 256 ? 3 : 4;            // smem: kStages * (kXRows * 256 + 8192) B <= 216 KB
    static constexpr int kXSubBytes = kXRows * 128;              // one 64-wide sub-tile
    static constexpr int kXStageBytes = 2 * kXSubBytes;
    static constexpr int kWStageBytes = kTileN * 64;             // packed codes: 128 rows x 64 B (TMA, 64-B swizzle)
    static constexpr int kStageBytes = kXStageBytes + kWStageBytes;
    static constexpr uint32_t kTmemCols = 512u;                  // D: [0, MT); W stage s: kWCol0 + 64 s
    static constexpr uint32_t kWCol0 = 256u;
    static_assert(kWCol0 + kStages * 64 <= kTmemCols, "TMEM budget");
};

template <typename T, int QT, int MT>
__global__ void __launch_bounds__(kThreads, 1)
    gemm4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const Gemm4Params p) {
    using Cfg = StageCfg<MT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kXSubBytes = Cfg::kXSubBytes;
    constexpr int kXStageBytes = Cfg::kXStageBytes;
    constexpr uint32_t kTmemCols = Cfg::kTmemCols;
    constexpr uint32_t kWCol0 = Cfg::kWCol0;

    // ------------------------------------------------------------------ shared memory
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kWStageBytes = Cfg::kWStageBytes;
    uint8_t* sx = smem;                              // [kStages][2][MT x 128 B]   activations
    uint8_t* sw = smem + kStages * kXStageBytes;     // [kStages][128 x 64 B]      packed codes
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full = bars;                   // [kStages] TMA (1 arrive + X bytes) + the 8 warps of one decode group -> MMA
    uint64_t* empty = bars + kStages;        // [kStages] MMA -> TMA producer + decode warps
    uint64_t* w_full = bars + 2 * kStages;   // [kStages] TMA (1 arrive + W bytes) -> decode warps
    uint64_t* acc_full = bars + 3 * kStages; // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Linear CTA id -> (tile, K-split): ids are ordered (tile, split)
    const int splits = p.splits;
    const int tile_id = blockIdx.x / splits;
    const int split = blockIdx.x - tile_id * splits;
    const int slot0 = tile_id * splits;  // first workspace slot of this tile
    const int n0 = (tile_id % p.n_tiles) * kTileN;
    const int m0 = (tile_id / p.n_tiles) * MT;
    // work is split in 128-wide stages
    const int per = (p.kblocks_total + splits - 1) / splits;
    const int st_begin = split * per;
    int st_end = st_begin + per;
    if (st_end > p.kblocks_total) st_end = p.kblocks_total;
    const int nst = st_end - st_begin;  // >= 1 by construction
    const int kb64_total = p.K / 64;    // the last stage may hold a single 64-wide block

    // ------------------------------------------------------------------ setup
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full[s], 1 + kDecodeWarps / 2);
            ptx::mbar_init(&empty[s], 1);
            ptx::mbar_init(&w_full[s], 1);
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
        // (elect.sync instead of `lane == 0`: a region ptxas KNOWS to be single-threaded keeps TMA / MMA operands in
        // uniform registers; with `lane == 0` every tcgen05.mma sits in an ELECT + R2UR loop, ~94 cycles each)
        if (ptx::elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nst; ++i) {
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                const int k0 = (st_begin + i) * kBK;
                // packed codes of this CTA's 128 output features: bytes [k0/2, k0/2 + 64) of rows n0..n0+127
                ptx::mbar_arrive_expect_tx(&w_full[s], kWStageBytes);
                ptx::tma_load_2d(sw + s * kWStageBytes, &tmap_w, &w_full[s], k0 / 2, n0);
                ptx::mbar_arrive_expect_tx(&full[s], kXStageBytes);
                uint8_t* dst = sx + s * kXStageBytes;
                // columns past K are out of bounds for the tensor map: TMA zero-fills them
#pragma unroll
                for (int h = 0; h < 2; ++h) ptx::tma_load_2d(dst + h * kXSubBytes, &tmap_x, &full[s], k0 + 64 * h, m0);
                if (++s == kStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one elected thread)
        // The commit that releases stage i-1 follows the second MMA of stage i, and the barrier of stage i+1 is
        // probed before the last MMAs of stage i (see gemm4_pair.cu): nothing but MMAs between two stages.
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc(/*D=F32*/ 1, TcFmt<T>::kFmt, TcFmt<T>::kFmt, /*M=*/128, /*N=*/MT);
            int s = 0, ps = 0;
            uint32_t ph = 0;
            ptx::mbar_wait(&full[0], 0);
            for (int i = 0; i < nst; ++i) {
                ptx::tc_fence_after();
                int ns = s + 1;
                uint32_t nph = ph;
                if (ns == kStages) {
                    ns = 0;
                    nph ^= 1u;
                }
                const bool more = i + 1 < nst;
                bool ok = false;
                const uint32_t xs = ptx::smem_u32(sx + s * kXStageBytes);
                const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(xs);
                const uint64_t bdesc1 = ptx::make_sw128_kmajor_desc(xs + kXSubBytes);
                const uint32_t a_tmem = tmem_base + kWCol0 + s * 64;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // K advances by 16 elements: +8 TMEM columns for A, +32 B (2 x 16 B) inside a sub-tile
                    ptx::mma_f16_ts(tmem_base, a_tmem + 8 * j, (j < 4 ? bdesc0 : bdesc1) + 2 * (j & 3), idesc,
                                    (i | j) != 0 ? 1u : 0u);
                    if (j == 1 && i > 0) ptx::tc_commit(&empty[ps]);
                    if (j == 6 && more) ok = ptx::mbar_try_wait(&full[ns], nph);
                }
                if (more && !ok) ptx::mbar_wait(&full[ns], nph);
                ps = s;
                s = ns;
                ph = nph;
            }
            ptx::tc_commit(&empty[ps]);
            ptx::tc_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ================================================================== decode warps
        const int dw = warp - 2;          // 0..15
        const int quarter = warp & 3;     // TMEM lane quarter this warp may touch
        const int grp = dw >> 3;          // decode group: owns the stages with (stage & 1) == grp
        const int half = (dw >> 2) & 1;   // which 64 of the stage's 128 k-elements
        const int khalf = dw >> 2;        // epilogue: which quarter of the accumulator columns (0..3)
        const int row = quarter * 32 + lane;
        const int n = n0 + row;
        const bool n_ok = n < p.N;
        const long long e_row = (long long)(n_ok ? n : 0) * p.K;
        ScaleSrc sc{p.absmax, p.absmax_8bit, p.absmax_code, p.absmax_offset ? __ldg(p.absmax_offset) : 0.0f};
        const bool two_scales = p.log2_bs == 5;  // blocksize 32: two quantisation blocks per 64 codes

        // The packed codes arrive by TMA (64-byte swizzle: 16-byte chunk c of row r sits at chunk
        // c ^ ((r >> 1) & 3)); this thread owns chunks 2*half and 2*half+1 of its row.
        const uint32_t sw_row = (uint32_t)row * 64u;
        const uint32_t sw_c0 = (uint32_t)(((2 * half) ^ ((row >> 1) & 3)) * 16);
        const uint32_t sw_c1 = (uint32_t)(((2 * half + 1) ^ ((row >> 1) & 3)) * 16);

        // Scales are scattered 4-byte loads (one row per lane): a register ring kScaleDepth stages
        // deep hides their latency.  The loop is unrolled by the ring depth so that slot j is a fixed
        // register (no rotation: a move out of a load's destination would wait for the load).
        float wsc[kScaleDepth][2];
        auto fetch = [&](int j, int t) {
            wsc[j][0] = wsc[j][1] = 0.f;
            const int stage_idx = 2 * t + grp;                 // this group's t-th stage
            const int kb = 2 * (st_begin + stage_idx) + half;  // 64-wide k-block of this warp
            if (stage_idx < nst && n_ok && kb < kb64_total) {
                const long long e = e_row + (long long)kb * 64;
                wsc[j][0] = sc.load(e >> p.log2_bs);
                if (two_scales) wsc[j][1] = sc.load((e + 32) >> p.log2_bs);
            }
        };
#pragma unroll
        for (int j = 0; j < kScaleDepth; ++j) fetch(j, j);

        const int cnt = (nst - grp + 1) >> 1;  // number of stages this group owns
        for (int t0 = 0; t0 < cnt; t0 += kScaleDepth) {
#pragma unroll
            for (int j = 0; j < kScaleDepth; ++j) {
                const int t = t0 + j;
                if (t < cnt) {
                    const int i = 2 * t + grp;
                    const int s = i % kStages;
                    const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                    const float sc0 = wsc[j][0], sc1 = wsc[j][1];
                    fetch(j, t + kScaleDepth);

                    ptx::mbar_wait(&w_full[s], ph);  // this stage's codes have landed
                    const uint8_t* wt = sw + s * kWStageBytes + sw_row;
                    const uint4 q0 = *reinterpret_cast<const uint4*>(wt + sw_c0);
                    const uint4 q1 = *reinterpret_cast<const uint4*>(wt + sw_c1);

                    // 64 codes of row n -> 32 registers of T pairs.  (Rows past N and k-blocks past K are
                    // zero-filled by TMA and carry scale 0: they decode to +-0.)
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

                    // w_full[s] completing implies the producer saw empty[s]; waiting on it here as well
                    // makes this warp itself an observer of the MMA completion before it overwrites TMEM.
                    ptx::mbar_wait(&empty[s], ph ^ 1u);
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + kWCol0 + s * 64 + half * 32;
                    ptx::tmem_st_x32(taddr, r);
                    ptx::tmem_wait_st();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&full[s]);
                }
            }
        }

        // ================================================================== epilogue
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();

        // this warp: lanes [quarter*32, +32) (= output features), columns [khalf*MT/4, +MT/4)
        constexpr int kColsPerWarp = MT / 4;
        constexpr int kChunk = (kColsPerWarp >= 32) ? 32 : kColsPerWarp;  // 4 (MT=16), 8, 16, 32
        const int col0 = khalf * kColsPerWarp;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        T* outp = reinterpret_cast<T*>(p.out);
        float bias_v = 0.f;
        if (p.bias != nullptr && n_ok) bias_v = DT<T>::to_f32(reinterpret_cast<const T*>(p.bias)[n]);

        if (splits == 1) {
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
                    if (n_ok && m < p.M) {
                        const T val = DT<T>::from_f32(__uint_as_float(v[t]) + bias_v);
                        const long long idx = (long long)m * p.ldc + n;
                        outp[idx] = val;
                        for (int r = 0; r < p.n_peers; ++r) reinterpret_cast<T*>(p.peer_out[r])[idx] = val;
                    }
                }
            }
        } else {
            // ---- split-K: every split CTA publishes its fp32 partial tile (layout [column m][row n], so
            // that both the write and the later reads are 128-byte coalesced), the splits of a tile
            // rendezvous on a counter, and EACH of them then reduces a 1/splits share of the columns
            // (in split order: deterministic).  The launch is COOPERATIVE (all CTAs of the <= one-wave grid are
            // resident together by contract), so the short wait cannot starve; it is bounded anyway.
            float* ws_tile = p.ws_partial + (long long)slot0 * kTileN * MT;
            float* my = ws_tile + (long long)split * kTileN * MT;
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
                for (int t = 0; t < kChunk; ++t) my[(col0 + c + t) * kTileN + row] = __uint_as_float(v[t]);
            }
            __threadfence();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            int* arrive = p.ws_counter + tile_id;
            int* done = p.ws_counter + p.tiles_total + tile_id;
            if (threadIdx.x == 64) {
                atomicAdd(arrive, 1);
                unsigned long long t0 = 0;
                unsigned spins = 0;
                while (atomicAdd(arrive, 0) < splits) {
                    __nanosleep(64);
                    if ((++spins & 0xFFF) == 0) {  // bounded: report + trap after 10 s instead of hanging the device
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 10000000000ull) {
                            printf("bnb200: split-K rendezvous timed out (tile %d split %d of %d)\n", tile_id, split, splits);
                            __trap();
                        }
                    }
                }
                __threadfence();
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            {
                const int e = threadIdx.x - 64;          // 0..511 over the 16 epilogue warps
                const int rn = e & (kTileN - 1);         // output feature inside the tile
                const int cg = e >> 7;                   // 0..3
                const int nn = n0 + rn;
                float bias_r = 0.f;
                if (p.bias != nullptr && nn < p.N) bias_r = DT<T>::to_f32(reinterpret_cast<const T*>(p.bias)[nn]);
                for (int c = split + splits * cg; c < MT; c += splits * 4) {
                    const int m = m0 + c;
                    if (m >= p.M) break;
                    float acc = 0.f;
                    for (int sp = 0; sp < splits; ++sp)
                        acc += __ldcg(ws_tile + ((long long)sp * MT + c) * kTileN + rn);
                    if (nn < p.N) {
                        const T val = DT<T>::from_f32(acc + bias_r);
                        const long long idx = (long long)m * p.ldc + nn;
                        outp[idx] = val;
                        for (int r = 0; r < p.n_peers; ++r) reinterpret_cast<T*>(p.peer_out[r])[idx] = val;
                    }
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (threadIdx.x == 64) {
                // last split out resets the tile's counters for the next launch
                if (atomicAdd(done, 1) == splits - 1) {
                    *arrive = 0;
                    *done = 0;
                    __threadfence();
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

// Split-K scratch: one FIXED-SIZE block per (device, stream), so that launches on different streams never
// share partials or counters and the launch path never reallocates.  32 MB covers every split this
// library launches (one-CTA kernel: <= one wave of 128 x 256 fp32 tiles = 19 MB; pair kernel: <= 148 split
// CTAs x 128 x 384 fp32 = 29 MB).  The only allocation happens on the first split-K call of a stream; it is
// made capture-safe (relaxed capture mode around cudaMalloc) so that a CUDA-graph capture whose first
// split-K GEMM is inside the capture still works.  The registry evicts its least recently used entry.
constexpr size_t kWsBytes = size_t(32) << 20;
constexpr size_t kWsCounters = 8192;
constexpr int kMaxWs = 64;
struct WsEntry {
    int device;
    cudaStream_t stream;
    Workspace ws;
    bool used;
    unsigned long long stamp;
};
WsEntry g_ws[kMaxWs];
unsigned long long g_ws_clock = 0;
std::mutex g_ws_mu;  // the registry is shared by every host thread that launches GEMMs

Workspace* get_workspace(cudaStream_t stream, size_t partial_bytes, size_t n_counters) {
    if (partial_bytes > kWsBytes || n_counters > kWsCounters) return nullptr;
    std::lock_guard<std::mutex> lk(g_ws_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    WsEntry* e = nullptr;
    WsEntry* lru = &g_ws[0];
    for (int i = 0; i < kMaxWs; ++i) {
        if (g_ws[i].used && g_ws[i].device == dev && g_ws[i].stream == stream) {
            e = &g_ws[i];
            break;
        }
        if (!g_ws[i].used) {
            if (lru->used) lru = &g_ws[i];
        } else if (lru->used && g_ws[i].stamp < lru->stamp) {
            lru = &g_ws[i];
        }
    }
    if (e != nullptr) {
        e->stamp = ++g_ws_clock;
        return &e->ws;
    }
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    cudaThreadExchangeStreamCaptureMode(&mode);
    e = lru;
    if (e->used) {
        // evict: cudaFree waits for the device, so no kernel can still be using the block
        int prev = dev;
        cudaSetDevice(e->device);
        cudaFree(e->ws.ptr);
        cudaFree(e->ws.counters);
        cudaSetDevice(prev);
        e->used = false;
    }
    Workspace w{};
    bool ok = cudaMalloc(&w.ptr, kWsBytes) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&w.counters), kWsCounters * sizeof(int)) == cudaSuccess &&
              cudaMemset(w.counters, 0, kWsCounters * sizeof(int)) == cudaSuccess;  // synchronous: visible to every stream
    cudaThreadExchangeStreamCaptureMode(&mode);
    if (!ok) {
        (void)cudaGetLastError();
        if (w.ptr) cudaFree(w.ptr);
        if (w.counters) cudaFree(w.counters);
        return nullptr;
    }
    w.bytes = kWsBytes;
    w.n_counters = kWsCounters;
    e->ws = w;
    e->device = dev;
    e->stream = stream;
    e->used = true;
    e->stamp = ++g_ws_clock;
    return &e->ws;
}

template <typename T, int QT, int MT>
bool launch_mt(const T* A, Gemm4Params& p, cudaStream_t stream) {
    using Cfg = StageCfg<MT>;
    constexpr size_t smem_bytes = 1024 /*align slack*/ + size_t(Cfg::kStages) * Cfg::kStageBytes + 256 /*barriers*/;
    // the shared-memory opt-in is PER DEVICE (one process may drive several GPUs)
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    auto kern = gemm4_tc_kernel<T, QT, MT>;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            set_last_error("gemm4_tc smem attr", cudaGetLastError());
            return false;
        }
        attr_set[dev] = true;
    }
    CUtensorMap tmap, tmap_w;
    if (!encode_tmap_2d(&tmap, A, 2, 128, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, (uint32_t)MT, 64u)) return false;
    // packed codes as a [N, K/2] byte matrix, 128 x 64-byte boxes, 64-byte swizzle
    if (!encode_tmap_2d(&tmap_w, p.B, 1, 64, (uint64_t)p.N, (uint64_t)p.K / 2, (uint64_t)p.K / 2, (uint32_t)kTileN, 64u))
        return false;
    const int n_tiles = (p.N + kTileN - 1) / kTileN;
    const int m_tiles = (p.M + MT - 1) / MT;
    const int sms = device_sm_count();
    const int tiles = n_tiles * m_tiles;

    // K-splitting for small problems: a uniform split so that ONE wave covers the machine.  Split CTAs exchange
    // fp32 partials through an L2-resident workspace and every split reduces its share of the columns.
    int splits = 1;
    if (tiles * 2 <= sms) {
        int v = sms / tiles;
        const int max_by_k = p.kblocks_total / 2 > 0 ? p.kblocks_total / 2 : 1;  // >= two 128-wide stages per split
        if (v > max_by_k) v = max_by_k;
        if (v > 16) v = 16;
        if (v < 1) v = 1;
        const int per = (p.kblocks_total + v - 1) / v;
        splits = (p.kblocks_total + per - 1) / per;  // no empty split
    }
    p.splits = splits;
    p.n_tiles = n_tiles;
    p.tiles_total = tiles;
    p.ws_partial = nullptr;
    p.ws_counter = nullptr;
    if (splits > 1) {
        Workspace* ws = get_workspace(stream, size_t(tiles) * splits * kTileN * MT * sizeof(float), 2 * (size_t)tiles);
        if (ws == nullptr) {
            set_last_error_msg("gemm4_tc: could not allocate the split-K workspace");
            return false;
        }
        p.ws_partial = reinterpret_cast<float*>(ws->ptr);
        p.ws_counter = ws->counters;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles * splits, 1, 1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (splits > 1) {
        // the splits of a tile rendezvous in the epilogue: a COOPERATIVE launch makes the runtime schedule the
        // whole (<= one wave) grid at once, so the wait cannot starve behind other streams' kernels
        attr[na].id = cudaLaunchAttributeCooperative;
        attr[na].val.cooperative = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap, tmap_w, p);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_last_error("gemm4_tc launch", e);
        return false;
    }
    BNB200_CHECK_LAUNCH("gemm4_tc");
    return true;
}

} // namespace

struct Gemm4Workspace {
    float* partial;
    int* counters;
};
bool gemm4_get_workspace(cudaStream_t stream, size_t partial_bytes, size_t n_counters, Gemm4Workspace* out) {
    Workspace* ws = get_workspace(stream, partial_bytes, n_counters);
    if (ws == nullptr) return false;
    out->partial = reinterpret_cast<float*>(ws->ptr);
    out->counters = ws->counters;
    return true;
}

template <typename T>
bool launch_gemm4_pair(const T* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit,
                       const float* absmax_code, const float* absmax_offset, T* out, const T* bias, int M, int N, int K,
                       int ldc, int blocksize, int quant_type, cudaStream_t stream, void* const* peers, int n_peers,
                       int mt_override, int force_splits, long long* trace);

// Returns true if the tensor-core path handled the call.
// `peers` / `n_peers`: up to 7 additional output bases (same ldc) that receive a copy of every element.
template <typename T>
bool launch_gemm4_tc(const T* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit,
                     const float* absmax_code, const float* absmax_offset, T* out, const T* bias, int M, int N, int K,
                     int ldc, int blocksize, int quant_type, cudaStream_t stream, void* const* peers, int n_peers) {
    if (n_peers < 0 || n_peers > 7) return false;
    if (M <= 0 || N <= 0) return true;
    if (K < 64 || (K % 64) != 0) return false;
    if (blocksize < 32 || (blocksize & (blocksize - 1)) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0) return false;
    if (quant_type != kNF4 && quant_type != kFP4) return false;

    // Large token counts: the CTA-pair kernel (gemm4_pair.cu).  BNB_B200_PAIR_KERNEL=0 keeps the one-CTA kernel.
    {
        static int use_pair = -1;
        if (use_pair < 0) {
            const char* e = getenv("BNB_B200_PAIR_KERNEL");
            use_pair = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        if (use_pair && M >= 512 &&
            launch_gemm4_pair<T>(A, B, absmax, absmax_8bit, absmax_code, absmax_offset, out, bias, M, N, K, ldc,
                                 blocksize, quant_type, stream, peers, n_peers, 0, 0, nullptr))
            return true;
    }

    int MT = 256;
    if (M <= 16) MT = 16;
    else if (M <= 32) MT = 32;
    else if (M <= 64) MT = 64;
    else if (M <= 128) MT = 128;

    Gemm4Params p{};
    p.B = B;
    p.absmax = absmax;
    p.absmax_8bit = absmax_8bit;
    p.absmax_code = absmax_code;
    p.absmax_offset = absmax_offset;
    p.bias = bias;
    p.out = out;
    p.n_peers = n_peers;
    for (int r = 0; r < n_peers; ++r) p.peer_out[r] = peers[r];
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = ldc;
    p.log2_bs = ilog2_pow2(blocksize);
    p.kblocks_total = (K + kBK - 1) / kBK;

#define BNB200_DISPATCH_MT(QT)                                                                                         \
    switch (MT) {                                                                                                      \
    case 16: return launch_mt<T, QT, 16>(A, p, stream);                                                                \
    case 32: return launch_mt<T, QT, 32>(A, p, stream);                                                                \
    case 64: return launch_mt<T, QT, 64>(A, p, stream);                                                                \
    case 128: return launch_mt<T, QT, 128>(A, p, stream);                                                              \
    default: return launch_mt<T, QT, 256>(A, p, stream);                                                               \
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
                                             int, int, int, int, int, cudaStream_t, void* const*, int);
template bool launch_gemm4_tc<__half>(const __half*, const uint8_t*, const float*, const uint8_t*, const float*,
                                      const float*, __half*, const __half*, int, int, int, int, int, int,
                                      cudaStream_t, void* const*, int);

} // namespace bnb200
