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
//     X[m0:m0+MT, k0:k0+128] (two 128-byte-swizzled sub-tiles).  In a cluster of CL n-tiles
//     each CTA fetches 1/CL of the activation rows and MULTICASTS them to its peers.
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
    int kblocks_per_split;       // unused by the kernel (kept for bookkeeping)
    int n_tiles;                 // N tiles of 128 (tile = m_tile * n_tiles + n_tile)
    int tiles_total;
    int tiles_main;              // tiles [0, tiles_main) run with `splits` K-splits, the rest (the partial last
    int splits_tail;             //   wave) with `splits_tail` K-splits so that it fills the machine
    int splits;
    int debug;                   // developer knobs (BNB_B200_DEBUG): 1 skip decode math, 2 skip tcgen05.st, 4 skip TMA, 8 skip MMA
};

template <typename T> struct TcFmt;
template <> struct TcFmt<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };
template <> struct TcFmt<__half> { static constexpr uint32_t kFmt = 0; };



// Pipeline stage = 128 k-elements: two 64-wide (128-byte, swizzle-atom) activation sub-tiles in
// shared memory and 64 TMEM columns of decoded weights.  One `full` and one `empty` mbarrier per
// stage keep the synchronisation cost on the single MMA-issuing thread at one wait + one commit
// per eight tcgen05.mma (it was the bottleneck with 64-wide stages and separate barriers).
//
// PAIR (cta_group::2, BNB_B200_PAIR=1, measured in round 1: bit-identical, 148 us vs 130 us at 4096^3):
// the two CTAs of a cluster own adjacent 128-feature tiles of the SAME token tile.  One tcgen05.mma
// issued by the leader drives both SMs (M = 256: 128 decoded rows from each CTA's TMEM); the
// activation tile is split between the two shared memories (MT/2 tokens each), so every SM ingests
// half the activation bytes: 40 KB per stage instead of 72 KB at MT = 256.  72 KB per 1024 MMA cycles
// is 70 B/clk against the ~41 B/clk/SM the L2 -> SM path sustains (11.5 TB/s chip-wide), which is
// the 0.6 roofline fraction of the single-CTA kernel; the pair removes that bound but its decode ->
// MMA hand-off crosses SMs every stage and needs a redesign before it wins.
template <int MT, bool PAIR = false> struct StageCfg {
    static constexpr int kXRows = PAIR ? MT / 2 : MT;            // token rows staged by this CTA
    static constexpr int kStages = (MT == 256 && !PAIR) ? 3 : 4; // smem: kStages * (kXRows * 256 + 8192) B <= 216 KB
    static constexpr int kXSubBytes = kXRows * 128;              // one 64-wide sub-tile
    static constexpr int kXStageBytes = 2 * kXSubBytes;
    static constexpr int kWStageBytes = kTileN * 64;             // packed codes: 128 rows x 64 B (TMA, 64-B swizzle)
    static constexpr int kStageBytes = kXStageBytes + kWStageBytes;
    static constexpr uint32_t kTmemCols = 512u;                  // D: [0, MT); W stage s: kWCol0 + 64 s
    static constexpr uint32_t kWCol0 = 256u;
    static_assert(kWCol0 + kStages * 64 <= kTmemCols, "TMEM budget");
};

// D16 (BNB_B200_DECODE16=1, experimental, written at the end of round 1 and not yet run): all 16 decode
// warps work on EVERY stage (each thread 32 codes of its row, one tcgen05.st.x16) instead of two
// groups of 8 that alternate stages.  The table of a 64-element quantisation block is then built by
// two threads (+14 % decode instructions), but a stage is decoded in one stage time instead of two,
// which is what the pair variant's cross-SM hand-off needs.
template <typename T, int QT, int MT, int CL, bool PAIR, bool D16>
__global__ void __launch_bounds__(kThreads, 1)
    gemm4_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const Gemm4Params p) {
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
    using Cfg = StageCfg<MT, PAIR>;
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
    uint64_t* empty = bars + kStages;        // [kStages] MMA of every cluster CTA -> TMA producer + decode warps
    uint64_t* w_full = bars + 2 * kStages;   // [kStages] TMA (1 arrive + W bytes) -> decode warps
    uint64_t* acc_full = bars + 3 * kStages; // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Linear CTA id -> (tile, K-split).  Within a region the ids are ordered
    // (tile group of CL tiles, split, tile in group) so that a cluster = CL consecutive ids shares
    // the m-tile and the split and covers CL consecutive n-tiles.
    int lin = blockIdx.x;
    int splits = p.splits, tile_base = 0, slot_base = 0;
    const int main_ctas = p.tiles_main * p.splits;
    if (lin >= main_ctas) {
        lin -= main_ctas;
        splits = p.splits_tail;
        tile_base = p.tiles_main;
        slot_base = p.splits > 1 ? main_ctas : 0;
    }
    const int tg = lin / (splits * CL);
    const int rem = lin - tg * (splits * CL);
    const int split = rem / CL;
    const int tile_id = tile_base + tg * CL + (rem - split * CL);
    const int slot0 = slot_base + (tile_id - tile_base) * splits;  // first workspace slot of this tile
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
            // pair, leader: + one relayed arrival for the peer's decode group, and the activation bytes of
            // both CTAs; pair, peer: full[s] only collects its own decode group for the relay (warp 1)
            constexpr int kArrivers = D16 ? kDecodeWarps : kDecodeWarps / 2;  // decode warps per stage
            ptx::mbar_init(&full[s], PAIR ? (ptx::cluster_ctarank() == 0 ? 2 + kArrivers : kArrivers) : 1 + kArrivers);
            ptx::mbar_init(&empty[s], PAIR ? 1 : CL);
            ptx::mbar_init(&w_full[s], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            ptx::tmem_alloc_pair<kTmemCols>(tmem_slot);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc<kTmemCols>(tmem_slot);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if constexpr (CL > 1) {
        ptx::cluster_sync();  // peers' barriers are initialised before anyone multicasts / arrives remotely
    } else {
        __syncthreads();
    }
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
    const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nst; ++i) {
                if constexpr (PAIR) {
                    ptx::mbar_wait_bounded(&empty[s], ph ^ 1u, 1, i);
                } else {
                    ptx::mbar_wait(&empty[s], ph ^ 1u);  // every CTA of the cluster has consumed this stage
                }
                const int k0 = (st_begin + i) * kBK;
                if (p.debug & 32) {
                    ptx::mbar_arrive(&w_full[s]);
                } else {
                    // packed codes of this CTA's 128 output features: bytes [k0/2, k0/2 + 64) of rows n0..n0+127
                    ptx::mbar_arrive_expect_tx(&w_full[s], kWStageBytes);
                    ptx::tma_load_2d(sw + s * kWStageBytes, &tmap_w, &w_full[s], k0 / 2, n0);
                }
                if constexpr (PAIR) {
                    // this CTA stages tokens [rank*MT/2, +MT/2) in its OWN shared memory; the bytes of both
                    // CTAs complete on the leader's barrier, which the leader's producer arms for both
                    const uint32_t lead_full = ptx::mapa_u32(ptx::smem_u32(&full[s]), 0);
                    if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kXStageBytes);
                    uint8_t* dst = sx + s * kXStageBytes;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        ptx::tma_load_2d_pair(dst + h * kXSubBytes, &tmap_x, lead_full, k0 + 64 * h,
                                              m0 + (int)cta_rank * (MT / 2));
                } else if (p.debug & 4) {
                    ptx::mbar_arrive(&full[s]);
                } else {
                    ptx::mbar_arrive_expect_tx(&full[s], kXStageBytes);
                    uint8_t* dst = sx + s * kXStageBytes;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // columns past K are out of bounds for the tensor map: TMA zero-fills them
                        if constexpr (CL == 1) {
                            ptx::tma_load_2d(dst + h * kXSubBytes, &tmap_x, &full[s], k0 + 64 * h, m0);
                        } else {
                            // this CTA fetches rows [rank*MT/CL, +MT/CL) once from L2 and multicasts them
                            // into the same stage of every CTA of the cluster
                            constexpr int kSliceRows = MT / CL;
                            ptx::tma_load_2d_multicast(dst + h * kXSubBytes + cta_rank * (kSliceRows * 128), &tmap_x,
                                                       &full[s], k0 + 64 * h, m0 + (int)cta_rank * kSliceRows,
                                                       kClusterMask);
                        }
                    }
                }
                if (++s == kStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        constexpr uint32_t idesc =
            ptx::make_idesc(/*D=F32*/ 1, TcFmt<T>::kFmt, TcFmt<T>::kFmt, /*M=*/PAIR ? 256 : 128, /*N=*/MT);
        int s = 0;
        uint32_t ph = 0;
        if (PAIR && cta_rank != 0) {
            // peer of a pair: this warp does not issue MMAs (the leader's drive both SMs).  It relays
            // "my decode group has filled TMEM slot s" to the leader's barrier with ONE cluster-scope
            // arrive per stage, so that the 16 decode warps only ever pay CTA-scope arrives
            // (178 -> 148 us at 4096^3 when every decode warp arrived remotely).
            for (int i = 0; i < nst; ++i) {
                ptx::mbar_wait_bounded(&full[s], ph, 6, i);
                if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&full[s]), 0));
                __syncwarp();
                if (++s == kStages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        // pair: only the leader issues; its instructions drive the tensor cores of both SMs
        const int nst_mma = (PAIR && cta_rank != 0) ? 0 : nst;
        for (int i = 0; i < nst_mma; ++i) {
            if constexpr (PAIR) {
                ptx::mbar_wait_bounded(&full[s], ph, 2, i);
            } else {
                ptx::mbar_wait(&full[s], ph);
            }
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint32_t xs = ptx::smem_u32(sx + s * kXStageBytes);
                const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(xs);
                const uint64_t bdesc1 = ptx::make_sw128_kmajor_desc(xs + kXSubBytes);
                const uint32_t a_tmem = tmem_base + kWCol0 + s * 64;
                if (!(p.debug & 8)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // K advances by 16 elements: +8 TMEM columns for A, +32 B (2 x 16 B) inside a sub-tile
                        if constexpr (PAIR)
                            ptx::mma_f16_ts_pair(tmem_base, a_tmem + 8 * j, (j < 4 ? bdesc0 : bdesc1) + 2 * (j & 3),
                                                 idesc, (i | j) != 0 ? 1u : 0u);
                        else
                            ptx::mma_f16_ts(tmem_base, a_tmem + 8 * j, (j < 4 ? bdesc0 : bdesc1) + 2 * (j & 3), idesc,
                                            (i | j) != 0 ? 1u : 0u);
                    }
                }
                if constexpr (PAIR) {
                    ptx::tc_commit_pair(&empty[s], kClusterMask);
                    if (i == nst - 1) ptx::tc_commit_pair(acc_full, kClusterMask);
                } else if constexpr (CL == 1) {
                    ptx::tc_commit(&empty[s]);
                    if (i == nst - 1) ptx::tc_commit(acc_full);
                } else {
                    ptx::tc_commit_multicast(&empty[s], kClusterMask);
                    if (i == nst - 1) ptx::tc_commit(acc_full);
                }
            }
            __syncwarp();
            if (++s == kStages) {
                s = 0;
                ph ^= 1u;
            }
        }
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

        if constexpr (D16) {
            // ---- every decode warp on every stage: thread = (row, 32 consecutive k)
            const int kq = dw >> 2;  // which 32 of the stage's 128 k-elements
            const uint32_t sw_c = (uint32_t)((kq ^ ((row >> 1) & 3)) * 16);
            float wsc1[kScaleDepth];
            auto fetch1 = [&](int j, int stage_idx) {
                wsc1[j] = 0.f;
                const int kb32 = 4 * (st_begin + stage_idx) + kq;  // 32-wide k-block of this thread
                if (stage_idx < nst && n_ok && kb32 * 32 < p.K) wsc1[j] = sc.load((e_row + (long long)kb32 * 32) >> p.log2_bs);
            };
#pragma unroll
            for (int j = 0; j < kScaleDepth; ++j) fetch1(j, j);
            for (int i0 = 0; i0 < nst; i0 += kScaleDepth) {
#pragma unroll
                for (int j = 0; j < kScaleDepth; ++j) {
                    const int i = i0 + j;
                    if (i < nst) {
                        const int s = i % kStages;
                        const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                        const float sc0 = wsc1[j];
                        fetch1(j, i + kScaleDepth);
                        ptx::mbar_wait_bounded(&w_full[s], ph, 23, i);
                        const uint4 q = *reinterpret_cast<const uint4*>(sw + s * kWStageBytes + sw_row + sw_c);
                        uint32_t r[16];
                        DecodeTable tab;
                        build_table<T, QT>(sc0, tab);
                        decode_word(q.x, tab, r + 0);
                        decode_word(q.y, tab, r + 4);
                        decode_word(q.z, tab, r + 8);
                        decode_word(q.w, tab, r + 12);
                        ptx::mbar_wait_bounded(&empty[s], ph ^ 1u, 24, i);
                        ptx::tc_fence_after();
                        ptx::tmem_st_x16(tmem_base + (uint32_t(quarter * 32) << 16) + kWCol0 + s * 64 + kq * 16, r);
                        ptx::tmem_wait_st();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&full[s]);
                    }
                }
            }
        } else {
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

                    if constexpr (PAIR) {
                        ptx::mbar_wait_bounded(&w_full[s], ph, 3, i);
                    } else {
                        ptx::mbar_wait(&w_full[s], ph);  // this stage's codes have landed
                    }
                    const uint8_t* wt = sw + s * kWStageBytes + sw_row;
                    const uint4 q0 = *reinterpret_cast<const uint4*>(wt + sw_c0);
                    const uint4 q1 = *reinterpret_cast<const uint4*>(wt + sw_c1);

                    // 64 codes of row n -> 32 registers of T pairs.  (Rows past N and k-blocks past K are
                    // zero-filled by TMA and carry scale 0: they decode to +-0.)
                    uint32_t r[32];
                    if (p.debug & 1) {
#pragma unroll
                        for (int z = 0; z < 32; ++z) r[z] = q0.x + z;
                    } else {
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
                    }

                    // w_full[s] completing implies the producer saw empty[s]; waiting on it here as well
                    // makes this warp itself an observer of the MMA completion before it overwrites TMEM.
                    if constexpr (PAIR) {
                        ptx::mbar_wait_bounded(&empty[s], ph ^ 1u, 4, i);
                    } else {
                        ptx::mbar_wait(&empty[s], ph ^ 1u);
                    }
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + kWCol0 + s * 64 + half * 32;
                    if (!(p.debug & 2)) {
                        ptx::tmem_st_x32(taddr, r);
                        ptx::tmem_wait_st();
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&full[s]);
                }
            }
        }

        }  // !D16

        // ================================================================== epilogue
        if constexpr (PAIR) {
            ptx::mbar_wait_bounded(acc_full, 0, 5);
        } else {
            ptx::mbar_wait(acc_full, 0);
        }
        ptx::tc_fence_after();

        // this warp: lanes [quarter*32, +32) (= output features), columns [khalf*MT/2, +MT/2)
        constexpr int kColsPerWarp = MT / 4;
        constexpr int kChunk = (kColsPerWarp >= 32) ? 32 : kColsPerWarp;  // 8 (MT=16), 16, 32
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
                    if (n_ok && m < p.M && !(p.debug & 16)) {
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
            // (in split order: deterministic).  All split CTAs of a tile are resident at the same time
            // by construction (the split grid never exceeds one wave), so the short spin cannot deadlock.
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
                while (atomicAdd(arrive, 0) < splits) __nanosleep(64);
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
    if constexpr (CL > 1) {
        ptx::cluster_sync();  // no CTA may exit while peers can still multicast into / arrive on its smem
    } else {
        __syncthreads();
    }
    if (warp == 1) {
        ptx::tc_fence_after();
        if constexpr (PAIR)
            ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
        else
            ptx::tmem_dealloc_dyn(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ persistent variant
// EXPERIMENTAL (BNB_B200_PERSISTENT=1; written at the end of round 1, compiled, not yet run on a
// GPU).  Same tile, same stage pipeline, same numerics as gemm4_tc_kernel for CL = 1 and no split-K,
// but one CTA per SM walks the tiles w = blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * TMEM allocation, barrier set-up and descriptor prefetch happen once per CTA, not per tile;
//   * the stage ring runs THROUGH tile boundaries (global stage counter): while the 16 decode warps
//     drain the accumulator of tile t, the TMA producer already fills the ring for tile t+1;
//   * the accumulator is handed back to the MMA thread as soon as it sits in registers
//     (tcgen05.ld + wait::ld, then one arrive per warp on acc_empty) -- the conversion and the global
//     stores of tile t overlap the first stages of tile t+1.
// Every mbarrier wait is bounded (report + trap after 10 s).
template <typename T, int QT, int MT>
__global__ void __launch_bounds__(kThreads, 1)
    gemm4_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                               const Gemm4Params p) {
    using Cfg = StageCfg<MT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kXSubBytes = Cfg::kXSubBytes;
    constexpr int kXStageBytes = Cfg::kXStageBytes;
    constexpr int kWStageBytes = Cfg::kWStageBytes;
    constexpr uint32_t kTmemCols = Cfg::kTmemCols;
    constexpr uint32_t kWCol0 = Cfg::kWCol0;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sx = smem;
    uint8_t* sw = smem + kStages * kXStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full = bars;                     // [kStages] TMA (1 arrive + X bytes) + 8 decode warps -> MMA
    uint64_t* empty = bars + kStages;          // [kStages] MMA -> TMA producer + decode warps
    uint64_t* w_full = bars + 2 * kStages;     // [kStages] TMA (1 arrive + code bytes) -> decode warps
    uint64_t* acc_full = bars + 3 * kStages;   // MMA -> epilogue            (one phase per tile)
    uint64_t* acc_empty = acc_full + 1;        // 16 epilogue warps -> MMA   (one phase per tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nst = p.kblocks_total;  // stages per tile (no split-K in this variant)
    const int kb64_total = p.K / 64;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full[s], 1 + kDecodeWarps / 2);
            ptx::mbar_init(&empty[s], 1);
            ptx::mbar_init(&w_full[s], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::mbar_init(acc_empty, kDecodeWarps);
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
            uint32_t it = 0;  // global stage counter
            for (int w = blockIdx.x; w < p.tiles_total; w += gridDim.x) {
                const int n0 = (w % p.n_tiles) * kTileN;
                const int m0 = (w / p.n_tiles) * MT;
                for (int i = 0; i < nst; ++i, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    ptx::mbar_wait_bounded(&empty[s], ph ^ 1u, 11, (int)it, w);
                    const int k0 = i * kBK;
                    ptx::mbar_arrive_expect_tx(&w_full[s], kWStageBytes);
                    ptx::tma_load_2d(sw + s * kWStageBytes, &tmap_w, &w_full[s], k0 / 2, n0);
                    ptx::mbar_arrive_expect_tx(&full[s], kXStageBytes);
                    uint8_t* dst = sx + s * kXStageBytes;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        ptx::tma_load_2d(dst + h * kXSubBytes, &tmap_x, &full[s], k0 + 64 * h, m0);
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        constexpr uint32_t idesc = ptx::make_idesc(/*D=F32*/ 1, TcFmt<T>::kFmt, TcFmt<T>::kFmt, /*M=*/128, /*N=*/MT);
        uint32_t it = 0, tc = 0;
        for (int w = blockIdx.x; w < p.tiles_total; w += gridDim.x, ++tc) {
            // the epilogue warps have the previous tile's accumulator in registers
            ptx::mbar_wait_bounded(acc_empty, (tc & 1u) ^ 1u, 12, (int)tc, w);
            ptx::tc_fence_after();
            for (int i = 0; i < nst; ++i, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1u;
                ptx::mbar_wait_bounded(&full[s], ph, 13, (int)it, w);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t xs = ptx::smem_u32(sx + s * kXStageBytes);
                    const uint64_t bdesc0 = ptx::make_sw128_kmajor_desc(xs);
                    const uint64_t bdesc1 = ptx::make_sw128_kmajor_desc(xs + kXSubBytes);
                    const uint32_t a_tmem = tmem_base + kWCol0 + s * 64;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        ptx::mma_f16_ts(tmem_base, a_tmem + 8 * j, (j < 4 ? bdesc0 : bdesc1) + 2 * (j & 3), idesc,
                                        (i | j) != 0 ? 1u : 0u);
                    ptx::tc_commit(&empty[s]);
                    if (i == nst - 1) ptx::tc_commit(acc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================== decode / epilogue warps
        const int dw = warp - 2;
        const int quarter = warp & 3;
        const int grp = dw >> 3;
        const int half = (dw >> 2) & 1;
        const int khalf = dw >> 2;
        const int row = quarter * 32 + lane;
        ScaleSrc sc{p.absmax, p.absmax_8bit, p.absmax_code, p.absmax_offset ? __ldg(p.absmax_offset) : 0.0f};
        const bool two_scales = p.log2_bs == 5;
        const uint32_t sw_row = (uint32_t)row * 64u;
        const uint32_t sw_c0 = (uint32_t)(((2 * half) ^ ((row >> 1) & 3)) * 16);
        const uint32_t sw_c1 = (uint32_t)(((2 * half + 1) ^ ((row >> 1) & 3)) * 16);
        constexpr int kColsPerWarp = MT / 4;  // 64 at MT = 256
        static_assert(kColsPerWarp % 32 == 0, "the persistent variant serves the large-M tiles");
        const int col0 = khalf * kColsPerWarp;
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        T* outp = reinterpret_cast<T*>(p.out);
        const int cnt = (nst - grp + 1) >> 1;  // stages of a tile this group decodes (those with i % 2 == grp)

        uint32_t it0 = 0, tc = 0;  // global index of the tile's first stage; tiles done
        for (int w = blockIdx.x; w < p.tiles_total; w += gridDim.x, it0 += (uint32_t)nst, ++tc) {
            const int n0 = (w % p.n_tiles) * kTileN;
            const int m0 = (w / p.n_tiles) * MT;
            const int n = n0 + row;
            const bool n_ok = n < p.N;
            const long long e_row = (long long)(n_ok ? n : 0) * p.K;

            float wsc[kScaleDepth][2];
            auto fetch = [&](int j, int t) {
                wsc[j][0] = wsc[j][1] = 0.f;
                const int stage_idx = 2 * t + grp;
                const int kb = 2 * stage_idx + half;
                if (stage_idx < nst && n_ok && kb < kb64_total) {
                    const long long e = e_row + (long long)kb * 64;
                    wsc[j][0] = sc.load(e >> p.log2_bs);
                    if (two_scales) wsc[j][1] = sc.load((e + 32) >> p.log2_bs);
                }
            };
#pragma unroll
            for (int j = 0; j < kScaleDepth; ++j) fetch(j, j);

            for (int t0 = 0; t0 < cnt; t0 += kScaleDepth) {
#pragma unroll
                for (int j = 0; j < kScaleDepth; ++j) {
                    const int t = t0 + j;
                    if (t < cnt) {
                        const int i = 2 * t + grp;          // stage inside the tile
                        const uint32_t gi = it0 + (uint32_t)i;  // global stage: ring slot and phase
                        const int s = gi % kStages;
                        const uint32_t ph = (gi / kStages) & 1u;
                        const float sc0 = wsc[j][0], sc1 = wsc[j][1];
                        fetch(j, t + kScaleDepth);

                        ptx::mbar_wait_bounded(&w_full[s], ph, 14, (int)gi, w);
                        const uint8_t* wt = sw + s * kWStageBytes + sw_row;
                        const uint4 q0 = *reinterpret_cast<const uint4*>(wt + sw_c0);
                        const uint4 q1 = *reinterpret_cast<const uint4*>(wt + sw_c1);
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

                        ptx::mbar_wait_bounded(&empty[s], ph ^ 1u, 15, (int)gi, w);
                        ptx::tc_fence_after();
                        const uint32_t taddr =
                            tmem_base + (uint32_t(quarter * 32) << 16) + kWCol0 + s * 64 + half * 32;
                        ptx::tmem_st_x32(taddr, r);
                        ptx::tmem_wait_st();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&full[s]);
                    }
                }
            }

            // ---------------- epilogue of this tile: accumulator -> registers, hand TMEM back, then store
            ptx::mbar_wait_bounded(acc_full, tc & 1u, 16, (int)tc, w);
            ptx::tc_fence_after();
            uint32_t v[kColsPerWarp];
#pragma unroll
            for (int c = 0; c < kColsPerWarp; c += 32) {
                uint32_t v32[32];
                ptx::tmem_ld_x32(lane_addr + col0 + c, v32);
#pragma unroll
                for (int z = 0; z < 32; ++z) v[c + z] = v32[z];
            }
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acc_empty);  // the MMA thread may overwrite the accumulator now

            float bias_v = 0.f;
            if (p.bias != nullptr && n_ok) bias_v = DT<T>::to_f32(reinterpret_cast<const T*>(p.bias)[n]);
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) {
                const int m = m0 + col0 + c;
                if (n_ok && m < p.M) {
                    const T val = DT<T>::from_f32(__uint_as_float(v[c]) + bias_v);
                    const long long idx = (long long)m * p.ldc + n;
                    outp[idx] = val;
                    for (int r2 = 0; r2 < p.n_peers; ++r2) reinterpret_cast<T*>(p.peer_out[r2])[idx] = val;
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

int tail_split_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BNB_B200_TAIL_SPLIT");
        v = e ? atoi(e) : 0;  // measured: helps 2048x14336x4096 (219 -> 208 us), hurts 4096^3 (131 -> 141 us)
    }
    return v;
}

int persistent_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BNB_B200_PERSISTENT");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;  // experimental, not yet measured: off
    }
    return v;
}

int cluster_override() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("BNB_B200_CLUSTER");
        v = e ? atoi(e) : -1;
    }
    return v;
}

template <typename T, int QT, int MT, int CL, bool PAIR = false, bool D16 = false>
bool launch_mt(const T* A, Gemm4Params& p, cudaStream_t stream) {
    using Cfg = StageCfg<MT, PAIR>;
    constexpr size_t smem_bytes = 1024 /*align slack*/ + size_t(Cfg::kStages) * Cfg::kStageBytes + 256 /*barriers*/;
    // the shared-memory opt-in and the resident-CTA count are PER DEVICE (one process may drive several GPUs)
    static bool attr_set[64] = {};
    static int wave_ctas_dev[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    auto kern = gemm4_tc_kernel<T, QT, MT, CL, PAIR, D16>;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
            set_last_error("gemm4_tc smem attr", cudaGetLastError());
            return false;
        }
        attr_set[dev] = true;
    }
    CUtensorMap tmap, tmap_w;
    if (!encode_tmap_2d(&tmap, A, 2, 128, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, (uint32_t)(MT / CL), 64u)) {
        return false;
    }
    // packed codes as a [N, K/2] byte matrix, 128 x 64-byte boxes, 64-byte swizzle
    if (!encode_tmap_2d(&tmap_w, p.B, 1, 64, (uint64_t)p.N, (uint64_t)p.K / 2, (uint64_t)p.K / 2, (uint32_t)kTileN, 64u)) {
        return false;
    }
    const int n_tiles = (p.N + kTileN - 1) / kTileN;
    const int m_tiles = (p.M + MT - 1) / MT;

    // K-splitting.  (a) small problems: a uniform split so that one wave covers the machine;
    // (b) large problems: full tiles for the whole waves, and the partial LAST wave split along K
    // so that it also fills the machine ("tail split": 4096^3 has 512 tiles = 3.46 waves of 148;
    // the last 68 tiles run as 136 half-K CTAs and cost half a wave instead of a full one).
    // Split CTAs exchange fp32 partials through an L2-resident workspace (last arriver reduces,
    // in split order).  Every region size is a multiple of the cluster size.
    // CTAs that can be resident at once: SM count for CL == 1; for clusters the hardware may strand a
    // few SMs (GPC granularity), so ask the occupancy API.
    int& wave_ctas = wave_ctas_dev[dev];
    if (wave_ctas == 0) {
        wave_ctas = device_sm_count();
        if (CL > 1) {
            cudaLaunchConfig_t qc{};
            qc.gridDim = dim3(CL * 64, 1, 1);
            qc.blockDim = dim3(kThreads);
            qc.dynamicSmemBytes = smem_bytes;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CL;
            qa[0].val.clusterDim.y = 1;
            qa[0].val.clusterDim.z = 1;
            qc.attrs = qa;
            qc.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &qc) == cudaSuccess && nclusters > 0) {
                wave_ctas = nclusters * CL;
            } else {
                (void)cudaGetLastError();
            }
        }
    }
    const int sms = wave_ctas;
    const int tiles = n_tiles * m_tiles;

    // Experimental persistent variant (see gemm4_tc_persistent_kernel): more tiles than SMs, no clusters.
    if constexpr (CL == 1 && MT >= 128 && !PAIR && !D16) {
        if (persistent_enabled() && tiles > sms) {
            static bool pattr_set_dev[64] = {};
            bool& pattr_set = pattr_set_dev[dev];
            auto pkern = gemm4_tc_persistent_kernel<T, QT, MT>;
            if (!pattr_set) {
                if (cudaFuncSetAttribute(pkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) !=
                    cudaSuccess) {
                    set_last_error("gemm4_tc persistent smem attr", cudaGetLastError());
                    return false;
                }
                pattr_set = true;
            }
            p.splits = 1;
            p.splits_tail = 1;
            p.tiles_main = tiles;
            p.n_tiles = n_tiles;
            p.tiles_total = tiles;
            p.kblocks_per_split = p.kblocks_total;
            p.ws_partial = nullptr;
            p.ws_counter = nullptr;
            pkern<<<sms, kThreads, smem_bytes, stream>>>(tmap, tmap_w, p);
            BNB200_CHECK_LAUNCH("gemm4_tc_persistent");
            return true;
        }
    }
    const int max_by_k = p.kblocks_total / 2 > 0 ? p.kblocks_total / 2 : 1;  // >= two 128-wide stages per split
    auto clamp_splits = [&](int v) {
        if (v > max_by_k) v = max_by_k;
        if (v > 16) v = 16;
        if (v < 1) v = 1;
        const int per = (p.kblocks_total + v - 1) / v;
        return (p.kblocks_total + per - 1) / per;  // no empty split
    };
    int splits = 1, splits_tail = 1, tiles_main = tiles;
    if (tiles * 2 <= sms) {
        splits = clamp_splits(sms / tiles);
    } else if (tiles > sms && tail_split_enabled()) {
        int full = (tiles / sms) * sms;
        full -= full % CL;
        const int rest = tiles - full;
        if (rest > 0 && rest * 2 <= sms) {
            const int st = clamp_splits(sms / rest);
            if (st > 1) {
                tiles_main = full;
                splits_tail = st;
            }
        }
    }
    p.splits = splits;
    p.splits_tail = splits_tail;
    p.tiles_main = tiles_main;
    p.n_tiles = n_tiles;
    p.tiles_total = tiles;
    p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
    p.ws_partial = nullptr;
    p.ws_counter = nullptr;
    const int split_slots = (splits > 1 ? tiles_main * splits : 0) + (splits_tail > 1 ? (tiles - tiles_main) * splits_tail : 0);
    if (split_slots > 0) {
        size_t bytes = size_t(split_slots) * kTileN * MT * sizeof(float);
        Workspace* ws = get_workspace(stream, bytes, 2 * (size_t)tiles);
        if (ws == nullptr) {
            set_last_error_msg("gemm4_tc: could not allocate the split-K workspace");
            return false;
        }
        p.ws_partial = reinterpret_cast<float*>(ws->ptr);
        p.ws_counter = ws->counters;
    }
    const int grid_ctas = tiles_main * splits + (tiles - tiles_main) * splits_tail;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid_ctas, 1, 1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (split_slots > 0) {
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
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("BNB_B200_DEBUG");
            dbg = e ? atoi(e) : 0;
        }
        p.debug = dbg;
    }

    // Cluster of CL n-tiles sharing one activation tile by TMA multicast: only worth it when the
    // activation tile is the dominant L2 traffic (large M) and the n-tile count divides.
    const int n_tiles = (N + kTileN - 1) / kTileN;
    // Measured on B200 (round 1): multicast does not pay yet -- the kernel is decode(ALU)-bound, not
    // L2-bound, and clusters of 4 strand 16 of the 148 SMs -- so the default is CL = 1
    // (4096^3: 131 us with CL=1, 136 us with CL=2, 138 us with CL=4).  BNB_B200_CLUSTER=2|4 enables it.
    int CL = 1;
    const int ov = cluster_override();
    if (ov == 1 || (ov == 2 && n_tiles % 2 == 0 && MT >= 128) || (ov == 4 && n_tiles % 4 == 0 && MT >= 128)) CL = ov;

    // CTA pairs (cta_group::2) for the large-M tile: BNB_B200_PAIR=1 (measured slower in round 1; off)
    // (both switches are read per call so that one process can A/B them: tools/probe_pair.py)
    const char* pair_e = getenv("BNB_B200_PAIR");
    const bool pair = pair_e != nullptr && pair_e[0] == '1' && MT == 256 && n_tiles % 2 == 0;
    const char* d16_e = getenv("BNB_B200_DECODE16");
    const bool d16 = d16_e != nullptr && d16_e[0] == '1';  // experimental, not yet run: off

#define BNB200_DISPATCH_MT(QT)                                                                                         \
    switch (MT) {                                                                                                      \
    case 16: return launch_mt<T, QT, 16, 1>(A, p, stream);                                                             \
    case 32: return launch_mt<T, QT, 32, 1>(A, p, stream);                                                             \
    case 64: return launch_mt<T, QT, 64, 1>(A, p, stream);                                                             \
    case 128:                                                                                                          \
        if (CL == 4) return launch_mt<T, QT, 128, 4>(A, p, stream);                                                    \
        if (CL == 2) return launch_mt<T, QT, 128, 2>(A, p, stream);                                                    \
        return launch_mt<T, QT, 128, 1>(A, p, stream);                                                                 \
    default:                                                                                                           \
        if (pair && d16) return launch_mt<T, QT, 256, 2, true, true>(A, p, stream);                                    \
        if (pair) return launch_mt<T, QT, 256, 2, true>(A, p, stream);                                                 \
        if (d16 && CL == 1) return launch_mt<T, QT, 256, 1, false, true>(A, p, stream);                                \
        if (CL == 4) return launch_mt<T, QT, 256, 4>(A, p, stream);                                                    \
        if (CL == 2) return launch_mt<T, QT, 256, 2>(A, p, stream);                                                    \
        return launch_mt<T, QT, 256, 1>(A, p, stream);                                                                 \
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
