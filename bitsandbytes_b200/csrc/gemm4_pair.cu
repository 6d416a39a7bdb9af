// gemm4_pair.cu -- NF4/FP4 dequant-fused GEMM for LARGE token counts: a CTA PAIR (cta_group::2) per
// 256-feature x MT-token output tile, decoded weights through tensor memory.
//
// Same contract and numerics as gemm4_tc.cu (reference _ops.py:239-295, gemm_4bit_mma.cuh:99-101;
// on B200 the reference itself takes dequantize + cuBLAS for these shapes, backends/cuda/ops.py:617-623,
// 904-916):   out[m, n] = T( sum_k X[m, k] * rn_T(value(code[n, k]) * scale[n, k / bs])  + bias[n] ).
//
// Why a second kernel.  The one-CTA kernel of gemm4_tc.cu is bound by what an SM can pull out of L2
// (round 1: 72 KB per 1024 MMA cycles = 70 B/clk against the ~45 B/clk/SM the L2 sustains chip-wide).
// Here the two CTAs of a cluster own adjacent 128-feature halves of one 256-feature tile and ONE
// tcgen05.mma.cta_group::2 drives both SMs' tensor cores (UMMA M = 256); the activation tile is split
// between the two shared memories, so each SM ingests HALF the activation bytes.  The pipeline is
// rebuilt around that:
//   * a-stage = 64 k-elements = 32 TMEM columns of decoded weights per CTA.  The ring of a-stages (TMEM
//     A slots + the matching activation slots in shared memory) has one `full` barrier per stage on the
//     LEADER (TMA bytes of both CTAs + the decode warps of BOTH CTAs: all of them arrive with a relaxed
//     cluster-scope arrive -- a cluster-scope RELEASE costs a MEMBAR.ALL.GPU, ~1300 cycles, and there is
//     no generic-proxy memory to publish, only "my tcgen05.st has completed") and one `empty` barrier per
//     stage in each CTA (tcgen05.commit multicast).
//   * packed codes travel in their OWN ring (128 k-elements = 8 KB per stage, 6 stages) fed by a separate
//     producer thread, so the decode warps run ahead of the tensor core instead of starting a stage's
//     decode only after the MMA that frees the matching TMEM slot has retired.
//   * 16 decode warps = 4 groups x 4 warps (one per TMEM lane quarter); group g decodes the a-stages
//     i = g (mod 4): a thread owns one feature row and the 64 codes of the stage (= one quantisation block at
//     the default block size, so the 16-entry table is built once per 64 weights), PRMT-decodes them in
//     registers (decode4.cuh) and writes them with two tcgen05.st.32x32b.x16.
//   * MT = 384 tokens per tile (two N = 192 MMAs per k-step, 384 accumulator columns + 4 x 32 weight
//     columns = the whole 512-column TMEM): a decoded weight feeds 384 MACs instead of 256, which takes
//     the ALU pipe (the PRMT decode, 64 lanes/clk/SM) off the critical path.  MT = 256 / 128 (one MMA per
//     k-step, 8 weight slots) serve smaller token counts.
//   * PERSISTENT: one cluster per SM pair walks the items c, c + #clusters, ...; only the first item pays the
//     launch, the TMEM allocation, the cluster rendezvous and the cold pipeline (~10.8 k cycles, measured).
//     Every item restarts its rings at slot 0 (compile-time slots in the unrolled MMA loop); the mbarrier
//     phases keep counting through a per-slot parity base (ring_base_after).
//   * epilogue: accumulators -> registers -> +bias -> T -> a row-major [tokens][128 features] tile in the
//     TOP four activation slots -> ONE TMA bulk store per decode group (and per destination: the fused all-gather
//     of the column-sharded layer is the same bulk store into each peer GPU's buffer, asynchronous, so the NVLink
//     writes drain while the cluster already runs its next item).  The producers refill the other slots meanwhile
//     and wait for `tile_free` before the staged ones; `acc_empty` lets the MMA thread restart as soon as the
//     accumulators have been read.  2-byte scalar stores took 15-23 k cycles per tile.
//   * the partial last round is split 2 ways along K; the two halves of a tile exchange one fp32 partial
//     through an L2-resident workspace and the LAST ARRIVER (atomic counter, no spinning) adds the other
//     half to its own accumulators -- a + b is commutative, so the result does not depend on who is last.
//   * all shared memory is addressed through 32-bit shared-space addresses off one base (sm100_ptx.cuh "_a"
//     helpers): the decode loop is issue-bound, generic pointers cost it ~90 instructions per stage.
//
// Warp roles (608 threads): warp 0 activation producer (TMA), warp 1 MMA issuer (leader only) + TMEM
// allocator, warps 2..17 decode then epilogue, warp 18 code producer (TMA).
// Measurements behind the numbers in these comments: profiles/r02_pair_trace.md.
#include "common.cuh"
#include "decode4.cuh"
#include "sm100_ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace bnb200 {

// shared with gemm4_tc.cu (split-K scratch per (device, stream))
struct Gemm4Workspace {
    float* partial;
    int* counters;
};
bool gemm4_get_workspace(cudaStream_t stream, size_t partial_bytes, size_t n_counters, Gemm4Workspace* out);

namespace {

constexpr int kTileN = 128;        // features per CTA (TMEM lanes); the pair owns 256
constexpr int kAK = 64;            // a-stage: k-elements per TMEM A slot (32 columns)
constexpr int kCK = 128;           // code stage: k-elements per TMA box of packed codes (64 B per row)
constexpr int kCodeStageBytes = kTileN * (kCK / 2);  // 8 KB
constexpr int kNC = 6;             // code ring depth (12 a-stages of run-ahead for the decode warps)
constexpr int kDecodeWarps = 16;   // 4 groups x 4 warps
constexpr int kThreads = 32 * (2 + kDecodeWarps + 1);
constexpr int kScaleDepth = 4;     // per-thread register ring of scales (stages of one group)
constexpr int kTraceStages = 256;  // TRACE builds: events are kept for the first 256 a-stages of cluster 0
constexpr int kTraceRoles = 10;
constexpr int kMaxOuts = 8;        // local output + up to 7 peer buffers

struct OutMaps {
    CUtensorMap m[kMaxOuts];
};

struct PairParams {
    const uint8_t* B;
    const float* absmax;
    const uint8_t* absmax_8bit;
    const float* absmax_code;
    const float* absmax_offset;
    const void* bias;
    void* out;
    void* peer_out[7];
    int n_peers;
    int tma_out;         // 1: the epilogue stores through the OutMaps tensor maps (ldc % 8 == 0, aligned bases)
    float* ws_partial;   // [split slots][MT columns][128 rows] fp32
    int* ws_counter;     // one per split (tile, CTA rank); zero on entry, reset by the last arriver
    long long* trace;    // TRACE builds only
    int trace_lite;      // 1: only the per-tile landmarks (first MMA, last MMA issued, accumulator ready, epilogue done)
    int M, N, K, ldc;
    int log2_bs;
    int ka_total;        // a-stages in K
    int n_pairs;         // 256-feature tiles along N
    int tiles_total;
    int tiles_main;      // tiles [0, tiles_main) are whole items, the rest is cut into `splits_tail` K ranges
    int n_items;         // tiles_main + (tiles_total - tiles_main) * splits_tail
    int splits_tail;
};

template <typename T> struct TcFmt;
template <> struct TcFmt<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };
template <> struct TcFmt<__half> { static constexpr uint32_t kFmt = 0; };

template <int MT> struct PairCfg {
    static_assert(MT == 128 || MT == 256 || MT == 384, "token tile");
    static constexpr int kNSub = MT == 384 ? 2 : 1;          // MMAs per k-step
    static constexpr int kUmmaN = MT / kNSub;                // 128 / 256 / 192
    static constexpr int kBoxRows = kUmmaN / 2;              // token rows this CTA stages per MMA
    static constexpr int kSubBytes = kBoxRows * 128;         // one [kBoxRows x 64] bf16 box, 128-byte swizzle
    static constexpr int kXStageBytes = kNSub * kSubBytes;   // MT/2 tokens x 128 B
    static constexpr int kNA = (512 - MT) / 32 > 8 ? 8 : (512 - MT) / 32;  // TMEM A-slot ring depth: 4 (MT=384) or 8
    // activation ring (shared memory), decoupled from the A ring: as deep as the budget allows, because the
    // L2 -> SM path answers a tile load only after ~3000 cycles under load (measured) and the bytes in flight
    // per SM set the ingest rate
    static constexpr int kNX = MT == 384 ? 6 : 8;
    static constexpr int kUnroll = MT == 384 ? 12 : 8;       // lcm(kNX, kNA): the MMA loop is unrolled over one ring period
    static constexpr uint32_t kACol0 = MT;                   // D: [0, MT); A slot s: MT + 32 s
    static constexpr size_t kSmemBytes = 1024 + size_t(kNX) * kXStageBytes + size_t(kNC) * kCodeStageBytes + 1024;
    static_assert(4 * kXStageBytes == MT * 256 && kNX > 4, "the output tile is staged in four slots of the activation ring");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(kSubBytes % 1024 == 0, "128-byte swizzle atoms");
    static_assert(kUnroll % kNX == 0 && kUnroll % kNA == 0, "unroll period");
};

template <bool TRACE> __device__ __forceinline__ void trace_ev(const PairParams& p, int role, int i, bool landmark = false) {
    if constexpr (TRACE) {
        if (p.trace != nullptr && blockIdx.x < 2 && i < kTraceStages && (landmark || !p.trace_lite))
            p.trace[((long long)blockIdx.x * kTraceRoles + role) * kTraceStages + i] = clock64();
    }
}
// wall clock (ns) next to a cycle stamp, to convert cycles into time
template <bool TRACE> __device__ __forceinline__ void trace_ns(const PairParams& p, int role, int i) {
    if constexpr (TRACE) {
        if (p.trace != nullptr && blockIdx.x < 2) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[((long long)blockIdx.x * kTraceRoles + role) * kTraceStages + i] = (long long)t;
        }
    }
}

// per-cluster timeline (TRACE builds, lite mode):
//   [kTraceTimeline + 8 cid + {0: start ns, 1: end ns, 2: SM id, 3: cycles entry -> first MMA, 4..7: end ns of item 0..3}]
constexpr int kTraceTimeline = 2 * kTraceRoles * kTraceStages;
constexpr int kTimelineSlots = 8;
template <bool TRACE> __device__ __forceinline__ void trace_cluster(const PairParams& p, int slot, long long v = -1) {
    if constexpr (TRACE) {
        if (p.trace != nullptr && p.trace_lite && (blockIdx.x >> 1) < 512 && slot < kTimelineSlots) {
            if (v < 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                v = (long long)t;
            }
            p.trace[kTraceTimeline + kTimelineSlots * (blockIdx.x >> 1) + slot] = v;
        }
    }
}

// One unit of work of a cluster: an output tile (256 features x MT tokens), or one K half of a tile of the
// partial last round.
struct PairItem {
    int tile, split, splits;
    int n0, m0;          // this CTA's first feature / the tile's first token
    int st_begin, nst;   // K range in a-stages (nst >= 1: the host only splits when every split gets work)
};
template <int MT> __device__ __forceinline__ PairItem pair_item(const PairParams& p, int w, uint32_t cta_rank) {
    PairItem it;
    it.split = 0;
    it.splits = 1;
    if (w < p.tiles_main) {
        it.tile = w;
    } else {
        const int r = w - p.tiles_main;
        it.splits = p.splits_tail;
        it.tile = p.tiles_main + r / it.splits;
        it.split = r - (r / it.splits) * it.splits;
    }
    it.n0 = (it.tile % p.n_pairs) * (2 * kTileN) + (int)cta_rank * kTileN;
    it.m0 = (it.tile / p.n_pairs) * MT;
    // boundaries on even a-stages (= whole code stages)
    int per = (p.ka_total + it.splits - 1) / it.splits;
    per += per & 1;
    it.st_begin = it.split * per;
    int st_end = it.st_begin + per;
    if (st_end > p.ka_total) st_end = p.ka_total;
    it.nst = st_end - it.st_begin;
    return it;
}
// Every item restarts its rings at slot 0 (so the MMA loop can be unrolled with compile-time slots); the
// mbarrier phases of course keep counting.  `base` holds, per slot, the parity of the uses completed by the
// earlier items: an item of n stages uses slot x (n - x + R - 1) / R times.
template <int R> __device__ __forceinline__ uint32_t ring_base_after(uint32_t base, int n) {
    const int fullr = n / R, rem = n - fullr * R;
    if (fullr & 1) base ^= (1u << R) - 1u;
    return base ^ ((1u << rem) - 1u);
}

template <typename T, int QT, int MT, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
    gemm4_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ OutMaps omaps, const PairParams p) {
    using Cfg = PairCfg<MT>;
    constexpr int kNA = Cfg::kNA;
    constexpr int kNX = Cfg::kNX;
    constexpr int kXStageBytes = Cfg::kXStageBytes;
    constexpr uint32_t kACol0 = Cfg::kACol0;
    static_assert(kNA == 4 || (kNA == 8 && kNX == 8), "decode-side ring bookkeeping");

    // Everything in shared memory is addressed by 32-bit shared-space addresses off ONE base (sm100_ptx.cuh, "_a"
    // helpers): no generic pointers, no cvta, no 64-bit address arithmetic in the hot loops.
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sx = sbase;                                  // [kNX][kNSub][kBoxRows x 128 B]
    const uint32_t sw = sx + kNX * kXStageBytes;                // [kNC][128 x 64 B]
    // the output tile ([MT token rows][128 features] of T = 4 activation slots) is staged in the TOP four slots of
    // the activation ring: the producer may refill the lower ones for the next item while the epilogue runs
    const uint32_t so = sx + (kNX - 4) * kXStageBytes;
    // ONE barrier pair per a-stage, indexed by the activation slot i % kNX (the TMEM A slot is i % kNA):
    //   full[i % kNX]   the stage's activation bytes (both CTAs) + the 8 decode warps (both CTAs) -> MMA  (leader's)
    //   empty[i % kNX]  MMA(i) has retired -> the activation producer (slot reusable at stage i + kNX) AND the
    //                   decode warps (TMEM A slot reusable at stage i + kNA <= i + kNX)
    // so the MMA thread pays one wait and one commit per stage while the activation ring is deeper than the A ring.
    const uint32_t full = sw + kNC * kCodeStageBytes;           // [kNX]
    const uint32_t empty = full + 8 * kNX;                      // [kNX]
    const uint32_t c_full = empty + 8 * kNX;                    // [kNC]
    const uint32_t c_empty = c_full + 8 * kNC;                  // [kNC]
    const uint32_t acc_full = c_empty + 8 * kNC;   // accumulators of the item complete -> the decode warps' epilogue
    const uint32_t acc_empty = acc_full + 8;       // accumulators read out (32 warps of both CTAs) -> MMA of the next item (leader's)
    const uint32_t tile_free = acc_empty + 8;      // the 4 decode groups' bulk stores have read the staged tile -> activation producer
    const uint32_t tmem_slot = tile_free + 8;
    const uint32_t s_flag = tmem_slot + 4;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = ptx::cluster_ctarank();
    const bool leader = cta_rank == 0;
    long long t_entry = 0;
    if constexpr (TRACE) {
        t_entry = clock64();
        if (threadIdx.x == 32 && leader) {
            trace_cluster<TRACE>(p, 0);
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            trace_cluster<TRACE>(p, 2, (long long)smid);
        }
    }

    // Persistent: cluster c works on the items c, c + #clusters, ... (whole tiles first, the K halves of the
    // partial last round at the end).  Only the first item pays the launch, the TMEM allocation, the cluster
    // rendezvous and the cold pipeline (~10 k cycles, measured: profiles/r02_pair_trace.md); on the later ones
    // the producers are already ahead when the epilogue of the previous item ends.
    const int cid = blockIdx.x >> 1;
    const int ncl = gridDim.x >> 1;
    const PairItem first = pair_item<MT>(p, cid, cta_rank);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_x);
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < kNX; ++s) {
            // leader: its producer's expect_tx arrival (bytes of both CTAs) + the 4 decode warps of each CTA
            ptx::mbar_init_a(full + 8 * s, 9);
            ptx::mbar_init_a(empty + 8 * s, 1);
        }
        ptx::mbar_init_a(acc_full, 1);
        ptx::mbar_init_a(acc_empty, 2 * kDecodeWarps);
        ptx::mbar_init_a(tile_free, 4);
        ptx::fence_barrier_init();
    }
    // The code ring is CTA-local (its barriers are never touched by the peer), so its producer initialises it and
    // starts the first kNC loads BEFORE the cluster rendezvous: their ~3000-cycle L2 latency then overlaps the
    // TMEM allocation and the cluster barrier instead of following them.
    const int ncs_first = (first.nst + 1) >> 1;
    const int ncs_early = ncs_first < kNC ? ncs_first : kNC;
    if (warp == 18 && ptx::elect_one()) {
        for (int s = 0; s < kNC; ++s) {
            ptx::mbar_init_a(c_full + 8 * s, 1);
            ptx::mbar_init_a(c_empty + 8 * s, 8);  // the 2 x 4 decode warps that read a code stage
        }
        ptx::fence_barrier_init();
        for (int j = 0; j < ncs_early; ++j) {
            ptx::mbar_arrive_expect_tx_a(c_full + 8 * j, kCodeStageBytes);
            ptx::tma_load_2d_a(sw + j * kCodeStageBytes, &tmap_w, c_full + 8 * j, ((first.st_begin + 2 * j) * kAK) / 2,
                               first.n0);
        }
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair_a<512>(tmem_slot);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ptx::lds_u32(tmem_slot);

    if (warp == 0) {
        // ================================================================== activation producer
        // (elect.sync, not `lane == 0`: ptxas then knows the region is single-threaded and keeps the TMA / MMA operands
        // in uniform registers instead of wrapping every instruction in an ELECT + R2UR loop -- measured: ~94 cycles
        // per tcgen05.mma and ~800 per stage with the loop, profiles/r02_pair_trace.md)
        if (ptx::elect_one()) {
            const uint32_t lead_full0 = ptx::mapa_u32(full, 0);
            uint32_t base = 0;
            int item = 0;
            for (int w = cid; w < p.n_items; w += ncl, ++item) {
                const PairItem it = pair_item<MT>(p, w, cta_rank);
                int s = 0;
                uint32_t ph = 0;
                bool tile_pending = item > 0;  // the previous item's output tile may still sit in slots kNX-4 .. kNX-1
                for (int i = 0; i < it.nst; ++i) {
                    if (tile_pending && s >= kNX - 4) {
                        ptx::mbar_wait_a(tile_free, (uint32_t)(item - 1) & 1u, 9, item);
                        tile_pending = false;
                    }
                    // the previous use of the slot (this item's stage i - kNX, or the previous item's) has retired
                    ptx::mbar_wait_a(empty + 8 * s, ((base >> s) & 1u) ^ ph ^ 1u, 1, i);
                    if (w == cid) trace_ev<TRACE>(p, 0, i);
                    const int k0 = (it.st_begin + i) * kAK;
                    if (leader) ptx::mbar_arrive_expect_tx_a(full + 8 * s, 2 * kXStageBytes);
                    const uint32_t dst = sx + s * kXStageBytes;
#pragma unroll
                    for (int sub = 0; sub < Cfg::kNSub; ++sub)
                        ptx::tma_load_2d_pair_a(dst + sub * Cfg::kSubBytes, &tmap_x, lead_full0 + 8u * s, k0,
                                                it.m0 + sub * Cfg::kUmmaN + (int)cta_rank * Cfg::kBoxRows);
                    if (++s == kNX) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                // (every phase of tile_free is observed, also by an item too short to reach the staging slots)
                if (tile_pending) ptx::mbar_wait_a(tile_free, (uint32_t)(item - 1) & 1u, 9, item);
                base = ring_base_after<kNX>(base, it.nst);
            }
        }
    } else if (warp == 18) {
        // ================================================================== code producer
        if (ptx::elect_one()) {
            uint32_t base = 0;
            for (int w = cid; w < p.n_items; w += ncl) {
                const PairItem it = pair_item<MT>(p, w, cta_rank);
                const int ncs = (it.nst + 1) >> 1;
                const int j0 = w == cid ? ncs_early : 0;  // the first item's first loads are already in flight
                int cs = j0 % kNC;
                uint32_t ph = (uint32_t)(j0 / kNC) & 1u;
                for (int j = j0; j < ncs; ++j) {
                    ptx::mbar_wait_a(c_empty + 8 * cs, ((base >> cs) & 1u) ^ ph ^ 1u, 2, j);
                    ptx::mbar_arrive_expect_tx_a(c_full + 8 * cs, kCodeStageBytes);
                    // bytes [k/2, k/2 + 64) of rows n0 .. n0+127 (rows past N / bytes past K/2: zero-filled)
                    ptx::tma_load_2d_a(sw + cs * kCodeStageBytes, &tmap_w, c_full + 8 * cs,
                                       ((it.st_begin + 2 * j) * kAK) / 2, it.n0);
                    if (++cs == kNC) {
                        cs = 0;
                        ph ^= 1u;
                    }
                }
                base = ring_base_after<kNC>(base, ncs);
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (leader CTA, one thread)
        // Measured on B200 (round 2 traces, profiles/r02_pair_trace.md): the tensor core accepts a tcgen05.mma only
        // about one instruction ahead of the one it executes (an N = 192 pair MMA is accepted every ~95 cycles), a
        // multicast tcgen05.commit costs this thread ~100 cycles and an mbarrier probe ~70.  Anything done BETWEEN two
        // stages therefore drains the pipe.  So the loop is software-pipelined: the commit that releases stage i-1 is
        // issued after the second MMA of stage i (it then also covers those two: the slots are released ~200 cycles
        // later, which the rings absorb), and the barrier of stage i+1 is probed (one non-blocking try_wait) before
        // the last MMAs of stage i -- the operands are ready long before, the probe succeeds and its latency hides
        // under the queued MMAs; only a failed probe falls back to a blocking wait.
        if (leader && ptx::elect_one()) {
            constexpr uint32_t idesc =
                ptx::make_idesc(/*D=F32*/ 1, TcFmt<T>::kFmt, TcFmt<T>::kFmt, /*M=*/256, /*N=*/Cfg::kUmmaN);
            constexpr int kMmas = (kAK / 16) * Cfg::kNSub;  // 4 or 8 per a-stage
            constexpr int kU = Cfg::kUnroll;                // ring period: a multiple of kNX and of kNA
            uint32_t base = 0;
            int item = 0;
            for (int w = cid; w < p.n_items; w += ncl, ++item) {
                const PairItem it = pair_item<MT>(p, w, cta_rank);
                const int nst = it.nst;
                const bool tr = TRACE && item == 0;
                // One a-stage.  `xs`, `nxs` (activation / barrier slots of this and the next stage), `pxs` (previous)
                // and `s` (TMEM A slot) are compile-time constants in the unrolled main loop below: the single issuing
                // thread then executes nothing but the tcgen05 instructions, one commit and one probe per stage.  (With
                // run-time slot indices the loop body was ~45 dependent instructions, five of them R2UR: ~650 cycles
                // per stage, measured, against 512 cycles of MMA.)
                auto stage = [&](int i, int xs, int nxs, int pxs, int s, uint32_t nxph, bool more) {
                    if (tr) trace_ev<TRACE>(p, 1, i, i == 0);
                    if constexpr (TRACE) {
                        if (tr && i == 0) {
                            trace_ns<TRACE>(p, 9, 40);
                            trace_cluster<TRACE>(p, 3, clock64() - t_entry);
                        }
                    }
                    ptx::tc_fence_after();
                    bool ok = false;
                    // (opaque to the optimiser: otherwise the 96 loop-invariant 64-bit descriptors of the unrolled ring
                    // period are hoisted into vector registers, spilled, and fed back through ~20 R2UR per stage)
                    uint32_t xa = sx + xs * kXStageBytes;
                    uint32_t a_tmem = tmem_base + kACol0 + s * 32;
                    asm volatile("" : "+r"(xa), "+r"(a_tmem));
#pragma unroll
                    for (int j = 0; j < kMmas; ++j) {
                        const int k = j / Cfg::kNSub, sub = j % Cfg::kNSub;
                        // K advances by 16 elements: +8 TMEM columns of A, +32 B inside the 128-byte swizzle row
                        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(xa + sub * Cfg::kSubBytes) + 2 * k;
                        ptx::mma_f16_ts_pair(tmem_base + sub * Cfg::kUmmaN, a_tmem + 8 * k, bdesc, idesc,
                                             (i | k) != 0 ? 1u : 0u);
                        // stage i-1 (and the two MMAs above) retired -> its slots are free
                        if (j == 1 && i > 0) ptx::tc_commit_pair_a(empty + 8 * pxs, 0x3);
                        if (j == kMmas - 2 && more) ok = ptx::mbar_try_wait_a(full + 8 * nxs, nxph);
                    }
                    if (tr) trace_ev<TRACE>(p, 2, i, i == nst - 1);
                    if (more && !ok) ptx::mbar_wait_slow_a(full + 8 * nxs, nxph, 4, i + 1);
                    if (tr) trace_ev<TRACE>(p, 7, i);
                };
                // the accumulators of the previous item have been read out by both CTAs' epilogues
                if (item > 0) ptx::mbar_wait_a(acc_empty, (uint32_t)(item - 1) & 1u, 8, item);
                ptx::mbar_wait_a(full, base & 1u, 4, 0);
                int i = 0;
                for (; i + kU <= nst; i += kU) {
                    const uint32_t par = (uint32_t)(i / kNX) & 1u;  // parity of the activation ring at the body's start
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const int xs = u % kNX, nxs = (u + 1) % kNX, pxs = (u + kNX - 1) % kNX, s = u % kNA;
                        const uint32_t nxph = par ^ (uint32_t)(((u + 1) / kNX) & 1) ^ ((base >> nxs) & 1u);
                        stage(i + u, xs, nxs, pxs, s, nxph, u + 1 < kU ? true : (i + kU < nst));
                    }
                }
                // remaining stages (fewer than a ring period): run-time slot indices
                for (; i < nst; ++i) {
                    const int xs = i % kNX, nxs = (i + 1) % kNX, pxs = (i + kNX - 1) % kNX, s = i % kNA;
                    stage(i, xs, nxs, pxs, s, ((uint32_t)((i + 1) / kNX) & 1u) ^ ((base >> nxs) & 1u), i + 1 < nst);
                }
                ptx::tc_commit_pair_a(empty + 8 * ((nst - 1) % kNX), 0x3);
                ptx::tc_commit_pair_a(acc_full, 0x3);
                base = ring_base_after<kNX>(base, nst);
            }
        }
        __syncwarp();
    } else {
        // ================================================================== decode warps
        // The decode is ISSUE bound: each scheduler hosts one warp of every group, so a scheduler issues one warp's
        // whole stage per a-stage (~500 instructions of which 270 are the PRMT decode and the table build).  Hence
        // no div/mod by the ring sizes (incremental slots), no generic addresses, no branches around the arrives.
        const int dw = warp - 2;        // 0..15
        const int quarter = warp & 3;   // TMEM lane quarter this warp may touch
        const int grp = dw >> 2;        // decodes the a-stages i with i % 4 == grp
        const int row = quarter * 32 + lane;
        ScaleSrc sc{p.absmax, p.absmax_8bit, p.absmax_code, p.absmax_offset ? __ldg(p.absmax_offset) : 0.0f};
        const bool two_scales = p.log2_bs == 5;
        // 64-byte swizzle: 16-byte chunk c of code row r sits at c ^ ((r >> 1) & 3); a stage reads chunks {0,1} or {2,3}
        const uint32_t code_off = sw + (uint32_t)row * 64u + (uint32_t)(((row >> 1) & 3) << 4);
        const bool tracer0 = TRACE && quarter == 0 && lane == 0;
        // the leader's barriers through shared::cluster addresses: the same (relaxed) arrive from both CTAs
        const uint32_t lead_full0 = ptx::mapa_u32(full, 0);
        const uint32_t lead_acc_empty = ptx::mapa_u32(acc_empty, 0);
        const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
        // this warp in the epilogue: lanes [quarter*32, +32) (= output features), columns [grp*MT/4, +MT/4) (= tokens)
        constexpr int kColsPerWarp = MT / 4;  // 32, 64 or 96
        constexpr int kChunks = kColsPerWarp / 32;
        const int col0 = grp * kColsPerWarp;
        T* outp = reinterpret_cast<T*>(p.out);

        // per-thread register ring of scales (the stages this group owns, kScaleDepth ahead)
        float wsc[kScaleDepth][2];
        PairItem it = first;
        long long e0 = 0;   // element index of (this thread's row, first k of the item); -1: row past N
        auto item_base = [&](const PairItem& q) {
            const int n = q.n0 + row;
            e0 = n < p.N ? (long long)n * p.K + (long long)q.st_begin * kAK : -1;
        };
        auto fetch = [&](const PairItem& q, int j, int t) {
            wsc[j][0] = wsc[j][1] = 0.f;
            const int i = 4 * t + grp;
            if (i < q.nst && e0 >= 0) {
                const long long e = e0 + (long long)i * kAK;
                wsc[j][0] = sc.load(e >> p.log2_bs);
                if (two_scales) wsc[j][1] = sc.load((e + 32) >> p.log2_bs);
            }
        };
        item_base(it);
#pragma unroll
        for (int j = 0; j < kScaleDepth; ++j) fetch(it, j, j);

        uint32_t base_x = 0, base_c = 0;
        int item = 0;
        for (int w = cid; w < p.n_items; w += ncl, ++item) {
            const int nst = it.nst;
            const bool tracer = tracer0 && item == 0;
            const int cnt = (nst - grp + 3) >> 2;  // a-stages this group owns
            // ring positions of stage i = grp + 4 t, advanced incrementally
            int xs = grp;                 // i % kNX (kNX > 4 > grp)
            uint32_t xph = 0;             // (i / kNX) & 1
            int pxs = 0;                  // slot / parity of stage i - 4 (kNA == 4: the stage whose MMA frees A slot)
            uint32_t pxph = 0;
            int cs = grp >> 1;            // (i / 2) % kNC
            uint32_t cph = 0;             // (i / 2 / kNC) & 1
            int i = grp;
            for (int t0 = 0; t0 < cnt; t0 += kScaleDepth) {
#pragma unroll
                for (int j = 0; j < kScaleDepth; ++j) {
                    const int t = t0 + j;
                    if (t < cnt) {
                        const float sc0 = wsc[j][0], sc1 = wsc[j][1];
                        fetch(it, j, t + kScaleDepth);

                        ptx::mbar_wait_a(c_full + 8 * cs, cph ^ ((base_c >> cs) & 1u), 5, i);
                        if (tracer) trace_ev<TRACE>(p, 3, i);
                        // which 32 bytes of the 64-byte code row: chunks {0,1} (even stage) or {2,3} (odd)
                        const uint32_t wt = (code_off + cs * kCodeStageBytes) ^ ((uint32_t)(i & 1) << 5);
                        const uint4 q0 = ptx::lds128(wt);
                        const uint4 q1 = ptx::lds128(wt ^ 16u);

                        // First half (32 codes) -> 16 registers -> TMEM while the second half is being decoded: the
                        // tcgen05.st and its completion latency (~500 cycles per warp, measured) overlap the PRMT work.
                        uint32_t ra[16], rb[16];
                        DecodeTable tab;
                        build_table<T, QT>(sc0, tab);
                        decode_word(q0.x, tab, ra + 0);
                        decode_word(q0.y, tab, ra + 4);
                        decode_word(q0.z, tab, ra + 8);
                        decode_word(q0.w, tab, ra + 12);
                        if (tracer) trace_ev<TRACE>(p, 4, i);
                        // TMEM A slot i % kNA was last read by MMA(i - kNA), which commits to empty[(i - kNA) % kNX]; the
                        // first kNA stages of an item follow the epilogue of the previous one (everything retired)
                        if (i >= kNA) {
                            if constexpr (kNA == 4)
                                ptx::mbar_wait_a(empty + 8 * pxs, pxph ^ ((base_x >> pxs) & 1u), 6, i);
                            else  // kNA == kNX == 8: same slot, one ring turn earlier
                                ptx::mbar_wait_a(empty + 8 * xs, xph ^ 1u ^ ((base_x >> xs) & 1u), 6, i);
                        }
                        if (tracer) trace_ev<TRACE>(p, 5, i);
                        ptx::tc_fence_after();
                        const uint32_t taddr = lane_addr + kACol0 + (uint32_t)(kNA == 4 ? grp : (i & 7)) * 32u;
                        ptx::tmem_st_x16(taddr, ra);
                        if (two_scales) build_table<T, QT>(sc1, tab);
                        decode_word(q1.x, tab, rb + 0);
                        decode_word(q1.y, tab, rb + 4);
                        decode_word(q1.z, tab, rb + 8);
                        decode_word(q1.w, tab, rb + 12);
                        // the codes are in registers (the decode consumed them): hand the code stage back.  An item with
                        // an odd number of a-stages leaves the second half of its last code stage unused: the group that
                        // decodes the last stage arrives for the missing one too (the producer reuses the ring in the
                        // next item).
                        __syncwarp();
                        if (lane == 0) {
                            ptx::mbar_arrive_a(c_empty + 8 * cs);
                            if (i == nst - 1 && (i & 1) == 0) ptx::mbar_arrive_a(c_empty + 8 * cs);
                        }
                        ptx::tmem_st_x16(taddr + 16, rb);
                        ptx::tmem_wait_st();
                        ptx::tc_fence_before();
                        __syncwarp();
                        // "this warp's 32 rows of the A slot are in TMEM": the payload is tensor memory, completed by
                        // wait::st above -- no generic-proxy data to release, hence the relaxed (cluster) arrive
                        if (lane == 0) ptx::mbar_arrive_cluster_relaxed(lead_full0 + 8u * xs);
                        if (tracer) trace_ev<TRACE>(p, 6, i);

                        // next owned stage: i + 4
                        pxs = xs;
                        pxph = xph;
                        i += 4;
                        xs += 4;
                        if (xs >= kNX) {
                            xs -= kNX;
                            xph ^= 1u;
                        }
                        cs += 2;
                        if (cs >= kNC) {
                            cs -= kNC;
                            cph ^= 1u;
                        }
                    }
                }
            }

            // the next item's scales travel while this item's epilogue runs
            const PairItem cur = it;
            if (w + ncl < p.n_items) {
                it = pair_item<MT>(p, w + ncl, cta_rank);
                item_base(it);
#pragma unroll
                for (int j = 0; j < kScaleDepth; ++j) fetch(it, j, j);
            }

            // ================================================================== epilogue
            ptx::mbar_wait_a(acc_full, (uint32_t)item & 1u, 7, item);
            ptx::tc_fence_after();
            if (tracer && grp == 0) trace_ev<TRACE>(p, 8, 0, true);

            const int n = cur.n0 + row;
            const bool n_ok = n < p.N;
            float bias_v = 0.f;
            if (p.bias != nullptr && n_ok) bias_v = DT<T>::to_f32(reinterpret_cast<const T*>(p.bias)[n]);

            bool finish = true;            // this CTA writes the output tile (false: first split to arrive)
            const float* other = nullptr;  // the other split's fp32 partial ([column][row]) to add, or NULL
            if (cur.splits == 2) {
                // ---- split K: publish the fp32 partial tile ([column][row]: coalesced both ways) and count in on the
                // (tile, CTA) counter.  The LAST of the two splits to arrive adds the other's partial to its own
                // accumulators (still in TMEM) and writes the tile; the first one is done.  Nobody waits.
                const int tt = (cur.tile - p.tiles_main) * 2 + (int)cta_rank;
                float* ws_tile = p.ws_partial + (long long)tt * 2 * (kTileN * MT);
                float* my = ws_tile + (long long)cur.split * (kTileN * MT);
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp; c += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_x32(lane_addr + col0 + c, v);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int t = 0; t < 32; ++t) __stcg(my + (col0 + c + t) * kTileN + row, __uint_as_float(v[t]));
                }
                __threadfence();
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (threadIdx.x == 64) {
                    const int old = atomicAdd(p.ws_counter + tt, 1);
                    ptx::sts_u32(s_flag, (uint32_t)old);
                    if (old == 1) p.ws_counter[tt] = 0;  // both have arrived: reset for the next launch
                    __threadfence();
                }
                asm volatile("bar.sync 1, 512;" ::: "memory");
                finish = ptx::lds_u32(s_flag) == 1u;
                other = ws_tile + (long long)(cur.split ^ 1) * (kTileN * MT);
            }

            if (finish && p.tma_out) {
                // tile in shared memory: [MT token rows][128 features] of T, 256 B per row; the 4 warps of this decode
                // group own token rows [col0, col0 + MT/4) and send them with ONE bulk store per destination.  The
                // other split's partial is fetched one chunk AHEAD (32 loads in flight while the previous chunk is
                // converted and stored): issued one by one behind the shared-memory stores the loads cost an L2 round
                // trip EACH (~50 k cycles per tile, measured).
                const uint32_t trow = so + (uint32_t)(col0 * kTileN + row) * 2u;  // (token col0, this feature)
                float ov[2][32];
                auto load_other = [&](float* dst, int c) {
#pragma unroll
                    for (int t = 0; t < 32; ++t) dst[t] = __ldcg(other + (col0 + c + t) * kTileN + row);
                };
                if (other != nullptr) load_other(ov[0], 0);
#pragma unroll
                for (int ci = 0; ci < kChunks; ++ci) {
                    const int c = ci * 32;
                    uint32_t v[32];
                    ptx::tmem_ld_x32(lane_addr + col0 + c, v);
                    if (other != nullptr && ci + 1 < kChunks) load_other(ov[(ci + 1) & 1], c + 32);
                    ptx::tmem_wait_ld();
                    if (ci == kChunks - 1) {
                        // accumulators read out: the MMA thread may start the next item
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive_cluster_relaxed(lead_acc_empty);
                    }
                    if (other != nullptr) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + ov[ci & 1][t]);
                    }
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const T val = DT<T>::from_f32(__uint_as_float(v[t]) + bias_v);
                        ptx::sts_b16(trow + (uint32_t)(c + t) * (kTileN * 2), reinterpret_cast<const uint16_t&>(val));
                    }
                }
                ptx::fence_proxy_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                if (quarter == 2 && ptx::elect_one()) {  // warp 2 + 4 grp (the first warp of the group)
                    if (cur.m0 + col0 < p.M) {
                        for (int d = 0; d <= p.n_peers; ++d)
                            ptx::tma_store_2d_a(&omaps.m[d], so + (uint32_t)(col0 * kTileN) * 2u, cur.n0, cur.m0 + col0);
                        ptx::tma_store_commit();
                        ptx::tma_store_wait_read();  // shared memory must outlive the reads (not the global writes)
                    }
                    ptx::mbar_arrive_a(tile_free);  // the activation producer may refill this group's part of the ring
                }
            } else {
                if (finish) {
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp; c += 32) {
                        uint32_t v[32];
                        ptx::tmem_ld_x32(lane_addr + col0 + c, v);
                        ptx::tmem_wait_ld();
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const int m = cur.m0 + col0 + c + t;
                            if (n_ok && m < p.M) {
                                float f = __uint_as_float(v[t]);
                                if (other != nullptr) f += __ldcg(other + (col0 + c + t) * kTileN + row);  // (rare path: unaligned ldc)
                                const T val = DT<T>::from_f32(f + bias_v);
                                const long long idx = (long long)m * p.ldc + n;
                                outp[idx] = val;
                                for (int r2 = 0; r2 < p.n_peers; ++r2) reinterpret_cast<T*>(p.peer_out[r2])[idx] = val;
                            }
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive_cluster_relaxed(lead_acc_empty);
                    if (quarter == 2) ptx::mbar_arrive_a(tile_free);  // nothing staged
                }
            }
            if (tracer && grp == 0) {
                trace_ev<TRACE>(p, 9, 0, true);
                trace_ns<TRACE>(p, 9, 41);
            }
            if constexpr (TRACE) {
                if (tracer0 && grp == 0 && leader) trace_cluster<TRACE>(p, 4 + item);
            }
            base_x = ring_base_after<kNX>(base_x, nst);
            base_c = ring_base_after<kNC>(base_c, (nst + 1) >> 1);
        }
    }

    // ------------------------------------------------------------------ teardown
    ptx::tc_fence_before();
    ptx::cluster_sync();  // no CTA may exit while its peer can still arrive on / commit into its shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, 512);
        if (leader && lane == 0) trace_cluster<TRACE>(p, 1);
    }
}

// ------------------------------------------------------------------ host side
// Tensor maps are pure functions of (base, shape, box): a small per-process cache keeps
// cuTensorMapEncodeTiled (a driver call) off the per-call host path.
struct TmapKey {
    const void* base;
    uint64_t rows, cols, stride;
    uint32_t box_rows, box_cols;
    int elem, swz;
};
struct TmapEntry {
    TmapKey key;
    CUtensorMap map;
    bool used;
};
constexpr int kTmapCache = 32;
TmapEntry g_tmaps[kTmapCache];
int g_tmap_next = 0;
std::mutex g_tmap_mu;

bool cached_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle, uint64_t rows, uint64_t cols,
                 uint64_t stride, uint32_t box_rows, uint32_t box_cols) {
    const TmapKey k{base, rows, cols, stride, box_rows, box_cols, elem_bytes, swizzle};
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    for (int i = 0; i < kTmapCache; ++i) {
        const TmapEntry& e = g_tmaps[i];
        if (e.used && e.key.base == k.base && e.key.rows == k.rows && e.key.cols == k.cols && e.key.stride == k.stride &&
            e.key.box_rows == k.box_rows && e.key.box_cols == k.box_cols && e.key.elem == k.elem && e.key.swz == k.swz) {
            *out = e.map;
            return true;
        }
    }
    if (!encode_tmap_2d(out, base, elem_bytes, swizzle, rows, cols, stride, box_rows, box_cols)) return false;
    TmapEntry& e = g_tmaps[g_tmap_next];
    g_tmap_next = (g_tmap_next + 1) % kTmapCache;
    e.key = k;
    e.map = *out;
    e.used = true;
    return true;
}

template <typename T, int QT, int MT, bool TRACE>
bool launch_pair_mt(const T* A, PairParams& p, cudaStream_t stream, int force_splits) {
    using Cfg = PairCfg<MT>;
    auto kern = gemm4_pair_kernel<T, QT, MT, TRACE>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};  // the opt-in is per device
    static int pairs_per_wave[64] = {};
    if (dev < 0 || dev >= 64) return false;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess) {
            set_last_error("gemm4_pair smem attr", cudaGetLastError());
            return false;
        }
        attr_set[dev] = true;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    if (pairs_per_wave[dev] == 0) {
        cudaLaunchConfig_t qc{};
        qc.gridDim = dim3(2 * 128, 1, 1);
        qc.blockDim = dim3(kThreads);
        qc.dynamicSmemBytes = Cfg::kSmemBytes;
        qc.attrs = attr;
        qc.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &qc) == cudaSuccess && nclusters > 0) {
            pairs_per_wave[dev] = nclusters;
        } else {
            (void)cudaGetLastError();
            pairs_per_wave[dev] = device_sm_count() / 2;
        }
    }
    const int P = pairs_per_wave[dev];

    CUtensorMap tmap_x, tmap_w;
    if (!cached_tmap(&tmap_x, A, 2, 128, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, (uint32_t)Cfg::kBoxRows, 64u))
        return false;
    if (!cached_tmap(&tmap_w, p.B, 1, 64, (uint64_t)p.N, (uint64_t)p.K / 2, (uint64_t)p.K / 2, (uint32_t)kTileN, 64u))
        return false;

    // output tiles leave through TMA stores when the destination rows are 16-byte aligned
    OutMaps omaps{};
    p.tma_out = 0;
    {
        bool ok = (p.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
        for (int r = 0; r < p.n_peers; ++r) ok = ok && (reinterpret_cast<uintptr_t>(p.peer_out[r]) & 15) == 0;
        if (ok) {
            for (int d = 0; d <= p.n_peers && ok; ++d)
                ok = encode_tmap_2d(&omaps.m[d], d == 0 ? p.out : p.peer_out[d - 1], 2, 0, (uint64_t)p.M, (uint64_t)p.N,
                                    (uint64_t)p.ldc * 2, (uint32_t)(MT / 4), (uint32_t)kTileN);
            p.tma_out = ok ? 1 : 0;
        }
    }

    p.n_pairs = (p.N + 2 * kTileN - 1) / (2 * kTileN);
    const int m_tiles = (p.M + MT - 1) / MT;
    const int tiles = p.n_pairs * m_tiles;
    p.tiles_total = tiles;

    // K split of the partial last wave (or of everything when the grid is less than half a wave)
    int splits = 1, tiles_main = tiles;
    const int rem = tiles % P;
    // production: >= 8 a-stages (512 k) per split; forced (tests): >= 2
    const int min_stages = force_splits > 0 ? 2 : 8;
    const int max_by_k = p.ka_total / min_stages > 0 ? p.ka_total / min_stages : 1;
    auto clamp = [&](int v) {
        if (v > 2) v = 2;  // two-way only: the last arriver adds ONE partial to its own accumulators
        if (v > max_by_k) v = max_by_k;
        if (v < 1) v = 1;
        // no empty split: per = ceil(ka/v) rounded up to even
        while (v > 1) {
            int per = (p.ka_total + v - 1) / v;
            per += per & 1;
            if (per * (v - 1) < p.ka_total) break;
            --v;
        }
        return v;
    };
    if (force_splits > 0) {
        splits = clamp(force_splits);
        tiles_main = splits > 1 ? 0 : tiles;
        if (force_splits >= 100) {  // 100 + s: split only the partial last wave, when the production rule would
            splits = (rem > 0 && rem * 2 <= P) ? clamp(force_splits - 100) : 1;
            tiles_main = splits > 1 ? tiles - rem : tiles;
        }
    } else if (rem > 0 && rem * 2 <= P) {
        splits = clamp(P / rem);
        tiles_main = splits > 1 ? tiles - rem : tiles;
    }
    p.tiles_main = tiles_main;
    p.splits_tail = splits;
    p.ws_partial = nullptr;
    p.ws_counter = nullptr;
    const int split_tiles = tiles - tiles_main;
    if (split_tiles > 0) {
        Gemm4Workspace ws{};
        if (!gemm4_get_workspace(stream, size_t(split_tiles) * 2 * splits * kTileN * MT * sizeof(float),
                                 size_t(split_tiles) * 2, &ws))
            return false;  // more split tiles than the fixed workspace holds (forced splits in tests): not served
        p.ws_partial = ws.partial;
        p.ws_counter = ws.counters;
    }

    cudaLaunchConfig_t cfg{};
    p.n_items = tiles_main + split_tiles * splits;
    cfg.gridDim = dim3(2 * (p.n_items < P ? p.n_items : P), 1, 1);  // persistent: one cluster per SM pair
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap_x, tmap_w, omaps, p);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_last_error("gemm4_pair launch", e);
        return false;
    }
    return true;
}

} // namespace

// Large-M route of the 4-bit GEMM.  Returns false when the shape is not served here (the caller falls back
// to the one-CTA kernel of gemm4_tc.cu).  `mt_override` (0 = automatic) and `force_splits` are for the
// probes / tests; `trace` (device buffer of 2 * kTraceRoles * kTraceStages + 4 * 1024 int64) selects the traced build.
template <typename T>
bool launch_gemm4_pair(const T* A, const uint8_t* B, const float* absmax, const uint8_t* absmax_8bit,
                       const float* absmax_code, const float* absmax_offset, T* out, const T* bias, int M, int N, int K,
                       int ldc, int blocksize, int quant_type, cudaStream_t stream, void* const* peers, int n_peers,
                       int mt_override, int force_splits, long long* trace) {
    if (K < 128 || (K % 64) != 0) return false;
    if (blocksize < 32 || (blocksize & (blocksize - 1)) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(B) & 15) != 0) return false;
    if (quant_type != kNF4 && quant_type != kFP4) return false;
    if (n_peers < 0 || n_peers > 7) return false;

    int MT = mt_override;
    if (MT == 0) {
        static const int env_mt = [] {
            const char* e = getenv("BNB_B200_PAIR_MT");  // developer override (profiling): 128 | 256 | 384
            return e ? atoi(e) : 0;
        }();
        if (env_mt == 128 || env_mt == 256 || env_mt == 384) MT = env_mt;
    }
    if (MT == 0) {
        if (M < 512) return false;  // the one-CTA kernel (with its split-K) serves the small token counts
        // Pick the token tile by the modelled time of the busiest cluster (cycles, measured on B200 in round 2,
        // profiles/r02_pair_trace.md): MT = 384 runs an a-stage in ~806 cycles (the two N = 192 MMAs per k-step
        // take 768), MT = 256 in ~656 (decode bound); an item costs ~9 k / ~7 k more for its drain, epilogue and
        // restart, a K half ~4 k on top for the partial exchange.
        const int n_pairs = (N + 255) / 256;
        const int P = device_sm_count() / 2;
        auto cost = [&](int mt, double stage_cycles, double item_cycles) {
            const int tiles = n_pairs * ((M + mt - 1) / mt);
            const int rem = tiles % P;
            const double item = (K / 64) * stage_cycles + item_cycles;
            double c = (tiles / P) * item;
            if (rem > 0) c += (rem * 2 <= P && K >= 1024) ? 0.5 * (K / 64) * stage_cycles + item_cycles + 4000.0 : item;
            return c;
        };
        MT = cost(384, 806.0, 9000.0) <= cost(256, 656.0, 7000.0) ? 384 : 256;
    }
    if (MT != 128 && MT != 256 && MT != 384) return false;

    PairParams p{};
    p.B = B;
    p.absmax = absmax;
    p.absmax_8bit = absmax_8bit;
    p.absmax_code = absmax_code;
    p.absmax_offset = absmax_offset;
    p.bias = bias;
    p.out = out;
    p.n_peers = n_peers;
    for (int r = 0; r < n_peers; ++r) p.peer_out[r] = peers[r];
    p.trace = trace;
    {
        const char* tl = getenv("BNB_B200_TRACE_LITE");
        p.trace_lite = (tl != nullptr && tl[0] == '1') ? 1 : 0;
    }
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = ldc;
    p.log2_bs = ilog2_pow2(blocksize);
    p.ka_total = K / kAK;

#define BNB200_PAIR_MT(QT, TR)                                                                                         \
    switch (MT) {                                                                                                      \
    case 128: return launch_pair_mt<T, QT, 128, TR>(A, p, stream, force_splits);                                       \
    case 256: return launch_pair_mt<T, QT, 256, TR>(A, p, stream, force_splits);                                       \
    default: return launch_pair_mt<T, QT, 384, TR>(A, p, stream, force_splits);                                        \
    }
    if (trace != nullptr) {
        if (quant_type == kNF4) {
            BNB200_PAIR_MT(kNF4, true)
        } else {
            BNB200_PAIR_MT(kFP4, true)
        }
    }
    if (quant_type == kNF4) {
        BNB200_PAIR_MT(kNF4, false)
    } else {
        BNB200_PAIR_MT(kFP4, false)
    }
#undef BNB200_PAIR_MT
}

template bool launch_gemm4_pair<__nv_bfloat16>(const __nv_bfloat16*, const uint8_t*, const float*, const uint8_t*,
                                               const float*, const float*, __nv_bfloat16*, const __nv_bfloat16*, int,
                                               int, int, int, int, int, cudaStream_t, void* const*, int, int, int,
                                               long long*);
template bool launch_gemm4_pair<__half>(const __half*, const uint8_t*, const float*, const uint8_t*, const float*,
                                        const float*, __half*, const __half*, int, int, int, int, int, int,
                                        cudaStream_t, void* const*, int, int, int, long long*);

} // namespace bnb200
