// tcgen05 / TMA / TMEM contraction kernel for sm_100a (MODE_BF16 and MODE_BF16X3).
//
// Persistent, warp-specialised: one CTA per SM loops over 128 x BN output tiles.
//   warp 0      TMA producer   cp.async.bulk.tensor (128B-swizzled K-major boxes) into a multi-stage smem ring
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16),
//                              fp32 accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//                              overlaps the main loop of tile i+1; tcgen05.commit releases smem slots / publishes TMEM
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue (gemm.cuh) -> HBM
//                              (two warps per TMEM lane quarter, each draining half of the tile's columns)
// A k=5 dilated conv is five K-segments whose A boxes are the same activation matrix shifted by (j-2)*d rows
// (TMA zero-fills rows outside [0,R); utterances are separated by >= 16 zero rows, so no tap crosses a sequence).
// MODE_BF16X3 loads hi and lo planes of both operands and issues hi*hi + lo*hi + hi*lo into the same accumulator.
// The weight-grad GEMM contracts over rows: dY [R][M] and X [R][N] are read as MN-major UMMA operands straight from the
// row matrices (64-channel x 64-row boxes, 128B swizzle), so the tap shift is an outer (row) TMA coordinate and no
// transposed copies exist; split-K partial tiles are reduced with fp32 atomics.
// CTA-pair variant (CL = 2, tcgen05.mma.cta_group::2, UMMA 256 x BN x 16): used whenever the row-tile count is even.
// Roofline: tensor pipe (dense bf16, MEASURED_PEAKS.json); BF16X3 has one third of it.
#include <cuda.h>
#include <stdlib.h>
#include "gemm.cuh"

namespace radmmm {

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kEpiWarps = 8;                    // two epilogue warps per scheduler: each TMEM lane quarter is split in
                                                // two column halves
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kMaxMaps = 4;   // unique A maps and unique B maps per launch

struct TcSeg {
    int a_map, b_map;      // indices into the map tables
    int a_row_shift;       // row GEMM: shift on the A outer coordinate; weight-grad: unused
    int b_row_off;         // row GEMM: first weight row of this segment inside the B map; weight-grad: K shift of X
    int k_blocks;          // K / 64
};

struct TcParams {
    CUtensorMap a_hi[kMaxMaps], a_lo[kMaxMaps], b_hi[kMaxMaps], b_lo[kMaxMaps];
    TcSeg seg[kMaxSeg];
    int n_seg, n_a_maps, n_b_maps;
    int m_tiles, n_tiles, taps, split_k, k_blocks_total;   // weight-grad: k_blocks_total = R/64
    int acc_segs;                                          // weight-grad: all segments accumulate into one output
    int balanced;                                          // weight-grad: K blocks of all tiles cut into equal runs per CTA (pair)
    CUtensorMap out_map[2];                                // TMA store of epi.out0 (hi / lo plane): 32 x 32 bf16 boxes, 64-byte swizzle
    int out_tma;                                           // 1: the epilogue writes epi.out0 with cp.async.bulk.tensor stores
    int debug;                                             // probe bits (tools/gemm_probe.py): 1 no epilogue, 2 no MMA, 4 no TMA
    unsigned long long* trace;                             // tools/gemm_probe.py: [CTA][kTraceSlots] SM-clock / globaltimer stamps
    int trace_ctas;
    EpiParams epi;
};

// In-kernel timeline (diagnostic, tools/gemm_probe.py).  Slots per CTA: 0 entry, 1 set-up done, 2 first TMA issued,
// 3 first operand stage landed, 4 last MMA committed, 5 accumulator visible to the epilogue, 6 epilogue done (warp 4),
// 7 after the final CTA / cluster barrier (all clock64 of this SM); 8 / 9 globaltimer at entry / exit (ns, device-wide).
constexpr int kTraceSlots = 10;
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RADMMM_TRACE(slot) do { if (P.trace != nullptr && (int)blockIdx.x < P.trace_ctas) \
        P.trace[(size_t)blockIdx.x * kTraceSlots + (slot)] = (unsigned long long)clock64(); } while (0)

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1ll << 22)) {     // a pipeline bug must trap, never hang the device
            printf("radmmm gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// cta_group::2 form: executed by both CTAs of a pair; the transaction bytes update the LEADER's barrier (peer bit of the
// shared::cluster barrier address cleared, as CUTLASS' SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
template <int CL>
__device__ __forceinline__ void tma_load(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    if constexpr (CL == 2) tma_load_2d_2sm(smem_dst, map, bar, c0, c1);
    else tma_load_2d(smem_dst, map, bar, c0, c1);
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int CL>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    if constexpr (CL == 2) {     // arrive on the barrier at this offset in BOTH CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}
template <int CL>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CL == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MN-major, 128B-swizzled operand tile (weight-grad): atoms of 64 channels (128 B) x 8 K-rows; atoms along K are
// 1024 B apart (SBO), 64-channel atom columns are BK*128 B = 8192 B apart (LBO); one UMMA_K step = 16 K-rows = 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((BK * 128) >> 4) << 16;            // leading byte offset: next 64-channel atom column
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, 16-byte units
    d |= (uint64_t)0 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn, bool mn_major, int cl) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((BM * cl) >> 4) << 24);      // M = 128 per CTA, 256 for a CTA pair
}

template <int MODE, int BN, int CL>
struct Cfg {
    static constexpr int planes = (MODE == MODE_BF16X3) ? 2 : 1;
    static constexpr int a_bytes = BM * BK * 2;                 // this CTA's 128 rows (row GEMM) / 128 channels (weight-grad)
    static constexpr int b_rows = BN / CL;                      // cta_group::2: each CTA of the pair holds half of B's N extent
    static constexpr int b_bytes = b_rows * BK * 2;
    static constexpr int stage_bytes = planes * (a_bytes + b_bytes);
#ifndef RADMMM_TC_SMEM_KB
#define RADMMM_TC_SMEM_KB 200
#endif
    static constexpr int stages = (RADMMM_TC_SMEM_KB * 1024) / stage_bytes > 8 ? 8 : (RADMMM_TC_SMEM_KB * 1024) / stage_bytes;
    static constexpr int tmem_cols = 2 * BN;      // 256 or 512: powers of two
    static constexpr int stage_tile_bytes = 32 * kStageLd * 2;     // epilogue staging tile, per epilogue warp
    static constexpr int smem_bytes = stages * stage_bytes + 1024 /*align slack*/ + 512 /*barriers; keeps the staging tiles 512-byte aligned*/ + kEpiWarps * stage_tile_bytes;
};

// CL == 1: one CTA per 128 x BN tile, tcgen05.mma.cta_group::1.
// CL == 2: a CTA pair (cluster of 2) per 256 x BN tile, tcgen05.mma.cta_group::2 issued by the leader (rank 0): each
//          CTA stages its own 128 rows of A and HALF of B, the pair's tensor cores read both halves, each CTA keeps its
//          128 accumulator rows in its own TMEM.  Per-CTA shared-memory traffic per MMA drops from 48 KB to 32 KB
//          (BN = 256), which buys 6 pipeline stages instead of 4 -- the 1-CTA kernel was TMA-latency bound with the
//          tensor pipe 57 % active (profiles/).
#ifdef RADMMM_TC_MAXNREG
#define RADMMM_TC_BOUNDS __maxnreg__(RADMMM_TC_MAXNREG)
#else
#define RADMMM_TC_BOUNDS __launch_bounds__(kThreads, 1)
#endif
template <int MODE, int KIND, int BN, bool WGRAD, int CL>
__global__ void RADMMM_TC_BOUNDS gemm_tc_kernel(const __grid_constant__ TcParams P) {
    using C = Cfg<MODE, BN, CL>;
    static_assert(CL == 1 || CL == 2, "cluster size");
    const uint32_t crank = (CL == 2) ? cluster_rank() : 0u;
    const bool leader = crank == 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::stages * C::stage_bytes);
    uint64_t* full = bars;                       // [stages]  (CL == 2: only the leader's are waited on)
    uint64_t* empty = bars + C::stages;          // [stages]
    uint64_t* tfull = bars + 2 * C::stages;      // [2]
    uint64_t* tempty = bars + 2 * C::stages + 2; // [2]       (CL == 2: only the leader's are waited on)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::stages + 4);
    __nv_bfloat16* stage_tiles = reinterpret_cast<__nv_bfloat16*>(smem + C::stages * C::stage_bytes + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        RADMMM_TRACE(0);
        if (P.trace != nullptr && (int)blockIdx.x < P.trace_ctas) P.trace[(size_t)blockIdx.x * kTraceSlots + 8] = gtimer();
    }

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_a_maps; ++i) { prefetch_tmap(&P.a_hi[i]); if (C::planes == 2) prefetch_tmap(&P.a_lo[i]); }
        for (int i = 0; i < P.n_b_maps; ++i) { prefetch_tmap(&P.b_hi[i]); if (C::planes == 2) prefetch_tmap(&P.b_lo[i]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::stages; ++i) { mbar_init(&full[i], CL); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CL * 32 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CL == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // the peer's barriers exist before any remote arrive / 2-CTA TMA / MMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) RADMMM_TRACE(1);
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) touches no memory
    // the previous kernel of the stream produces, so with the launch attribute set it overlaps that kernel's tail; from here on
    // this grid reads its operands.  Without the attribute both instructions are no-ops.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Tile walk.  A "walk" position is one 128 x BN tile (CL == 1) or one 256 x BN pair tile (CL == 2, CTA `crank` takes
    // row tile 2*pm + crank).  Weight-grad: the (tap, split-K range) pair is the slowest index.
    const int tiles_mn = P.m_tiles * P.n_tiles;
    const int walk_mn = tiles_mn / CL;
    const int walk_end = WGRAD ? walk_mn * P.taps * P.split_k : walk_mn;
    const int walk_begin = (int)blockIdx.x / CL, walk_step = (int)gridDim.x / CL;
    auto decode = [&](int walk, int& m_blk, int& n_blk, int& tap, int& split) {
        const int mn = walk % walk_mn, ts = walk / walk_mn;
        tap = WGRAD ? ts / P.split_k : 0;
        split = WGRAD ? ts % P.split_k : 0;
        m_blk = (mn / P.n_tiles) * CL + (int)crank;
        n_blk = mn % P.n_tiles;
    };
    // weight-grad split-K range (in 64-row K blocks)
    auto k_range = [&](int split, int& kb0, int& kb1) {
        const int per = (P.k_blocks_total + P.split_k - 1) / P.split_k;
        kb0 = min(P.k_blocks_total, split * per);
        kb1 = min(P.k_blocks_total, kb0 + per);
    };
    const bool all_segs = !WGRAD || P.acc_segs;     // segments accumulate into one tile (row GEMM / accumulating weight-grad)
    // Balanced weight-grad walk (P.balanced): the job is walk_mn * taps tiles of k_blocks_total K blocks each; the K blocks of
    // all tiles are laid end to end and cut into gridDim.x / CL equal contiguous runs, one per CTA (pair).  A run crosses
    // tile boundaries, so a CTA works on up to ceil(run / k_blocks_total) + 1 partial tiles and every tile is reduced
    // with red.add into a zero-filled output -- no worker idles for the quantisation tail a whole-tile walk leaves
    // (5 taps x 32 pair tiles on 74 pairs: 3 rounds where 2.16 would do).
    const bool balanced = WGRAD && P.balanced != 0;
    int u_begin = 0, u_end = 0;
    if (balanced) {
        const long long units = (long long)walk_mn * P.taps * P.k_blocks_total;
        u_begin = (int)(units * walk_begin / walk_step);
        u_end = (int)(units * (walk_begin + 1) / walk_step);
    }
    // One "item" = one accumulator tile's worth of work for this CTA: (tile, tap, K-block range).  All three roles walk the
    // same item sequence.
    auto item_first = [&]() -> int { return balanced ? u_begin : walk_begin; };
    auto item_valid = [&](int pos) -> bool { return balanced ? pos < u_end : pos < walk_end; };
    auto item_decode = [&](int pos, int& m_blk, int& n_blk, int& tap, int& kb0, int& kb1) {
        if (balanced) {
            const int tile = pos / P.k_blocks_total;
            kb0 = pos - tile * P.k_blocks_total;
            kb1 = min(P.k_blocks_total, kb0 + (u_end - pos));
            const int mn = tile % walk_mn;
            tap = tile / walk_mn;
            m_blk = (mn / P.n_tiles) * CL + (int)crank;
            n_blk = mn % P.n_tiles;
        } else {
            int split;
            decode(pos, m_blk, n_blk, tap, split);
            kb0 = 0;
            kb1 = 0;
            if (WGRAD) k_range(split, kb0, kb1);
        }
    };
    auto item_next = [&](int pos, int kb0, int kb1) -> int { return balanced ? pos + (kb1 - kb0) : pos + walk_step; };

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            bool first_load = true;
            int m_blk, n_blk, tap, ikb0 = 0, ikb1 = 0;
            for (int pos = item_first(); item_valid(pos); pos = item_next(pos, ikb0, ikb1)) {
                item_decode(pos, m_blk, n_blk, tap, ikb0, ikb1);
                const int seg_begin = all_segs ? 0 : tap, seg_end = all_segs ? P.n_seg : tap + 1;
                for (int s = seg_begin; s < seg_end; ++s) {
                    const TcSeg sg = P.seg[s];
                    const int kb0 = WGRAD ? ikb0 : 0, kb1 = WGRAD ? ikb1 : sg.k_blocks;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* st = smem + stage * C::stage_bytes;
                        uint8_t* sa_hi = st;
                        uint8_t* sa_lo = st + C::a_bytes;
                        uint8_t* sb_hi = st + C::planes * C::a_bytes;
                        uint8_t* sb_lo = sb_hi + C::b_bytes;
                        uint64_t* fb = &full[stage];
                        if (P.debug & 4) {                        // probe: no loads, just hand the slot over
                            if (CL == 2 && !leader) mbar_arrive_remote(fb, 0);
                            else mbar_arrive(fb);
                            if (++stage == C::stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        if (CL == 1) mbar_expect_tx(fb, C::stage_bytes);
                        if (!WGRAD) {
                            const int a_c0 = kb * BK, a_c1 = m_blk * BM + sg.a_row_shift;
                            const int b_c0 = kb * BK, b_c1 = n_blk * BN + (int)crank * C::b_rows + sg.b_row_off;
                            tma_load<CL>(sa_hi, &P.a_hi[sg.a_map], fb, a_c0, a_c1);
                            tma_load<CL>(sb_hi, &P.b_hi[sg.b_map], fb, b_c0, b_c1);
                            if (C::planes == 2) {
                                tma_load<CL>(sa_lo, &P.a_lo[sg.a_map], fb, a_c0, a_c1);
                                tma_load<CL>(sb_lo, &P.b_lo[sg.b_map], fb, b_c0, b_c1);
                            }
                        } else {
                            // 64-channel x 64-row boxes; inner coordinate = channel, outer = row (tap shift on X)
                            const int ra = kb * BK, rb = kb * BK + sg.b_row_off;
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i) {
                                tma_load<CL>(sa_hi + i * (BK * 128), &P.a_hi[sg.a_map], fb, m_blk * BM + i * 64, ra);
                                if (C::planes == 2) tma_load<CL>(sa_lo + i * (BK * 128), &P.a_lo[sg.a_map], fb, m_blk * BM + i * 64, ra);
                            }
#pragma unroll
                            for (int i = 0; i < C::b_rows / 64; ++i) {
                                const int col = n_blk * BN + (int)crank * C::b_rows + i * 64;
                                tma_load<CL>(sb_hi + i * (BK * 128), &P.b_hi[sg.b_map], fb, col, rb);
                                if (C::planes == 2) tma_load<CL>(sb_lo + i * (BK * 128), &P.b_lo[sg.b_map], fb, col, rb);
                            }
                        }
                        if (CL == 2) {
                            // both CTAs' bytes land on the LEADER's barrier: the leader posts the expected total, the
                            // peer contributes its arrival remotely
                            if (leader) mbar_expect_tx(fb, 2 * C::stage_bytes);
                            else mbar_arrive_remote(fb, 0);
                        }
                        if (first_load) { RADMMM_TRACE(2); first_load = false; }
                        if (++stage == C::stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer (leader only)
        if (leader) {
            constexpr uint32_t idesc = make_idesc(BN, WGRAD, CL);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            int m_blk, n_blk, tap, ikb0 = 0, ikb1 = 0;
            for (int pos = item_first(); item_valid(pos); pos = item_next(pos, ikb0, ikb1), ++it) {
                item_decode(pos, m_blk, n_blk, tap, ikb0, ikb1);
                int total_kb = 0;
                if (!WGRAD) {
                    for (int s = 0; s < P.n_seg; ++s) total_kb += P.seg[s].k_blocks;
                } else {
                    total_kb = (ikb1 - ikb0) * (P.acc_segs ? P.n_seg : 1);
                }
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (it == 0 && kb == 0 && lane == 0) RADMMM_TRACE(3);
                    if (P.debug & 2) {                            // probe: no MMAs, release the slot at once
                        if (elect_one()) {
                            mbar_arrive(&empty[stage]);
                            if (CL == 2) mbar_arrive_remote(&empty[stage], 1);
                            if (kb == total_kb - 1) tc_commit<CL>(&tfull[as]);
                        }
                    } else if (elect_one()) {
                        const uint32_t st = smem_u32(smem + stage * C::stage_bytes);
                        const uint32_t a_hi = st, a_lo = st + C::a_bytes;
                        const uint32_t b_hi = st + C::planes * C::a_bytes, b_lo = b_hi + C::b_bytes;
                        // K-major: one UMMA_K step = 32 B inside the 128 B swizzle row; MN-major: 16 K-rows = 2048 B
                        constexpr uint32_t kstep = WGRAD ? (UMMA_K * 128) : (UMMA_K * 2);
                        const uint64_t da_hi = WGRAD ? make_smem_desc_mn(a_hi) : make_smem_desc(a_hi);
                        const uint64_t db_hi = WGRAD ? make_smem_desc_mn(b_hi) : make_smem_desc(b_hi);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t koff = (uint64_t)((k * kstep) >> 4);
                            umma_bf16<CL>(tmem_d, da_hi + koff, db_hi + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        if (C::planes == 2) {
                            const uint64_t da_lo = WGRAD ? make_smem_desc_mn(a_lo) : make_smem_desc(a_lo);
                            const uint64_t db_lo = WGRAD ? make_smem_desc_mn(b_lo) : make_smem_desc(b_lo);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                const uint64_t koff = (uint64_t)((k * kstep) >> 4);
                                umma_bf16<CL>(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
                                umma_bf16<CL>(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                            }
                        }
                        tc_commit<CL>(&empty[stage]);                       // smem slot(s) reusable once these MMAs retire
                        if (kb == total_kb - 1) { tc_commit<CL>(&tfull[as]); RADMMM_TRACE(4); }  // accumulator complete (both CTAs' epilogues)
                    }
                    __syncwarp();
                    if (++stage == C::stages) { stage = 0; phase ^= 1; }
                }
                if (total_kb == 0 && elect_one()) tc_commit<CL>(&tfull[as]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------------------------------ epilogue
        const int q = (warp - 4) & 3;                // TMEM lane quarter of this warp (== warp % 4)
        const int half = (warp - 4) >> 2;            // which column half of the tile this warp drains
        constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
        const Stager stager{stage_tiles + (warp - 4) * (32 * kStageLd), lane, P.out_tma ? &P.out_map[0] : nullptr,
                            P.out_tma ? &P.out_map[1] : nullptr, P.out_tma ? P.epi.out0.ptr : nullptr};
        int it = 0;
        int m_blk, n_blk, tap, ikb0 = 0, ikb1 = 0;
        for (int pos = item_first(); item_valid(pos); pos = item_next(pos, ikb0, ikb1), ++it) {
            item_decode(pos, m_blk, n_blk, tap, ikb0, ikb1);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int row = m_blk * BM + q * 32 + lane;
            if constexpr (!WGRAD && KIND == EPI_DH) {
                // the saved activations this warp needs are fetched BEFORE waiting for the accumulator, two 32-column
                // chunks ahead, so their latency hides behind the main loop (one tile per CTA: nothing else overlaps it)
                constexpr int NCH = kColsPerWarp / 32;
                RawTile<MODE> raw[2];
                const int nb = n_blk * BN + half * kColsPerWarp;
                staged_fetch32<MODE>(stager, P.epi.h, row, nb, raw[0]);
                if (NCH > 1) staged_fetch32<MODE>(stager, P.epi.h, row, nb + 32, raw[1]);
                mbar_wait(&tfull[as], aphase);
                tc_fence_after();
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    if (P.debug & 1) break;
                    float v[32], hv[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + half * kColsPerWarp + ci * 32), v);
                    staged_unpack32<MODE>(stager, raw[ci & 1], hv);
                    if (ci + 2 < NCH) staged_fetch32<MODE>(stager, P.epi.h, row, nb + (ci + 2) * 32, raw[ci & 1]);
                    epi_dh_with_h<MODE>(P.epi, stager, row, nb + ci * 32, v, hv);
                }
                tc_fence_before();
                if (CL == 2 && !leader) mbar_arrive_remote(&tempty[as], 0);
                else mbar_arrive(&tempty[as]);
                continue;
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            if (warp == 4 && lane == 0) RADMMM_TRACE(5);
            const bool has_acc = !WGRAD || ikb1 > ikb0;
#pragma unroll 1
            for (int c = half * kColsPerWarp; c < (half + 1) * kColsPerWarp; c += 32) {
                if (P.debug & 1) break;                           // probe: no epilogue
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c), v);
                const int n0 = n_blk * BN + c;
                if (WGRAD) {
                    if (has_acc) epi_wgrad<32>(P.epi, tap, row, n0, v);
                } else {
                    if (n0 < P.epi.N || KIND == EPI_RS || KIND == EPI_START || KIND == EPI_IN)
                        epi_apply<MODE, KIND, 32>(P.epi, stager, row, n0, v);
                }
            }
            tc_fence_before();
            if (warp == 4 && lane == 0) RADMMM_TRACE(6);
            if (CL == 2 && !leader) mbar_arrive_remote(&tempty[as], 0);     // the leader's MMA warp owns the wait
            else mbar_arrive(&tempty[as]);
        }
    }

    if (warp >= 4 && lane == 0 && P.out_tma) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's TMA stores are done
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // no CTA may exit while its peer can still arrive on its barriers / read its smem
    if (threadIdx.x == 0) {
        RADMMM_TRACE(7);
        if (P.trace != nullptr && (int)blockIdx.x < P.trace_ctas) P.trace[(size_t)blockIdx.x * kTraceSlots + 9] = gtimer();
    }
    if (warp == 2) {
        tc_fence_after();
        if (CL == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

// bf16 matrix [outer][inner] with row pitch ld elements; box = 64 x box_rows, 128B swizzle, zero fill out of bounds
// bf16 output matrix [outer][ld]: 32-column x 32-row boxes, 64-byte swizzle (the epilogue's TMA stores)
static int make_store_map(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld) {
    EncodeFn enc = get_encode();
    RADMMM_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RADMMM_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled (store map) failed with code %d", (int)r);
    return RADMMM_OK;
}

static int make_map(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
    EncodeFn enc = get_encode();
    RADMMM_REQUIRE(enc != nullptr, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    RADMMM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, "gemm_tc: operand not 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * 2)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RADMMM_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed with code %d (inner=%lld outer=%lld ld=%lld box=%d)",
                   (int)r, inner, outer, ld, box_rows);
    return RADMMM_OK;
}

constexpr int kMaxDevices = 64;      // function attributes and SM counts are per device
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
static int sm_count() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

template <int MODE, int KIND, int BN, bool WGRAD, int CL>
static int launch_inst(const TcParams& P, int n_tiles_total, cudaStream_t st) {
    using C = Cfg<MODE, BN, CL>;
    auto kern = gemm_tc_kernel<MODE, KIND, BN, WGRAD, CL>;
    static bool configured_dev[kMaxDevices] = {};
    static int max_clusters_dev[kMaxDevices] = {};
    const int dev = current_device();
    bool& configured = configured_dev[dev];
    int& max_clusters = max_clusters_dev[dev];
    if (!configured) {
        RADMMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
        if (CL > 1) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(CL * 32); q.blockDim = dim3(kThreads); q.dynamicSmemBytes = C::smem_bytes;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                max_clusters = sm_count() / CL / 2;
            }
        }
        configured = true;
    }
    // RADMMM_B200_PDL=1: programmatic stream serialization (see the kernel) for every contraction launch
    static const bool pdl = []() { const char* e = getenv("RADMMM_B200_PDL"); return e && e[0] == '1'; }();
    int grid;
    if (CL == 1) {
        grid = n_tiles_total < sm_count() ? n_tiles_total : sm_count();
        if (grid < 1) grid = 1;
    } else {
        int clusters = n_tiles_total / CL;
        if (clusters > max_clusters) clusters = max_clusters;
        if (clusters < 1) clusters = 1;
        grid = clusters * CL;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (CL > 1) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = CL; attr[n_attr].val.clusterDim.y = 1; attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr; cfg.numAttrs = n_attr;
    RADMMM_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    count_launch();
    return RADMMM_OK;
}

template <int MODE, int BN, int CL>
static int launch_kind(const TcParams& P, int n_tiles_total, bool wgrad, cudaStream_t st) {
    if (wgrad) return launch_inst<MODE, EPI_WGRAD, BN, true, CL>(P, n_tiles_total, st);
    switch (P.epi.kind) {
        case EPI_START: return launch_inst<MODE, EPI_START, BN, false, CL>(P, n_tiles_total, st);
        case EPI_IN: return launch_inst<MODE, EPI_IN, BN, false, CL>(P, n_tiles_total, st);
        case EPI_RS: return launch_inst<MODE, EPI_RS, BN, false, CL>(P, n_tiles_total, st);
        case EPI_END: return launch_inst<MODE, EPI_END, BN, false, CL>(P, n_tiles_total, st);
        case EPI_DOUT: return launch_inst<MODE, EPI_DOUT, BN, false, CL>(P, n_tiles_total, st);
        case EPI_DH: return launch_inst<MODE, EPI_DH, BN, false, CL>(P, n_tiles_total, st);
        case EPI_DH0: return launch_inst<MODE, EPI_DH0, BN, false, CL>(P, n_tiles_total, st);
        case EPI_DZ0: return launch_inst<MODE, EPI_DZ0, BN, false, CL>(P, n_tiles_total, st);
        case EPI_DCTX: return launch_inst<MODE, EPI_DCTX, BN, false, CL>(P, n_tiles_total, st);
        case EPI_F32: return launch_inst<MODE, EPI_F32, BN, false, CL>(P, n_tiles_total, st);
    }
    set_error("gemm_tc: unknown epilogue kind %d", P.epi.kind);
    return RADMMM_ERR_ARG;
}

struct MapKey { const void* ptr; long long ld, plane; };

// RADMMM_B200_TC_PAIR=0 forces the 1-CTA kernel (debugging / A-B measurements)
static bool pair_mode_enabled() {
    const char* e = getenv("RADMMM_B200_TC_PAIR");
    return !(e && e[0] == '0');
}
static int probe_bits() {
    const char* e = getenv("RADMMM_B200_TC_PROBE");
    return e ? atoi(e) : 0;
}
static unsigned long long* g_trace = nullptr;
static int g_trace_ctas = 0, g_trace_launches = 0, g_trace_next = 0;

}  // namespace

void gemm_tc_set_trace(void* buf, int max_ctas, int max_launches) {
    g_trace = reinterpret_cast<unsigned long long*>(buf);
    g_trace_ctas = max_ctas; g_trace_launches = max_launches; g_trace_next = 0;
}

int launch_gemm_tc(const GemmArgs& args, int mode, cudaStream_t stream) {
    RADMMM_REQUIRE(mode == MODE_BF16 || mode == MODE_BF16X3, "gemm_tc: bad mode %d", mode);
    RADMMM_REQUIRE(args.R % BM == 0, "gemm_tc: R=%d must be a multiple of %d", args.R, BM);
    RADMMM_REQUIRE(args.n_seg >= 1 && args.n_seg <= kMaxSeg, "gemm_tc: bad segment count %d", args.n_seg);
    const bool x3 = mode == MODE_BF16X3;
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.epi = args.epi;
    P.n_seg = args.n_seg;
    P.debug = probe_bits();
    if (g_trace != nullptr && g_trace_next < g_trace_launches) {       // launch i writes region i of the buffer
        P.trace = g_trace + (size_t)g_trace_next++ * g_trace_ctas * kTraceSlots;
        P.trace_ctas = g_trace_ctas;
    }

    // output tiling
    const int N = args.epi.N;
    int n_pad = (int)round_up(N, 128);
    // weight-grads whose width is an odd number of 128-column tiles (the context / LSTM input weights: N = 1088) take one more
    // all-zero half tile so that they run 256-wide: MN-major 128-wide tiles stream twice the operand bytes per tensor cycle
    // and run at less than half the rate (see the weight-grad tiling note below)
    if (args.wgrad && !x3 && N >= 512 && n_pad % 256 != 0) n_pad += 128;
    int BN = (!x3 && n_pad % 256 == 0) ? 256 : 128;
    if (args.wgrad && BN == 256 && args.split_k < 1) {
        // Weight-grad tiling.  128-wide tiles (160 pair tiles for the 5-tap dilated conv: plain stores, no split-K) looked
        // attractive for their tile count, but a 256 x 128 pair tile streams 24 KB of MN-major operands per 64-row K block
        // for 256 tensor cycles -- 96 B/clk per SM, ~14 KB/clk over the chip, more than twice what L2 delivers: measured
        // 0.46 us per K block instead of 0.13 (profiles/r2_gemm_timeline.md).  256-wide tiles stream 32 KB per 512 cycles,
        // the same 64 B/clk per SM as the forward convolutions, which run at the tensor peak; their 80 pair tiles run whole
        // (no split along K, plain stores -- see the reduction-strategy note further down).
        // RADMMM_B200_WGRAD_BN128=1 restores the old choice (A/B measurements).
        static const bool bn128 = []() { const char* e = getenv("RADMMM_B200_WGRAD_BN128"); return e && e[0] == '1'; }();
        const int taps = args.wgrad == 2 ? 1 : args.n_seg;
        const long long slots = sm_count();
        const long long t256 = (long long)cdiv(args.epi.M, BM) * (n_pad / 256) * taps;
        const long long t128 = (long long)cdiv(args.epi.M, BM) * (n_pad / 128) * taps;
        if (bn128 && 2 * t256 < 3 * slots && 2 * t128 >= 3 * slots) BN = 128;
    }
    // narrow outputs (the `end` conv: N = C <= 160 -> one 256-wide tile per 128 rows, 26 CTAs at B=8): 128-wide tiles double
    // the CTA count and halve each CTA's main loop
    if (!args.wgrad && BN == 256 && (long long)(args.R / BM) * (n_pad / 256) * 2 <= sm_count()) BN = 128;
    P.n_tiles = n_pad / BN;
    int n_tiles_total;

    MapKey a_keys[kMaxMaps], b_keys[kMaxMaps];
    int n_a = 0, n_b = 0;
    auto find_or_add = [&](MapKey* keys, int& n, const void* ptr, long long ld, long long plane) -> int {
        for (int i = 0; i < n; ++i)
            if (keys[i].ptr == ptr && keys[i].ld == ld && keys[i].plane == plane) return i;
        if (n == kMaxMaps) return -1;
        keys[n] = MapKey{ptr, ld, plane};
        return n++;
    };

    bool use_cl = false;
    if (!args.wgrad) {
        P.m_tiles = args.R / BM;
        n_tiles_total = P.m_tiles * P.n_tiles;
        use_cl = (P.m_tiles % 2 == 0) && pair_mode_enabled();
        const int a_box = BM, b_box = use_cl ? BN / 2 : BN;
        // B maps: segments whose weight matrices sit in one allocation (same ld / plane stride, row-aligned offsets)
        // share a map anchored at the lowest pointer; the segment carries its row offset.
        for (int s = 0; s < args.n_seg; ++s) {
            const GemmSeg& g = args.seg[s];
            RADMMM_REQUIRE(g.K % BK == 0 && g.K > 0, "gemm_tc: segment K=%d must be a positive multiple of %d", g.K, BK);
            RADMMM_REQUIRE(g.a.ptr && g.w.ptr, "gemm_tc: null operand");
            int ai = find_or_add(a_keys, n_a, g.a.ptr, g.a.ld, g.a.plane_stride);
            RADMMM_REQUIRE(ai >= 0, "gemm_tc: too many distinct A operands");
            int bi = -1, row_off = 0;
            for (int i = 0; i < n_b; ++i) {
                if (b_keys[i].ld != g.w.ld || b_keys[i].plane != g.w.plane_stride) continue;
                const long long diff = (const char*)g.w.ptr - (const char*)b_keys[i].ptr;
                const long long row_bytes = g.w.ld * 2;
                if (diff % row_bytes == 0 && diff >= 0 && diff / row_bytes < (1 << 24)) {
                    bi = i; row_off = (int)(diff / row_bytes);
                    break;
                }
            }
            if (bi < 0) {
                bi = find_or_add(b_keys, n_b, g.w.ptr, g.w.ld, g.w.plane_stride);
                RADMMM_REQUIRE(bi >= 0, "gemm_tc: too many distinct B operands");
            }
            P.seg[s] = TcSeg{ai, bi, g.shift, row_off, g.K / BK};
        }
        // NB the B-map sharing above requires the anchor to be the LOWEST address; callers list segments in
        // ascending weight order (the WN builders do).  Outer extents: A = R rows exactly (zero fill beyond), B large.
        for (int i = 0; i < n_a; ++i) {
            long long kmax = 0;
            for (int s = 0; s < args.n_seg; ++s) if (P.seg[s].a_map == i) kmax = kmax > args.seg[s].K ? kmax : args.seg[s].K;
            RADMMM_TRY(make_map(&P.a_hi[i], a_keys[i].ptr, kmax, args.R, a_keys[i].ld, a_box));
            if (x3) RADMMM_TRY(make_map(&P.a_lo[i], (const __nv_bfloat16*)a_keys[i].ptr + a_keys[i].plane, kmax, args.R, a_keys[i].ld, a_box));
        }
        for (int i = 0; i < n_b; ++i) {
            long long kmax = 0, rows = 0;
            for (int s = 0; s < args.n_seg; ++s)
                if (P.seg[s].b_map == i) {
                    kmax = kmax > args.seg[s].K ? kmax : args.seg[s].K;
                    const long long need = (long long)P.seg[s].b_row_off + n_pad;
                    rows = rows > need ? rows : need;
                }
            RADMMM_TRY(make_map(&P.b_hi[i], b_keys[i].ptr, kmax, rows, b_keys[i].ld, b_box));
            if (x3) RADMMM_TRY(make_map(&P.b_lo[i], (const __nv_bfloat16*)b_keys[i].ptr + b_keys[i].plane, kmax, rows, b_keys[i].ld, b_box));
        }
    } else {
        // weight-grad: A = dY [R][M], B = X [R][N] as MN-major operands; one output tile set per tap, K = rows split
        // into split_k ranges
        const GemmSeg& g0 = args.seg[0];
        RADMMM_REQUIRE(g0.a.ptr && g0.w.ptr, "gemm_tc: null weight-grad operand");
        RADMMM_REQUIRE(g0.a.ld % 64 == 0 && g0.w.ld % 64 == 0, "gemm_tc: weight-grad operands need ld %% 64 == 0");
        const int M = args.epi.M;
        const bool grouped = args.wgrad == 3;
        RADMMM_REQUIRE(!grouped || args.n_seg <= 4, "gemm_tc: at most 4 grouped weight-grad problems per launch");
        P.m_tiles = cdiv(M, BM);
        // an odd number of 128-row tiles (the LSTM's dW_ih: M = 8 x 528 = 33 tiles) would fall back to the 1-CTA kernel; one
        // more all-zero tile (TMA zero-fills beyond M, the epilogue skips rows >= M) keeps it on CTA pairs: 59 -> ~30 us
        if (P.m_tiles > 1 && (P.m_tiles & 1) && pair_mode_enabled()) P.m_tiles += 1;
        P.acc_segs = args.wgrad == 2;
        P.taps = P.acc_segs ? 1 : args.n_seg;
        P.k_blocks_total = args.R / BK;
        // split K (= rows) so that the persistent grid sees >= ~4 tiles per SM while every tile keeps >= 8 K blocks
        const int tiles0 = P.m_tiles * P.n_tiles * P.taps;
        int split = args.split_k;
        // Whole tiles, no split-K (plain stores, no zero-fill, no red.add) whenever there are at least as many tiles as CTAs.
        // Inside the train step several weight-grad launches and the dgrad chain share the SMs, so a launch's own
        // quantisation tail (160 tiles on 148 CTAs) is filled by its neighbours and what counts is the total work: measured
        // 8.57 ms per step against 8.84 (two-way split-K) and 8.80 (balanced K-block runs) -- profiles/r2_wgrad_reduction_ab.md.
        // RADMMM_B200_WGRAD_WHOLE=0 restores the split; RADMMM_B200_WGRAD_BALANCED=1 selects the balanced walk.
        static const int whole_mode = []() { const char* e = getenv("RADMMM_B200_WGRAD_WHOLE"); return e ? atoi(e) : 2; }();
        const bool whole = whole_mode != 0;
        if (split < 1 && whole && tiles0 >= sm_count()) split = 1;
        // Few tiles AND a short K (<= 128 blocks of 64 rows, i.e. the benchmark's B=8 x T=800): still no split.  The launch
        // fills only part of the machine, but it runs on a weight-grad lane next to three others and the dgrad chain, and the
        // memset + red.add a split costs is pure extra work: 8.50 vs 8.64 ms per step.  Long K keeps the split (a handful of
        // CTAs walking thousands of K blocks would become the step's tail).  RADMMM_B200_WGRAD_WHOLE=1: tiles >= SMs only.
        if (split < 1 && whole_mode >= 2 && P.k_blocks_total <= 128) split = 1;
        if (split < 1 && 2 * (long long)tiles0 >= 3 * sm_count()) split = 1;      // enough tiles already
        if (split < 1) {
            split = cdiv(2 * sm_count(), tiles0);
            const int max_split = P.k_blocks_total / 8 > 1 ? P.k_blocks_total / 8 : 1;
            if (split > max_split) split = max_split;
        }
        if (split > P.k_blocks_total) split = P.k_blocks_total;
        if (split < 1) split = 1;
        {   // no empty ranges
            const int per = cdiv(P.k_blocks_total, split);
            split = cdiv(P.k_blocks_total, per);
        }
        use_cl = (P.m_tiles % 2 == 0) && pair_mode_enabled();
        // Whole tiles or equal K-block runs?  Whole tiles cost `rounds * K blocks per item`; the balanced walk costs
        // `units / workers` plus roughly one more epilogue (counted as 4 K blocks).  Take the balanced walk when it wins by
        // more than 10 % (the 5-tap dilated conv: 80 pair tiles on 74 pairs -- 78 K-block times split in two, 56 balanced).
        // Off by default (see above); RADMMM_B200_WGRAD_BALANCED=1 enables it for A/B measurements.
        static const bool allow_balanced = []() { const char* e = getenv("RADMMM_B200_WGRAD_BALANCED"); return e && e[0] == '1'; }();
        if (allow_balanced && args.zero_output && !P.acc_segs && args.split_k < 1) {
            split = args.split_k;                      // undo the whole-tile choice above: compare against the split it replaces
            if (split < 1) {
                split = cdiv(2 * sm_count(), tiles0);
                const int max_split = P.k_blocks_total / 8 > 1 ? P.k_blocks_total / 8 : 1;
                if (split > max_split) split = max_split;
                const int per = cdiv(P.k_blocks_total, split);
                split = cdiv(P.k_blocks_total, per);
            }
            const int cl = use_cl ? 2 : 1;
            const long long workers = sm_count() / cl;
            const long long items = (long long)(tiles0 / cl) * split;
            const long long tile_cost = cdiv(items, workers) * cdiv(P.k_blocks_total, split);
            const long long even = cdiv((long long)(tiles0 / cl) * P.k_blocks_total, workers) + 4;
            if (items > workers / 2 && 10 * even < 9 * tile_cost) { P.balanced = 1; split = 1; }
        }
        P.split_k = split;
        n_tiles_total = P.m_tiles * P.n_tiles * P.taps * P.split_k;
        if (P.balanced && n_tiles_total < 2 * sm_count()) n_tiles_total = 2 * sm_count();     // grid = every CTA (pair) slot
        if (args.zero_output) {                 // the launcher owns the reduction strategy
            P.epi.atomic = split > 1 || P.balanced;
            if (P.epi.atomic) RADMMM_TRY(zero_wgrad_output(args, stream));
        }
        RADMMM_REQUIRE((split == 1 && !P.balanced) || P.epi.atomic, "gemm_tc: split-K weight-grad needs the atomic epilogue");
        // inner extents stop at the logical widths (rounded to the 64-channel box) so that operands which are column
        // slices of wider matrices never read past their rows
        const long long a_in = round_up(M, 64) < g0.a.ld ? round_up(M, 64) : g0.a.ld;
        const long long b_in = round_up(N, 64) < g0.w.ld ? round_up(N, 64) : g0.w.ld;
        if (!grouped) {
            RADMMM_TRY(make_map(&P.a_hi[0], g0.a.ptr, a_in, args.R, g0.a.ld, 64));
            if (x3) RADMMM_TRY(make_map(&P.a_lo[0], (const __nv_bfloat16*)g0.a.ptr + g0.a.plane_stride, a_in, args.R, g0.a.ld, 64));
            n_a = 1;
        }
        for (int s = 0; s < args.n_seg; ++s) {
            const GemmSeg& g = args.seg[s];
            int ai = 0;
            if (grouped) {                       // every group brings its own dY (and X)
                RADMMM_REQUIRE(g.a.ptr && g.w.ptr && g.a.ld == g0.a.ld && g.w.ld == g0.w.ld, "gemm_tc: grouped weight-grad operands must share their row pitch");
                ai = find_or_add(a_keys, n_a, g.a.ptr, g.a.ld, g.a.plane_stride);
                RADMMM_REQUIRE(ai >= 0, "gemm_tc: too many distinct dY operands");
            } else {
                RADMMM_REQUIRE(g.a.ptr == g0.a.ptr, "gemm_tc: weight-grad segments must share dY");
                RADMMM_REQUIRE(P.acc_segs || g.w.ptr == g0.w.ptr, "gemm_tc: weight-grad taps must share X");
            }
            int bi = find_or_add(b_keys, n_b, g.w.ptr, g.w.ld, g.w.plane_stride);
            RADMMM_REQUIRE(bi >= 0, "gemm_tc: too many distinct X operands");
            P.seg[s] = TcSeg{ai, bi, 0, g.shift, P.k_blocks_total};
        }
        if (grouped) {
            for (int i = 0; i < n_a; ++i) {
                RADMMM_TRY(make_map(&P.a_hi[i], a_keys[i].ptr, a_in, args.R, a_keys[i].ld, 64));
                if (x3) RADMMM_TRY(make_map(&P.a_lo[i], (const __nv_bfloat16*)a_keys[i].ptr + a_keys[i].plane, a_in, args.R, a_keys[i].ld, 64));
            }
        }
        for (int i = 0; i < n_b; ++i) {
            RADMMM_TRY(make_map(&P.b_hi[i], b_keys[i].ptr, b_in, args.R, b_keys[i].ld, 64));
            if (x3) RADMMM_TRY(make_map(&P.b_lo[i], (const __nv_bfloat16*)b_keys[i].ptr + b_keys[i].plane, b_in, args.R, b_keys[i].ld, 64));
        }
    }
    P.n_a_maps = n_a;
    P.n_b_maps = n_b;
    // TMA stores for the kinds whose only activation output is out0 (RADMMM_B200_TMA_STORE=0: per-lane 16-byte stores)
    static const bool tma_store = []() { const char* e = getenv("RADMMM_B200_TMA_STORE"); return !(e && e[0] == '0'); }();
    const int kind = args.epi.kind;
    if (tma_store && !args.wgrad && (kind == EPI_START || kind == EPI_IN || kind == EPI_RS || kind == EPI_DH0 || kind == EPI_DH) &&
        args.epi.out0.ptr != nullptr && (reinterpret_cast<uintptr_t>(args.epi.out0.ptr) & 15) == 0 && (args.epi.out0.ld * 2) % 16 == 0) {
        const ActMat& o = args.epi.out0;
        RADMMM_TRY(make_store_map(&P.out_map[0], o.ptr, o.ld, args.R, o.ld));
        if (x3) RADMMM_TRY(make_store_map(&P.out_map[1], (const __nv_bfloat16*)o.ptr + o.plane_stride, o.ld, args.R, o.ld));
        P.out_tma = 1;
    }

    const bool wg = args.wgrad != 0;
    if (use_cl) {
        if (x3) return launch_kind<MODE_BF16X3, 128, 2>(P, n_tiles_total, wg, stream);
        if (BN == 256) return launch_kind<MODE_BF16, 256, 2>(P, n_tiles_total, wg, stream);
        return launch_kind<MODE_BF16, 128, 2>(P, n_tiles_total, wg, stream);
    }
    if (x3) return launch_kind<MODE_BF16X3, 128, 1>(P, n_tiles_total, wg, stream);
    if (BN == 256) return launch_kind<MODE_BF16, 256, 1>(P, n_tiles_total, wg, stream);
    return launch_kind<MODE_BF16, 128, 1>(P, n_tiles_total, wg, stream);
}

}  // namespace radmmm
