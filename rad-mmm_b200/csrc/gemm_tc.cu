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
// Cluster variant (CL = 4, a 2x2 group of tiles): the two CTAs that share a row tile each load half of the A box and
// multicast it to both, the two CTAs that share a column tile do the same for B, so every operand byte crosses the
// L2 -> SM fabric once per CTA pair instead of once per CTA (the k=5 conv at B=8, T=800 moves 400 MB per launch without
// it and is L2-bandwidth bound at ~730 TFLOP/s).  Stage release is a multicast tcgen05.commit to the three CTAs that
// write into this CTA's stage.
// Roofline: tensor pipe (dense bf16, MEASURED_PEAKS.json); BF16X3 has one third of it.
#include <cuda.h>
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
    EpiParams epi;
};

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
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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
__host__ __device__ constexpr uint32_t make_idesc(int bn, bool mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int MODE, int BN>
struct Cfg {
    static constexpr int planes = (MODE == MODE_BF16X3) ? 2 : 1;
    static constexpr int a_bytes = BM * BK * 2;
    static constexpr int b_bytes = BN * BK * 2;
    static constexpr int stage_bytes = planes * (a_bytes + b_bytes);
    static constexpr int stages = (200 * 1024) / stage_bytes > 8 ? 8 : (200 * 1024) / stage_bytes;
    static constexpr int tmem_cols = 2 * BN;      // 256 or 512: powers of two
    static constexpr int stage_tile_bytes = 32 * kStageLd * 2;     // epilogue staging tile, per epilogue warp
    static constexpr int smem_bytes = stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiWarps * stage_tile_bytes;
};

template <int MODE, int KIND, int BN, bool WGRAD, int CL>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams P) {
    using C = Cfg<MODE, BN>;
    static_assert(CL == 1 || CL == 4, "cluster size");
    // 2x2 cluster geometry: rank = cm + 2*cn; cm selects the row tile of the pair, cn the column tile
    const uint32_t crank = (CL == 4) ? cluster_rank() : 0u;
    const int cm = crank & 1, cn = crank >> 1;
    const uint16_t mask_a = (uint16_t)((1u << crank) | (1u << (crank ^ 2)));     // CTAs sharing my row tile (same cm)
    const uint16_t mask_b = (uint16_t)((1u << crank) | (1u << (crank ^ 1)));     // CTAs sharing my column tile (same cn)
    const uint16_t mask_rel = (uint16_t)(mask_a | mask_b);                        // CTAs that write into my stages
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::stages * C::stage_bytes);
    uint64_t* full = bars;                       // [stages]
    uint64_t* empty = bars + C::stages;          // [stages]
    uint64_t* tfull = bars + 2 * C::stages;      // [2]
    uint64_t* tempty = bars + 2 * C::stages + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::stages + 4);
    __nv_bfloat16* stage_tiles = reinterpret_cast<__nv_bfloat16*>(smem + C::stages * C::stage_bytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_a_maps; ++i) { prefetch_tmap(&P.a_hi[i]); if (C::planes == 2) prefetch_tmap(&P.a_lo[i]); }
        for (int i = 0; i < P.n_b_maps; ++i) { prefetch_tmap(&P.b_hi[i]); if (C::planes == 2) prefetch_tmap(&P.b_lo[i]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL == 4 ? 3 : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 4) cluster_sync_all();          // peers' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Tile walk.  CL == 1: CTA b takes tiles b, b + grid, ...  CL == 4: cluster c takes 2x2 super-tiles c, c + #clusters, ...
    // and CTA (cm, cn) of the cluster works on tile (2*sm + cm, 2*sn + cn); `tile` below always is the CTA's own linear
    // tile index in the (m_tiles x n_tiles [x taps x split]) space, `walk` the scheduler position.
    const int tiles_mn = P.m_tiles * P.n_tiles;
    const int n_tiles_total = WGRAD ? tiles_mn * P.taps * P.split_k : tiles_mn;
    const int walk_begin = (CL == 4) ? (int)(blockIdx.x >> 2) : (int)blockIdx.x;
    const int walk_step = (CL == 4) ? (int)(gridDim.x >> 2) : (int)gridDim.x;
    const int walk_end = (CL == 4) ? n_tiles_total / 4 : n_tiles_total;
    auto tile_of = [&](int walk) -> int {
        if (CL == 1) return walk;
        const int sn_tiles = P.n_tiles >> 1, smn = (P.m_tiles >> 1) * sn_tiles;
        const int outer = walk / smn, inner = walk % smn;          // outer = (tap, split) for weight-grad
        const int sm = inner / sn_tiles, sn = inner % sn_tiles;
        return outer * tiles_mn + (2 * sm + cm) * P.n_tiles + (2 * sn + cn);
    };

    // weight-grad split-K range (in 64-row K blocks)
    auto k_range = [&](int split, int& kb0, int& kb1) {
        const int per = (P.k_blocks_total + P.split_k - 1) / P.split_k;
        kb0 = min(P.k_blocks_total, split * per);
        kb1 = min(P.k_blocks_total, kb0 + per);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int walk = walk_begin; walk < walk_end; walk += walk_step) {
                const int tile = tile_of(walk);
                int mn = tile, tap = 0, split = 0;
                if (WGRAD) { mn = tile % tiles_mn; const int ts = tile / tiles_mn; tap = ts / P.split_k; split = ts % P.split_k; }
                const int m_blk = mn / P.n_tiles, n_blk = mn % P.n_tiles;
                const int seg_begin = (WGRAD && !P.acc_segs) ? tap : 0, seg_end = (WGRAD && !P.acc_segs) ? tap + 1 : P.n_seg;
                for (int s = seg_begin; s < seg_end; ++s) {
                    const TcSeg sg = P.seg[s];
                    int kb0 = 0, kb1 = sg.k_blocks;
                    if (WGRAD) k_range(split, kb0, kb1);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* st = smem + stage * C::stage_bytes;
                        mbar_expect_tx(&full[stage], C::stage_bytes);
                        uint8_t* sa_hi = st;
                        uint8_t* sb_hi = st + C::planes * C::a_bytes;
                        uint8_t* sa_lo = st + C::a_bytes;
                        uint8_t* sb_lo = st + 2 * C::a_bytes + C::b_bytes;
                        if (!WGRAD && CL == 4) {
                            // my half of each box (A: 64 of the 128 rows, B: BN/2 rows), multicast to the sharing pair
                            const int a_c0 = kb * BK, a_c1 = m_blk * BM + cn * (BM / 2) + sg.a_row_shift;
                            const int b_c0 = kb * BK, b_c1 = n_blk * BN + cm * (BN / 2) + sg.b_row_off;
                            const int a_off = cn * (BM / 2) * 128, b_off = cm * (BN / 2) * 128;
                            tma_load_2d_mc(sa_hi + a_off, &P.a_hi[sg.a_map], &full[stage], a_c0, a_c1, mask_a);
                            tma_load_2d_mc(sb_hi + b_off, &P.b_hi[sg.b_map], &full[stage], b_c0, b_c1, mask_b);
                            if (C::planes == 2) {
                                tma_load_2d_mc(sa_lo + a_off, &P.a_lo[sg.a_map], &full[stage], a_c0, a_c1, mask_a);
                                tma_load_2d_mc(sb_lo + b_off, &P.b_lo[sg.b_map], &full[stage], b_c0, b_c1, mask_b);
                            }
                        } else if (!WGRAD) {
                            const int a_c0 = kb * BK, a_c1 = m_blk * BM + sg.a_row_shift;
                            const int b_c0 = kb * BK, b_c1 = n_blk * BN + sg.b_row_off;
                            tma_load_2d(sa_hi, &P.a_hi[sg.a_map], &full[stage], a_c0, a_c1);
                            tma_load_2d(sb_hi, &P.b_hi[sg.b_map], &full[stage], b_c0, b_c1);
                            if (C::planes == 2) {
                                tma_load_2d(sa_lo, &P.a_lo[sg.a_map], &full[stage], a_c0, a_c1);
                                tma_load_2d(sb_lo, &P.b_lo[sg.b_map], &full[stage], b_c0, b_c1);
                            }
                        } else if (CL == 4) {
                            // weight-grad: 64-channel x 64-row boxes; I load box cn of A and boxes [cm*BN/128, ...) of B
                            const int ra = kb * BK, rb = kb * BK + sg.b_row_off;
                            tma_load_2d_mc(sa_hi + cn * (BK * 128), &P.a_hi[0], &full[stage], m_blk * BM + cn * 64, ra, mask_a);
                            if (C::planes == 2) tma_load_2d_mc(sa_lo + cn * (BK * 128), &P.a_lo[0], &full[stage], m_blk * BM + cn * 64, ra, mask_a);
#pragma unroll
                            for (int j = 0; j < BN / 128; ++j) {
                                const int i = cm * (BN / 128) + j;
                                tma_load_2d_mc(sb_hi + i * (BK * 128), &P.b_hi[sg.b_map], &full[stage], n_blk * BN + i * 64, rb, mask_b);
                                if (C::planes == 2) tma_load_2d_mc(sb_lo + i * (BK * 128), &P.b_lo[sg.b_map], &full[stage], n_blk * BN + i * 64, rb, mask_b);
                            }
                        } else {
                            // 64-channel x 64-row boxes; inner coordinate = channel, outer = row (tap shift on X)
                            const int ra = kb * BK, rb = kb * BK + sg.b_row_off;
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i) {
                                tma_load_2d(sa_hi + i * (BK * 128), &P.a_hi[0], &full[stage], m_blk * BM + i * 64, ra);
                                if (C::planes == 2) tma_load_2d(sa_lo + i * (BK * 128), &P.a_lo[0], &full[stage], m_blk * BM + i * 64, ra);
                            }
#pragma unroll
                            for (int i = 0; i < BN / 64; ++i) {
                                tma_load_2d(sb_hi + i * (BK * 128), &P.b_hi[sg.b_map], &full[stage], n_blk * BN + i * 64, rb);
                                if (C::planes == 2) tma_load_2d(sb_lo + i * (BK * 128), &P.b_lo[sg.b_map], &full[stage], n_blk * BN + i * 64, rb);
                            }
                        }
                        if (++stage == C::stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc(BN, WGRAD);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int walk = walk_begin; walk < walk_end; walk += walk_step, ++it) {
            const int tile = tile_of(walk);
            int total_kb = 0;
            if (!WGRAD) {
                for (int s = 0; s < P.n_seg; ++s) total_kb += P.seg[s].k_blocks;
            } else {
                const int split = (tile / tiles_mn) % P.split_k;
                int kb0, kb1;
                k_range(split, kb0, kb1);
                total_kb = (kb1 - kb0) * (P.acc_segs ? P.n_seg : 1);
            }
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tempty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < total_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = smem_u32(smem + stage * C::stage_bytes);
                    // K-major: one UMMA_K step = 32 B inside the 128 B swizzle row; MN-major: 16 K-rows = 2048 B
                    constexpr uint32_t kstep = WGRAD ? (UMMA_K * 128) : (UMMA_K * 2);
                    const uint64_t da_hi = WGRAD ? make_smem_desc_mn(st) : make_smem_desc(st);
                    const uint64_t db_hi = WGRAD ? make_smem_desc_mn(st + C::planes * C::a_bytes)
                                                 : make_smem_desc(st + C::planes * C::a_bytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * kstep) >> 4);
                        umma_bf16(tmem_d, da_hi + koff, db_hi + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    if (C::planes == 2) {
                        const uint64_t da_lo = WGRAD ? make_smem_desc_mn(st + C::a_bytes) : make_smem_desc(st + C::a_bytes);
                        const uint64_t db_lo = WGRAD ? make_smem_desc_mn(st + 2 * C::a_bytes + C::b_bytes)
                                                     : make_smem_desc(st + 2 * C::a_bytes + C::b_bytes);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t koff = (uint64_t)((k * kstep) >> 4);
                            umma_bf16(tmem_d, da_lo + koff, db_hi + koff, idesc, 1u);
                            umma_bf16(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                        }
                    }
                    if (CL == 4) tc_commit_mc(&empty[stage], mask_rel);   // release the slot to every CTA that writes into it
                    else tc_commit(&empty[stage]);                  // smem slot reusable once these MMAs retire
                    if (kb == total_kb - 1) tc_commit(&tfull[as]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == C::stages) { stage = 0; phase ^= 1; }
            }
            if (total_kb == 0 && elect_one()) tc_commit(&tfull[as]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------------------------------ epilogue
        const int q = (warp - 4) & 3;                // TMEM lane quarter of this warp (== warp % 4)
        const int half = (warp - 4) >> 2;            // which column half of the tile this warp drains
        constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
        const Stager stager{stage_tiles + (warp - 4) * (32 * kStageLd), lane};
        int it = 0;
        for (int walk = walk_begin; walk < walk_end; walk += walk_step, ++it) {
            const int tile = tile_of(walk);
            int mn = tile, tap = 0, split = 0;
            if (WGRAD) { mn = tile % tiles_mn; const int ts = tile / tiles_mn; tap = ts / P.split_k; split = ts % P.split_k; }
            const int m_blk = mn / P.n_tiles, n_blk = mn % P.n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            bool has_acc = true;
            if (WGRAD) { int kb0, kb1; k_range(split, kb0, kb1); has_acc = kb1 > kb0; }
#pragma unroll 1
            for (int c = half * kColsPerWarp; c < (half + 1) * kColsPerWarp; c += 32) {
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
            mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 4) cluster_sync_all();          // no CTA may exit while a peer can still arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::tmem_cols) : "memory");
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

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int MODE, int KIND, int BN, bool WGRAD, int CL>
static int launch_inst(const TcParams& P, int n_tiles_total, cudaStream_t st) {
    using C = Cfg<MODE, BN>;
    auto kern = gemm_tc_kernel<MODE, KIND, BN, WGRAD, CL>;
    static bool configured = false;
    static int max_clusters = 0;
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
    if (CL == 1) {
        int grid = n_tiles_total < sm_count() ? n_tiles_total : sm_count();
        if (grid < 1) grid = 1;
        kern<<<grid, kThreads, C::smem_bytes, st>>>(P);
        RADMMM_LAUNCH_CHECK();
        return RADMMM_OK;
    }
    int clusters = n_tiles_total / CL;
    if (clusters > max_clusters) clusters = max_clusters;
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    RADMMM_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
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

}  // namespace

int launch_gemm_tc(const GemmArgs& args, int mode, cudaStream_t stream) {
    RADMMM_REQUIRE(mode == MODE_BF16 || mode == MODE_BF16X3, "gemm_tc: bad mode %d", mode);
    RADMMM_REQUIRE(args.R % BM == 0, "gemm_tc: R=%d must be a multiple of %d", args.R, BM);
    RADMMM_REQUIRE(args.n_seg >= 1 && args.n_seg <= kMaxSeg, "gemm_tc: bad segment count %d", args.n_seg);
    const bool x3 = mode == MODE_BF16X3;
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.epi = args.epi;
    P.n_seg = args.n_seg;

    // output tiling
    const int N = args.epi.N;
    const int n_pad = (int)round_up(N, 128);
    const int BN = (!x3 && n_pad % 256 == 0) ? 256 : 128;
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
        use_cl = (P.m_tiles % 2 == 0) && (P.n_tiles % 2 == 0);
        const int a_box = use_cl ? BM / 2 : BM, b_box = use_cl ? BN / 2 : BN;
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
        P.m_tiles = cdiv(M, BM);
        P.acc_segs = args.wgrad == 2;
        P.taps = P.acc_segs ? 1 : args.n_seg;
        P.k_blocks_total = args.R / BK;
        // split K (= rows) so that the persistent grid sees >= ~4 tiles per SM while every tile keeps >= 8 K blocks
        const int tiles0 = P.m_tiles * P.n_tiles * P.taps;
        int split = args.split_k;
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
        P.split_k = split;
        n_tiles_total = P.m_tiles * P.n_tiles * P.taps * P.split_k;
        use_cl = (P.m_tiles % 2 == 0) && (P.n_tiles % 2 == 0);
        RADMMM_REQUIRE(split == 1 || args.epi.atomic, "gemm_tc: split-K weight-grad needs the atomic epilogue");
        // inner extents stop at the logical widths (rounded to the 64-channel box) so that operands which are column
        // slices of wider matrices never read past their rows
        const long long a_in = round_up(M, 64) < g0.a.ld ? round_up(M, 64) : g0.a.ld;
        const long long b_in = round_up(N, 64) < g0.w.ld ? round_up(N, 64) : g0.w.ld;
        RADMMM_TRY(make_map(&P.a_hi[0], g0.a.ptr, a_in, args.R, g0.a.ld, 64));
        if (x3) RADMMM_TRY(make_map(&P.a_lo[0], (const __nv_bfloat16*)g0.a.ptr + g0.a.plane_stride, a_in, args.R, g0.a.ld, 64));
        n_a = 1;
        for (int s = 0; s < args.n_seg; ++s) {
            const GemmSeg& g = args.seg[s];
            RADMMM_REQUIRE(g.a.ptr == g0.a.ptr, "gemm_tc: weight-grad segments must share dY");
            RADMMM_REQUIRE(P.acc_segs || g.w.ptr == g0.w.ptr, "gemm_tc: weight-grad taps must share X");
            int bi = find_or_add(b_keys, n_b, g.w.ptr, g.w.ld, g.w.plane_stride);
            RADMMM_REQUIRE(bi >= 0, "gemm_tc: too many distinct X operands");
            P.seg[s] = TcSeg{0, bi, 0, g.shift, P.k_blocks_total};
        }
        for (int i = 0; i < n_b; ++i) {
            RADMMM_TRY(make_map(&P.b_hi[i], b_keys[i].ptr, b_in, args.R, b_keys[i].ld, 64));
            if (x3) RADMMM_TRY(make_map(&P.b_lo[i], (const __nv_bfloat16*)b_keys[i].ptr + b_keys[i].plane, b_in, args.R, b_keys[i].ld, 64));
        }
    }
    P.n_a_maps = n_a;
    P.n_b_maps = n_b;

    const bool wg = args.wgrad != 0;
    if (use_cl) {
        if (x3) return launch_kind<MODE_BF16X3, 128, 4>(P, n_tiles_total, wg, stream);
        if (BN == 256) return launch_kind<MODE_BF16, 256, 4>(P, n_tiles_total, wg, stream);
        return launch_kind<MODE_BF16, 128, 4>(P, n_tiles_total, wg, stream);
    }
    if (x3) return launch_kind<MODE_BF16X3, 128, 1>(P, n_tiles_total, wg, stream);
    if (BN == 256) return launch_kind<MODE_BF16, 256, 1>(P, n_tiles_total, wg, stream);
    return launch_kind<MODE_BF16, 128, 1>(P, n_tiles_total, wg, stream);
}

}  // namespace radmmm
