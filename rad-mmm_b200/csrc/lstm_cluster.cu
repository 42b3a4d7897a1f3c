// Cluster-resident bidirectional LSTM recurrence on the tensor cores (MODE_BF16), sm_100a.
//
// Replaces, for the throughput mode, the cooperative fp32 kernel of lstm.cu (66 CTAs per direction, W_hh in shared memory,
// h exchanged through L2 behind a global-atomic grid barrier: 3.7 us per time step forward, 5.7 us backward at B=8).
// Here ONE thread-block cluster of 16 CTAs runs a direction (2 clusters = 32 SMs in total; the rest of the GPU stays free
// for the weight preparation that overlaps the LSTM):
//   * W_hh (4H x H) is cut into 16 row slices of 33 hidden units x 4 gates; a CTA keeps its slice as bf16 mma.sync A
//     fragments IN REGISTERS for the whole sequence (18 warps x 17 tiles x 4 registers), so the recurrent mat-vec reads no
//     weights from anywhere (everything else a step needs -- cell state, the input projections of the next step, fetched with
//     cp.async one step ahead -- lives in shared memory, so the fragments are never spilled);
//   * a time step is gates[rows, batch] = W_slice . h: mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with h as the B operand,
//     fetched with ldmatrix.trans from a [k][8 sequences] bf16 tile in shared memory;
//   * the new h slice (33 units x 8 sequences, bf16, 16 bytes per unit) is PUSHED into the shared memory of all 16 CTAs of
//     the cluster with st.async (distributed shared memory) that completes on the receiver's mbarrier -- no cluster barrier,
//     no global memory, no atomics: a CTA starts step s+1 as soon as the 16 slices of h_s have landed (measured exchange
//     floor on B200: 0.76 us per step, tools/probes/cluster_xchg.cu);
//   * backward: the CTA that owns a unit slice also owns its gate gradients, so the transposed mat-vec dh = W^T dG is split
//     over K: every CTA multiplies its 132 gate rows against the full (H x 132) slice of W^T (again register-resident A
//     fragments), and the fp32 partial results are reduce-scattered to the unit owners with 16-byte st.async pushes.
// Variable lengths follow the packed-sequence semantics of the reference (models/radmmm.py:137-146): the reverse direction
// of sequence b starts at frame len_b - 1, frames beyond len_b stay zero.
// Numerics: bf16 W_hh and bf16 h in the recurrent product (fp32 accumulate, fp32 cell state and gate math) -- the precision
// contract of MODE_BF16; the fp32-grade modes keep the fp32 kernel of lstm.cu.  Latency-bound by construction.
#include <stdlib.h>
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int UPC = 33;                // hidden units per CTA
constexpr int MT = 9;                  // m-tiles per CTA: 4 * 33 = 132 gate rows -> 144
constexpr int ROWS = MT * 16;          // 144
constexpr int NT = 640;                // 20 warps (register allocation is per 4 warps: 640 threads -> 96 registers each);
                                       // forward: 18 of them hold W (9 m-tiles x 2 k-halves); backward: KT / 2 of them
// Cluster geometry.  L16 (hidden <= 528: the decoder's context LSTM): 16 CTAs, 34 unit slots per CTA in the K ordering (33 units +
// 1 zero slot: K = 544 = 34 k-tiles).  L4 (hidden <= 132: the attribute predictors' hidden-128 bi-LSTMs): 4 CTAs, 40 slots per CTA
// (K = 160 = 10 k-tiles) -- the same kernels on a quarter of the SMs with 3 instead of 15 exchange partners.
template <int CL_, int SLOT_>
struct LC {
    static constexpr int CL = CL_;             // CTAs per cluster = row slices of W_hh
    static constexpr int SLOT = SLOT_;         // unit slots per CTA in the K ordering (>= UPC; the rest are zero columns)
    static constexpr int KP = CL * SLOT;       // K of the recurrent product
    static constexpr int KT = KP / 16;         // k-tiles
    static constexpr int TPW = KT / 2;         // k-tiles per warp (forward: two K halves)
    static constexpr int HMAX = CL * UPC;
    static_assert(KP % 32 == 0 && TPW % 2 == 1 && SLOT >= UPC, "cluster geometry");
};
using L16 = LC<16, 34>;
using L4 = LC<4, 40>;

struct ClParams {
    const float* xproj;       // [R][8H]  W_ih x + b_ih + b_hh, columns [dir][gate i,f,g,o][H]
    const float* whh[2];      // [4H][H] per direction
    const int* lens;          // [B] grouped lengths
    float* out;               // (B, Tp, 2H), zero-initialised
    float* gates;             // [R][8H] post-activation gates (saved for backward)
    float* cstate;            // [R][2H] cell state (saved for backward)
    const float* dout;        // backward: (B, Tp, 2H)
    float* dgates;            // backward: [R][8H] pre-activation gate gradients
    int B, NB, Tp, H, pitch, n_chunks;     // NB x 8 sequences per cluster, n_chunks clusters per direction
    int probe;                   // tools/lstm_cluster_probe.py (RADMMM_B200_LSTM_PROBE): 1 skip the saved-state stores, 2 skip the gate
                                 // transcendentals, 4 skip the cp.async prefetch
    int bulk;                    // 1: exchange with one bulk DSMEM copy per peer (cp.async.bulk shared::cta -> shared::cluster);
                                 // 0: one 16-byte st.async per (unit, peer)  (RADMMM_B200_LSTM_BULK=0, A/B measurements)
    unsigned long long* trace;   // diagnostic (radmmm_debug_trace): per CTA 8 accumulated SM-clock counters, see kernels
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
        if (++spins > (1ll << 24)) {     // an exchange bug must trap, never hang the device
            printf("radmmm lstm_cluster: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// 16-byte push into the shared memory of another CTA of the cluster; completes 16 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t remote_bar, uint4 v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into another CTA's; completes on that CTA's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (the bulk-copy engine) after the next barrier
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// K ordering shared by forward and backward: slot p in [0, 544) belongs to CTA p / 34; its unit is 33 * (p / 34) + p % 34 when
// p % 34 < 33 and that unit exists, else the slot is a zero column.
template <class C>
__device__ __forceinline__ int slot_unit(int p, int H) {
    const int j = p % C::SLOT, u = (p / C::SLOT) * UPC + j;
    return (j < UPC && u < H) ? u : -1;
}
// local gate row r in [0, 144): unit j = r / 4 of this CTA, gate r % 4 (i, f, g, o); rows >= 132 are padding
template <class C>
__device__ __forceinline__ float whh_at(const float* __restrict__ Whh, int H, int rank, int r, int p) {
    const int j = r >> 2, g = r & 3, unit = rank * UPC + j, col = slot_unit<C>(p, H);
    if (j >= UPC || unit >= H || col < 0) return 0.0f;
    return __ldg(Whh + (size_t)(g * H + unit) * H + col);
}

// =========================================================================================================== forward
// shared memory: [2 mbarriers | lens[32] | hs[2][NB][544][8] bf16 | part[2][144][8 NB] fp32 | hstage[NB][34][8] bf16 |
//                 (x2) | cst[33 * 8 NB] fp32 | xps[3][4][33 * 8 NB] fp32 | itab[33 * 8 NB] int4 | stash[6][33 * 8 NB] fp32]
template <bool TRACE, class C>
__global__ void __cluster_dims__(C::CL, 1, 1) __launch_bounds__(NT, 1) lstm_cl_fwd_kernel(const ClParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int H = p.H, NB = p.NB, NBN = 8 * NB, NIT = UPC * NBN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    int* slen = reinterpret_cast<int*>(smem + 128);                                          // [32]
    __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(smem + 256);                       // [2][NB][KP][8]
    float* part = reinterpret_cast<float*>(smem + 256 + (size_t)2 * NB * C::KP * 16);           // [2][ROWS][NBN]
    // hstage is double-buffered: the bulk copies of step s may still be reading it while step s+1 writes (they are known to
    // be complete once h_{s+1} of every peer has arrived, i.e. before step s+2 writes the same half again)
    __nv_bfloat16* hstage2 = reinterpret_cast<__nv_bfloat16*>(part + (size_t)2 * ROWS * NBN); // [2][NB][SLOT][8]
    float* cst = reinterpret_cast<float*>(hstage2 + (size_t)2 * NB * C::SLOT * 8);              // [NIT]
    float* xps = cst + NIT;                                                                   // [3][4][NIT]
    int4* itab = reinterpret_cast<int4*>(xps + (size_t)3 * 4 * NIT);                          // [NIT]
    float* stash = reinterpret_cast<float*>(itab + NIT);                                      // [6][NIT] values saved for backward
    const int rank = (int)cluster_rank(), dir = (int)(cluster_id_x() & 1), n0 = (int)(cluster_id_x() >> 1) * 8;     // cluster = (8-sequence chunk, direction)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool mma_warp = warp < 2 * MT;      // 18 of the 20 warps
    const int mt = warp >> 1, kh = warp & 1;
    const float* Whh = p.whh[dir];

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) slen[tid] = (tid < 8 * NB && n0 + tid < p.B) ? min(p.lens[n0 + tid], p.Tp) : 0;
    for (int i = tid; i < 2 * NB * C::KP; i += NT) reinterpret_cast<uint4*>(hs)[i] = make_uint4(0, 0, 0, 0);   // h_{-1} = 0, zero slots
    for (int i = tid; i < NIT; i += NT) cst[i] = 0.0f;
    // per-item constants (item it = j * NBN + n: unit j of this CTA, sequence n), so that the per-step gate code does no index
    // arithmetic: .x = length (0: nothing to do), .y = offset of (row of frame 0, this direction, this unit) in xproj / gates
    // ([R][8H]), .z = the same in cstate ([R][2H]), .w = offset of (frame 0) in out ((B, Tp, 2H))
    for (int it = tid; it < NIT; it += NT) {
        const int j = it / NBN, n = it - j * NBN, unit = rank * UPC + j;
        int4 e;
        const int ng = n0 + n;                       // sequence index in the batch
        e.x = (ng < p.B && unit < H) ? min(p.lens[ng], p.Tp) : 0;
        e.y = ng * p.pitch * 8 * H + dir * 4 * H + unit;
        e.z = ng * p.pitch * 2 * H + dir * H + unit;
        e.w = ng * p.Tp * 2 * H + dir * H + unit;
        itab[it] = e;
    }
    // W_hh slice -> A fragments (rows = local gate rows of m-tile `mt`, k = slots of this warp's K half)
    uint32_t wf[C::TPW][4];
    if (mma_warp) {
        const int r0 = mt * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
        for (int i = 0; i < C::TPW; ++i) {
            const int c0 = (kh * C::TPW + i) * 16 + 2 * (lane & 3);
            wf[i][0] = pack_bf16(whh_at<C>(Whh, H, rank, r0, c0), whh_at<C>(Whh, H, rank, r0, c0 + 1));
            wf[i][1] = pack_bf16(whh_at<C>(Whh, H, rank, r1, c0), whh_at<C>(Whh, H, rank, r1, c0 + 1));
            wf[i][2] = pack_bf16(whh_at<C>(Whh, H, rank, r0, c0 + 8), whh_at<C>(Whh, H, rank, r0, c0 + 9));
            wf[i][3] = pack_bf16(whh_at<C>(Whh, H, rank, r1, c0 + 8), whh_at<C>(Whh, H, rank, r1, c0 + 9));
        }
    }
    int tmax = 0;                                  // this chunk's longest sequence
    for (int b = n0; b < min(p.B, n0 + 8 * NB); ++b) tmax = max(tmax, min(p.lens[b], p.Tp));
    __syncthreads();
    cluster_sync_all();                       // every CTA's barriers and zeroed h tiles exist before the first push

    // pointwise work items: it = j * NBN + n (unit j of this CTA, sequence n), thread tid handles it = tid, tid + NT, ...
    // the input projections of step s are fetched into xps[s % 3] with cp.async TWO steps ahead (issued after the push of
    // step s - 2, so neither their issue nor their latency sits on a step's critical path)
    auto fetch_xp = [&](int s) {
        for (int it = tid; it < NIT; it += NT) {
            const int4 e = itab[it];
            if (s < e.x) {
                const int t = dir ? e.x - 1 - s : s;
                const float* x = p.xproj + (size_t)e.y + (size_t)t * 8 * H;
                float* d = xps + (size_t)(s % 3) * 4 * NIT + it;
#pragma unroll
                for (int g = 0; g < 4; ++g) cp_async4(d + (size_t)g * NIT, x + (size_t)g * H);
            }
        }
        cp_async_commit();
    };
    fetch_xp(0);
    fetch_xp(1);
    const uint32_t tx_bytes = (uint32_t)(C::CL * UPC * NB * 16);
    // diagnostic phase timers of thread 0: [0] wait for h, [1] mat-vec, [2] cp.async wait + barrier, [3] gates, [4] barrier, [5] push
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
    auto lap = [&](int i) { if (TRACE && tid == 0) { const long long now = clock64(); tacc[i] += now - tlast; tlast = now; } };

    for (int s = 0; s < tmax; ++s) {
        const int cur = s & 1, prev = cur ^ 1;
        __nv_bfloat16* hstage = hstage2 + (size_t)cur * NB * C::SLOT * 8;
        lap(5);
        if (s > 0) mbar_wait(&bars[prev], ((s - 1) >> 1) & 1);         // the 16 slices of h_{s-1} have landed in hs[prev]
        lap(0);
        // ---- recurrent mat-vec on the tensor cores: part[kh][row][n] = sum_{k in half kh} W[row][k] h_{s-1}[k][n]
        if (mma_warp) {
            for (int nb = 0; nb < NB; ++nb) {
                const uint32_t hbase = smem_u32(hs + ((size_t)(prev * NB + nb) * C::KP + (size_t)kh * C::TPW * 16) * 8);
                float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i + 1 < C::TPW; i += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(hbase + (uint32_t)(i * 16 + lane) * 16, b0, b1, b2, b3);
                    mma_bf16(acc0, wf[i], b0, b1);
                    mma_bf16(acc1, wf[i + 1], b2, b3);
                }
                {
                    uint32_t b0, b1;
                    ldmatrix_x2_trans(hbase + (uint32_t)((C::TPW - 1) * 16 + (lane & 15)) * 16, b0, b1);
                    mma_bf16(acc0, wf[C::TPW - 1], b0, b1);
                }
                float* o = part + ((size_t)kh * ROWS + mt * 16 + (lane >> 2)) * NBN + nb * 8 + 2 * (lane & 3);
                *reinterpret_cast<float2*>(o) = make_float2(acc0[0] + acc1[0], acc0[1] + acc1[1]);
                *reinterpret_cast<float2*>(o + 8 * NBN) = make_float2(acc0[2] + acc1[2], acc0[3] + acc1[3]);
            }
        }
        lap(1);
        cp_async_wait_but_one();               // this thread's input projections of step s are in xps[s % 3]
        __syncthreads();
        lap(2);
        // ---- gates, cell / hidden state of this CTA's units.  What the backward pass needs (gates, c, h: six scattered global
        // stores per item) is only parked in shared memory here and written out AFTER the push, while the exchange is in
        // flight -- off the step's critical path; so is the prefetch of the next step's input projections.
        for (int it = tid; it < NIT; it += NT) {
            const int4 e = itab[it];
            const int j = it / NBN, n = it - j * NBN;
            float h = 0.0f;
            if (s < e.x) {
                const float* pa = part + (size_t)(4 * j) * NBN + n;
                const float* pb = pa + (size_t)ROWS * NBN;
                const float* xp = xps + (size_t)(s % 3) * 4 * NIT + it;
                float gi = pa[0] + pb[0] + xp[0], gf = pa[NBN] + pb[NBN] + xp[NIT];
                float gg = pa[2 * NBN] + pb[2 * NBN] + xp[2 * NIT], go = pa[3 * NBN] + pb[3 * NBN] + xp[3 * NIT];
                if (!(p.probe & 2)) { gi = sigmoidf_(gi); gf = sigmoidf_(gf); gg = tanhf_(gg); go = sigmoidf_(go); }
                const float c = gf * cst[it] + gi * gg;
                h = go * ((p.probe & 2) ? c : tanhf_(c));
                cst[it] = c;
                float* sp = stash + it;
                sp[0] = gi; sp[NIT] = gf; sp[2 * NIT] = gg; sp[3 * NIT] = go; sp[4 * NIT] = c; sp[5 * NIT] = h;
            }
            hstage[((size_t)(n >> 3) * C::SLOT + j) * 8 + (n & 7)] = __float2bfloat16_rn(h);
        }
        auto write_saved = [&]() {             // each thread writes what it parked itself: no barrier needed
            if (p.probe & 1) return;
            for (int it = tid; it < NIT; it += NT) {
                const int4 e = itab[it];
                if (s < e.x) {
                    const int t = dir ? e.x - 1 - s : s;
                    const float* sp = stash + it;
                    float* gp = p.gates + (size_t)e.y + (size_t)t * 8 * H;
                    gp[0] = sp[0]; gp[H] = sp[NIT]; gp[2 * (size_t)H] = sp[2 * NIT]; gp[3 * (size_t)H] = sp[3 * NIT];
                    p.cstate[(size_t)e.z + (size_t)t * 2 * H] = sp[4 * NIT];
                    p.out[(size_t)e.w + (size_t)t * 2 * H] = sp[5 * NIT];
                }
            }
        };
        if (s + 1 == tmax) { write_saved(); break; }
        lap(3);
        if (p.bulk) fence_proxy_async();       // hstage was written through the generic proxy, the copy engine reads it
        __syncthreads();
        lap(4);
        // ---- push this CTA's h slice into every CTA of the cluster (own copy included)
        if (tid == 0) mbar_expect_tx(&bars[cur], tx_bytes);
        if (p.bulk) {                          // ONE bulk copy of 33 units x 16 bytes per (tile, peer)
            if (tid < C::CL * NB) {
                const int peer = tid % C::CL, nb = tid / C::CL;
                const uint32_t dst = smem_u32(hs + ((size_t)(cur * NB + nb) * C::KP + rank * C::SLOT) * 8);
                bulk_copy_to_peer(mapa(dst, peer), smem_u32(hstage + (size_t)nb * C::SLOT * 8), UPC * 16, mapa(smem_u32(&bars[cur]), peer));
            }
        } else {                               // 16 bytes per (unit, tile, peer)
            for (int i = tid; i < UPC * NB * C::CL; i += NT) {
                const int peer = i % C::CL, rest = i / C::CL, j = rest % UPC, nb = rest / UPC;
                const uint4 v = *reinterpret_cast<const uint4*>(hstage + ((size_t)nb * C::SLOT + j) * 8);
                const uint32_t dst = smem_u32(hs + ((size_t)(cur * NB + nb) * C::KP + rank * C::SLOT + j) * 8);
                st_async_v4(mapa(dst, peer), mapa(smem_u32(&bars[cur]), peer), v);
            }
        }
        write_saved();
        if (!(p.probe & 4)) fetch_xp(s + 2);   // issued while this CTA waits for the other 15 slices, needed two steps from now
        else cp_async_commit();
    }
    cp_async_wait_all();
    cluster_sync_all();                       // no CTA leaves while a peer could still push into its shared memory
    if (TRACE && p.trace != nullptr && tid == 0 && blockIdx.x < 2 * C::CL)          // the trace buffer holds the first two clusters
        for (int i = 0; i < 6; ++i) p.trace[blockIdx.x * 8 + i] = (unsigned long long)tacc[i];
}

// =========================================================================================================== backward
// Work split: CTA `rank` owns the 33 units [33 rank, 33 rank + 33) -- their dh, dc and gate gradients -- and the matching 132
// gate ROWS of W_hh.  dh_rec[unit, n] = sum over all 4H gate rows rho of W_hh[rho][unit] dG[rho][n] is split over K = rho:
// every CTA computes the partial sum over ITS 132 rows for ALL 544 unit slots (A = the transposed slice, 34 m-tiles x 9
// k-tiles: warp w < 17 holds m-tiles 2w and 2w+1 with the full K), and pushes each 16-unit x 8-sequence fp32 tile to the
// owner of those units; the owner adds the 16 partial slices.
// shared memory: [2 mbarriers | lens[32] | recv[2][CL][NB][SLOT][8] fp32 | dgs[NB][144][8] bf16 | dcn[33 * 8 NB] fp32 |
//                 sv[3][7][33 * 8 NB] fp32 | itab[33 * 8 NB] int4 | stash[4][33 * 8 NB] fp32 |
//                 pstage[2][NB][544][8] fp32 (bulk exchange only)]
template <bool TRACE, class C>
__global__ void __cluster_dims__(C::CL, 1, 1) __launch_bounds__(NT, 1) lstm_cl_bwd_kernel(const ClParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int H = p.H, NB = p.NB, NBN = 8 * NB, NIT = UPC * NBN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    int* slen = reinterpret_cast<int*>(smem + 128);
    float* recv = reinterpret_cast<float*>(smem + 256);                                      // [2][CL][NB][SLOT][8]
    __nv_bfloat16* dgs = reinterpret_cast<__nv_bfloat16*>(recv + (size_t)2 * C::CL * NB * C::SLOT * 8);   // [NB][ROWS][8]
    float* dcn = reinterpret_cast<float*>(dgs + (size_t)NB * ROWS * 8);                      // [NIT]
    float* sv = dcn + NIT;                                                                    // [3][7][NIT]
    int4* itab = reinterpret_cast<int4*>(sv + (size_t)3 * 7 * NIT);                           // [NIT]
    float* stash = reinterpret_cast<float*>(itab + NIT);                                      // [4][NIT] gate gradients for dgates
    float* pstage2 = stash + (size_t)4 * NIT;                                                 // [2][NB][KP][8]
    const int rank = (int)cluster_rank(), dir = (int)(cluster_id_x() & 1), n0 = (int)(cluster_id_x() >> 1) * 8;     // cluster = (8-sequence chunk, direction)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* Whh = p.whh[dir];
    constexpr int MTW = 2;                    // m-tiles (of 16 unit slots) per warp
    const bool mma_warp = warp < C::KT / MTW;    // 17 warps

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) slen[tid] = (tid < 8 * NB && n0 + tid < p.B) ? min(p.lens[n0 + tid], p.Tp) : 0;
    for (int i = tid; i < NB * ROWS; i += NT) reinterpret_cast<uint4*>(dgs)[i] = make_uint4(0, 0, 0, 0);   // padding rows stay zero
    for (int i = tid; i < NIT; i += NT) dcn[i] = 0.0f;
    for (int it = tid; it < NIT; it += NT) {      // per-item constants, as in the forward kernel (.y also addresses dgates)
        const int j = it / NBN, n = it - j * NBN, unit = rank * UPC + j;
        int4 e;
        const int ng = n0 + n;                       // sequence index in the batch
        e.x = (ng < p.B && unit < H) ? min(p.lens[ng], p.Tp) : 0;
        e.y = ng * p.pitch * 8 * H + dir * 4 * H + unit;
        e.z = ng * p.pitch * 2 * H + dir * H + unit;
        e.w = ng * p.Tp * 2 * H + dir * H + unit;
        itab[it] = e;
    }
    // A fragments of the transposed slice: A[m = unit slot][k = local gate row] = W_hh[row(k)][unit(m)]
    uint32_t wf[MTW][MT][4];
    if (mma_warp) {
#pragma unroll
        for (int a = 0; a < MTW; ++a) {
            const int m0 = (warp * MTW + a) * 16 + (lane >> 2), m1 = m0 + 8;
#pragma unroll
            for (int k = 0; k < MT; ++k) {
                const int c0 = k * 16 + 2 * (lane & 3);
                wf[a][k][0] = pack_bf16(whh_at<C>(Whh, H, rank, c0, m0), whh_at<C>(Whh, H, rank, c0 + 1, m0));
                wf[a][k][1] = pack_bf16(whh_at<C>(Whh, H, rank, c0, m1), whh_at<C>(Whh, H, rank, c0 + 1, m1));
                wf[a][k][2] = pack_bf16(whh_at<C>(Whh, H, rank, c0 + 8, m0), whh_at<C>(Whh, H, rank, c0 + 9, m0));
                wf[a][k][3] = pack_bf16(whh_at<C>(Whh, H, rank, c0 + 8, m1), whh_at<C>(Whh, H, rank, c0 + 9, m1));
            }
        }
    }
    int tmax = 0;                                  // this chunk's longest sequence
    for (int b = n0; b < min(p.B, n0 + 8 * NB); ++b) tmax = max(tmax, min(p.lens[b], p.Tp));
    __syncthreads();
    cluster_sync_all();

    // saved forward values of (unit, sequence) for time step s: gates i f g o, c, c_prev, dout -> sv[s % 3], two steps ahead
    auto load_saved = [&](int s) {
        for (int it = tid; it < NIT; it += NT) {
            const int4 e = itab[it];
            if (s >= 0 && s < e.x) {
                const int t = dir ? e.x - 1 - s : s;
                const float* gp = p.gates + (size_t)e.y + (size_t)t * 8 * H;
                const float* cp = p.cstate + (size_t)e.z + (size_t)t * 2 * H;
                float* d = sv + (size_t)(s % 3) * 7 * NIT + it;
#pragma unroll
                for (int g = 0; g < 4; ++g) cp_async4(d + (size_t)g * NIT, gp + (size_t)g * H);
                cp_async4(d + (size_t)4 * NIT, cp);
                if (s > 0) cp_async4(d + (size_t)5 * NIT, dir ? cp + 2 * (size_t)H : cp - 2 * (size_t)H);     // c of the previous step
                cp_async4(d + (size_t)6 * NIT, p.dout + (size_t)e.w + (size_t)t * 2 * H);
            }
        }
        cp_async_commit();
    };
    load_saved(tmax - 1);
    load_saved(tmax - 2);
    const uint32_t tx_bytes = (uint32_t)(C::CL * NB * C::SLOT * 8 * 4);          // 16 sources x (34 slots x 8 sequences) fp32 per tile
    // diagnostic phase timers of thread 0: [0] wait for dh, [1] cp.async wait, [2] gate gradients, [3] barrier, [4] mat-vec + push
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
    auto lap = [&](int i) { if (TRACE && tid == 0) { const long long now = clock64(); tacc[i] += now - tlast; tlast = now; } };

    for (int s = tmax - 1, step = 0; s >= 0; --s, ++step) {
        const int cur = step & 1, prev = cur ^ 1;
        lap(4);
        if (step > 0) mbar_wait(&bars[prev], ((step - 1) >> 1) & 1);      // the 16 partial slices of dh_rec have landed
        lap(0);
        cp_async_wait_but_one();
        lap(1);
        // ---- gate gradients of this CTA's units
        for (int it = tid; it < NIT; it += NT) {
            const int4 e = itab[it];
            const int j = it / NBN, n = it - j * NBN;
            float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f, dc_keep = 0.0f;
            if (s < e.x) {
                const float* v = sv + (size_t)(s % 3) * 7 * NIT + it;
                float dh = v[6 * NIT];
                if (step > 0) {
                    const float* rc = recv + (((size_t)prev * C::CL * NB + (n >> 3)) * C::SLOT + j) * 8 + (n & 7);
#pragma unroll
                    for (int src = 0; src < C::CL; ++src) dh += rc[(size_t)src * NB * C::SLOT * 8];
                }
                const float gi = v[0], gf = v[NIT], gg = v[2 * NIT], go = v[3 * NIT];
                const float c_prev = s > 0 ? v[5 * NIT] : 0.0f;
                const float tc = (p.probe & 2) ? v[4 * NIT] : tanhf_(v[4 * NIT]);
                const float dc = dh * go * (1.0f - tc * tc) + dcn[it];
                d_o = dh * tc * go * (1.0f - go);
                d_i = dc * gg * gi * (1.0f - gi);
                d_g = dc * gi * (1.0f - gg * gg);
                d_f = dc * c_prev * gf * (1.0f - gf);
                dc_keep = dc * gf;
                float* sp = stash + it;        // written to dgates after the exchange has been started (see write_saved)
                sp[0] = d_i; sp[NIT] = d_f; sp[2 * NIT] = d_g; sp[3 * NIT] = d_o;
            }
            dcn[it] = dc_keep;
            __nv_bfloat16* o = dgs + ((size_t)(n >> 3) * ROWS + 4 * j) * 8 + (n & 7);
            o[0] = __float2bfloat16_rn(d_i); o[8] = __float2bfloat16_rn(d_f);
            o[16] = __float2bfloat16_rn(d_g); o[24] = __float2bfloat16_rn(d_o);
        }
        auto write_saved = [&]() {             // each thread writes what it parked itself: no barrier needed
            if (p.probe & 1) return;
            for (int it = tid; it < NIT; it += NT) {
                const int4 e = itab[it];
                if (s < e.x) {
                    const int t = dir ? e.x - 1 - s : s;
                    const float* sp = stash + it;
                    float* dg = p.dgates + (size_t)e.y + (size_t)t * 8 * H;
                    dg[0] = sp[0]; dg[H] = sp[NIT]; dg[2 * (size_t)H] = sp[2 * NIT]; dg[3 * (size_t)H] = sp[3 * NIT];
                }
            }
        };
        if (s == 0) { write_saved(); break; }
        lap(2);
        __syncthreads();
        lap(3);
        // ---- partial dh_rec for all unit slots, pushed to the owners
        if (tid == 0) mbar_expect_tx(&bars[cur], tx_bytes);
        if (mma_warp) {
            for (int nb = 0; nb < NB; ++nb) {
                const uint32_t gbase = smem_u32(dgs + (size_t)nb * ROWS * 8);
                float acc[MTW][4];
#pragma unroll
                for (int a = 0; a < MTW; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.0f;
#pragma unroll
                for (int k = 0; k + 1 < MT; k += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(gbase + (uint32_t)(k * 16 + lane) * 16, b0, b1, b2, b3);
#pragma unroll
                    for (int a = 0; a < MTW; ++a) { mma_bf16(acc[a], wf[a][k], b0, b1); mma_bf16(acc[a], wf[a][k + 1], b2, b3); }
                }
                {
                    uint32_t b0, b1;
                    ldmatrix_x2_trans(gbase + (uint32_t)((MT - 1) * 16 + (lane & 15)) * 16, b0, b1);
#pragma unroll
                    for (int a = 0; a < MTW; ++a) mma_bf16(acc[a], wf[a][MT - 1], b0, b1);
                }
                if (p.bulk) {
                    // C fragment (rows lane/4 and +8, sequences 2 (lane%4) + {0,1}) -> local staging tile [slot][8] fp32
                    float* ps = pstage2 + ((size_t)(cur * NB + nb) * C::KP) * 8;
#pragma unroll
                    for (int a = 0; a < MTW; ++a) {
                        float* o = ps + (size_t)((warp * MTW + a) * 16 + (lane >> 2)) * 8 + 2 * (lane & 3);
                        *reinterpret_cast<float2*>(o) = make_float2(acc[a][0], acc[a][1]);
                        *reinterpret_cast<float2*>(o + 64) = make_float2(acc[a][2], acc[a][3]);
                    }
                } else {
                    // Lane pairs (xor 1) trade halves so that the even lane owns 4 consecutive sequences of row lane/4 and the odd
                    // lane those of row lane/4 + 8: one 16-byte push each.
#pragma unroll
                    for (int a = 0; a < MTW; ++a) {
                        const bool odd = lane & 1;
                        const float s0 = odd ? acc[a][0] : acc[a][2], s1 = odd ? acc[a][1] : acc[a][3];
                        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                        float4 v;
                        if (!odd) v = make_float4(acc[a][0], acc[a][1], r0, r1);        // row lane/4
                        else v = make_float4(r0, r1, acc[a][2], acc[a][3]);             // row lane/4 + 8
                        const int slot = (warp * MTW + a) * 16 + (lane >> 2) + (odd ? 8 : 0);
                        const int n0 = 4 * ((lane & 3) >> 1);                            // lanes 0,1 -> sequences 0..3; lanes 2,3 -> 4..7
                        const int owner = slot / C::SLOT, js = slot % C::SLOT;
                        const uint32_t dst = smem_u32(recv + ((((size_t)cur * C::CL + rank) * NB + nb) * C::SLOT + js) * 8 + n0);
                        st_async_v4(mapa(dst, owner), mapa(smem_u32(&bars[cur]), owner), *reinterpret_cast<uint4*>(&v));
                    }
                }
            }
        }
        if (p.bulk) {       // the staged partial sums go to their owners: ONE bulk copy of 34 slots x 32 bytes per (tile, owner)
            fence_proxy_async();
            __syncthreads();
            if (tid < C::CL * NB) {
                const int owner = tid % C::CL, nb = tid / C::CL;
                const uint32_t src = smem_u32(pstage2 + ((size_t)(cur * NB + nb) * C::KP + owner * C::SLOT) * 8);
                const uint32_t dst = smem_u32(recv + (((size_t)cur * C::CL + rank) * NB + nb) * C::SLOT * 8);
                bulk_copy_to_peer(mapa(dst, owner), src, C::SLOT * 32, mapa(smem_u32(&bars[cur]), owner));
            }
        }
        write_saved();
        if (!(p.probe & 4)) load_saved(s - 2); // issued while this CTA waits for the other partial sums, needed two steps from now
        else cp_async_commit();
    }
    cp_async_wait_all();
    cluster_sync_all();
    if (TRACE && p.trace != nullptr && tid == 0 && blockIdx.x < 2 * C::CL)          // the trace buffer holds the first two clusters
        for (int i = 0; i < 6; ++i) p.trace[blockIdx.x * 8 + i] = (unsigned long long)tacc[i];
}

}  // namespace



static unsigned long long* g_lstm_trace = nullptr;
void lstm_cluster_set_trace(void* buf) { g_lstm_trace = reinterpret_cast<unsigned long long*>(buf); }

bool lstm_cluster_supported(int B, int H) { return H >= 1 && H <= L16::HMAX && B >= 1 && B <= 256; }

// RADMMM_B200_LSTM_CL4=0 keeps every hidden size on the 16-CTA geometry (A/B measurements)
static bool use_small_geometry(int H) {
    static const bool allow = []() { const char* e = getenv("RADMMM_B200_LSTM_CL4"); return !(e && e[0] == '0'); }();
    return allow && H <= L4::HMAX;
}

template <class C>
static size_t fwd_smem(int NB) {
    const size_t nit = (size_t)UPC * 8 * NB;
    return 256 + (size_t)2 * NB * C::KP * 16 + (size_t)2 * ROWS * 8 * NB * 4 + (size_t)2 * NB * C::SLOT * 16 + nit * 4 + 3 * 4 * nit * 4 + nit * 16 + 6 * nit * 4;
}
template <class C>
static size_t bwd_smem(int NB) {
    const size_t nit = (size_t)UPC * 8 * NB;
    return 256 + (size_t)2 * C::CL * NB * C::SLOT * 8 * 4 + (size_t)NB * ROWS * 16 + nit * 4 + 3 * 7 * nit * 4 + nit * 16 + 4 * nit * 4 +
           (size_t)2 * NB * C::KP * 8 * 4;
}

template <bool FWD, bool TRACE, class C>
static int launch_cluster_t(const ClParams& p, cudaStream_t st) {
    static size_t smem_set[64] = {};          // function attributes are per device (and per instantiation)
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    const size_t smem = FWD ? fwd_smem<C>(p.NB) : bwd_smem<C>(p.NB);
    auto kern = FWD ? lstm_cl_fwd_kernel<TRACE, C> : lstm_cl_bwd_kernel<TRACE, C>;
    if (smem > smem_set[dev]) {
        if (C::CL > 8) RADMMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        RADMMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev] = smem;
    }
    kern<<<2 * C::CL * p.n_chunks, NT, smem, st>>>(p);       // one cluster per (chunk, direction); independent of each other
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

template <bool FWD>
static int launch_cluster(const ClParams& p_in, cudaStream_t st) {
    // exchange flavour (measured on B200 at B=8, T'=400, profiles/): forward 2.50 us/step with one bulk DSMEM copy per peer
    // vs 2.84 with 16-byte st.async pushes; backward 2.94 vs 2.63 (its payload comes straight from the accumulator registers,
    // the bulk copy needs a staging round trip through shared memory).  RADMMM_B200_LSTM_BULK=0/1 forces one for both.
    static const int forced = []() { const char* e = getenv("RADMMM_B200_LSTM_BULK"); return e ? (e[0] != '0' ? 1 : 0) : -1; }();
    ClParams p = p_in;
    p.bulk = forced >= 0 ? forced : (FWD ? 1 : 0);
    if (use_small_geometry(p.H))
        return p.trace != nullptr ? launch_cluster_t<FWD, true, L4>(p, st) : launch_cluster_t<FWD, false, L4>(p, st);
    return p.trace != nullptr ? launch_cluster_t<FWD, true, L16>(p, st) : launch_cluster_t<FWD, false, L16>(p, st);
}

static int fill(ClParams& p, const int* lens, int B, int Tp, int H) {
    RADMMM_REQUIRE(lstm_cluster_supported(B, H), "lstm_cluster: B=%d (<= 256) / H=%d (<= %d) out of range", B, H, L16::HMAX);
    memset(&p, 0, sizeof(p));
    // one cluster per (8 sequences, direction): a step costs the same for every batch size as long as the clusters are
    // co-resident (148 SMs hold 9 clusters of 16 CTAs, i.e. 32 sequences run in one wave, 64 in two)
    p.lens = lens; p.B = B; p.NB = 1; p.n_chunks = (B + 7) / 8; p.Tp = Tp; p.H = H; p.pitch = Tp + 16;
    p.trace = g_lstm_trace;
    p.bulk = -1;                                // decided per kernel in launch_cluster
    { const char* e = getenv("RADMMM_B200_LSTM_PROBE"); p.probe = e ? atoi(e) : 0; }
    return RADMMM_OK;
}

int lstm_cluster_forward(const float* xproj, const float* whh_f, const float* whh_r, const int* lens, int B, int Tp, int H,
                         float* out, float* gates, float* cstate, cudaStream_t st) {
    ClParams p;
    RADMMM_TRY(fill(p, lens, B, Tp, H));
    p.xproj = xproj; p.whh[0] = whh_f; p.whh[1] = whh_r; p.out = out; p.gates = gates; p.cstate = cstate;
    return launch_cluster<true>(p, st);
}

int lstm_cluster_backward(const float* dout, const float* gates, const float* cstate, const float* whh_f, const float* whh_r,
                          const int* lens, int B, int Tp, int H, float* dgates, cudaStream_t st) {
    ClParams p;
    RADMMM_TRY(fill(p, lens, B, Tp, H));
    p.dout = dout; p.gates = const_cast<float*>(gates); p.cstate = const_cast<float*>(cstate);
    p.whh[0] = whh_f; p.whh[1] = whh_r; p.dgates = dgates;
    return launch_cluster<false>(p, st);
}

}  // namespace radmmm
