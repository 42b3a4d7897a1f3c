// Persistent bidirectional LSTM recurrence for the decoder's context encoder (models/radmmm.py:137-146:
// pack_padded_sequence -> nn.LSTM(bidirectional) -> pad_packed_sequence).
//
// cuDNN runs this fp32 LSTM as ~10 launch-bound kernels per time step (measured: ~4000 launches, ~45 % of the train
// step at B=8, T'=400).  Here the input projections for all frames are one contraction (gemm_tc / gemm_ffma) and the
// sequential part is ONE cooperative kernel per pass:
//   * one CTA per (direction, 8 hidden units); its 32 x H slice of W_hh stays in shared memory for the whole sequence;
//   * per step each warp owns one hidden unit: lanes split the H-long dot products (4 gates x 8 sequences = 32 FMAs per
//     3 shared loads), a halving butterfly leaves lane v with the sum for (gate v/8, sequence v%8), lanes 0-7 apply the
//     gate non-linearities and keep c in a register;
//   * the new h slice goes to a double-buffered exchange array in L2 and a per-direction arrive/spin barrier
//     (one atomic per CTA per step) publishes it to the other CTAs of that direction.
// Both directions run concurrently; variable lengths follow the packed semantics (the reverse direction of sequence b
// starts at frame len_b-1; frames beyond len_b stay zero).  All arithmetic fp32.
// Latency-bound by construction: T' dependent steps of (17 KB exchange + barrier); no roofline applies.
#include <stdlib.h>
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int U = 8;          // hidden units per CTA (one per warp)
constexpr int NT = 256;
constexpr int BT = 8;         // sequences per register tile

struct LstmParams {
    const float* xproj;       // [R][8H]  W_ih x + b_ih + b_hh, columns [dir][gate i,f,g,o][H]
    const float* whh[2];      // [4H][H] per direction
    const int* lens;          // [B] grouped lengths
    float* out;               // (B, Tp, 2H)
    float* gates;             // [R][8H] post-activation gates (saved for backward)
    float* cstate;            // [R][2H] cell state (saved for backward)
    const float* dout;        // backward: (B, Tp, 2H)
    float* dgates;            // backward: [R][8H] pre-activation gate gradients
    float* xchg;              // exchange buffers
    unsigned int* counters;   // [2] per-direction barrier counters (zeroed before launch)
    int B, Bp, Tp, H, pitch, G;
    int probe;                // tools/lstm_probe.py: 1 skip the mat-vec, 2 skip the exchange reload, 4 skip the barrier
};

// Gate non-linearities on the MUFU exponential (abs error ~1e-7, far inside the fp32 parity tolerance); tanh through
// exp(-2|x|) so that it never overflows.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// L2 -> shared copy of n4 float4 with up to BATCH loads per thread in flight (the exchange buffers were just written by
// the other CTAs, so every load is an L2 round trip: issuing them back to back instead of load/store/load/store cuts the
// reload from ~n4/NT round trips to ~n4/(NT*BATCH); measured 35 % of the forward step before).  `src_index(i)` maps
// the destination float4 index to the source float4 index.
template <int BATCH, typename F>
__device__ __forceinline__ void copy_f4_batched(float4* __restrict__ dst, const float4* __restrict__ src, int n4, int tid,
                                                F src_index) {
    for (int base = 0; base < n4; base += BATCH * 256) {
        float4 t[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = base + j * 256 + tid;
            if (i < n4) t[j] = __ldcg(src + src_index(i));
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = base + j * 256 + tid;
            if (i < n4) dst[i] = t[j];
        }
    }
}

// Per-direction arrive/spin barrier.  The CTA's stores are ordered before thread 0's release-reduction by the bar.sync
// (cumulativity), so no per-thread MEMBAR is needed; consumers read the exchanged data after the acquire + bar.sync.
__device__ __forceinline__ void dir_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        long long spins = 0;
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (++spins > (1ll << 26)) { printf("radmmm lstm: barrier timed out\n"); __trap(); }
        }
    }
    __syncthreads();
}

// halving butterfly: in: v[32] partial sums per lane; out: the full sum of element `lane` (returned)
__device__ __forceinline__ float reduce32(float* v, int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}
__device__ __forceinline__ float reduce8(float* v, int lane) {       // 8 values -> lane (l & 7) holds element l & 7
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

template <bool ONE>      // ONE: a single sequence tile (B <= 8): per-tile state lives in registers, not local memory
__global__ void __launch_bounds__(NT, 1) lstm_fwd_kernel(const LstmParams p) {
    extern __shared__ float sm[];
    const int H = p.H, Bp = p.Bp;
    float4* Wt = reinterpret_cast<float4*>(sm);             // [U][H] float4 = the 4 gates of (unit w, input k)
    float4* hs4 = reinterpret_cast<float4*>(sm + (size_t)U * H * 4);     // [Bp/4][H] float4 = 4 sequences of input k
    const int dir = blockIdx.x / p.G, slice = blockIdx.x % p.G, u0 = slice * U;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* Whh = p.whh[dir];
    for (int i = tid; i < U * H; i += NT) {
        const int uu = i / H, k = i % H;
        Wt[i] = make_float4(Whh[(size_t)(0 * H + u0 + uu) * H + k], Whh[(size_t)(1 * H + u0 + uu) * H + k],
                            Whh[(size_t)(2 * H + u0 + uu) * H + k], Whh[(size_t)(3 * H + u0 + uu) * H + k]);
    }
    const int n4 = Bp / 4 * H;
    for (int i = tid; i < n4; i += NT) hs4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int tmax = 0;
    for (int b = 0; b < p.B; ++b) tmax = max(tmax, min(p.lens[b], p.Tp));
    __syncthreads();

    const int n_tiles = ONE ? 1 : Bp / BT;
    float c_state[8];                                       // supports Bp <= 64
#pragma unroll
    for (int i = 0; i < 8; ++i) c_state[i] = 0.0f;
    const int unit = u0 + w;
    float* xbuf = p.xchg + (size_t)dir * 2 * Bp * H;        // [2 parity][Bp/4][H][4]  (same layout as hs4)
    struct Kept { float gi, gf, gg, go, c, h; long long r, o; bool valid; };
    Kept keep = {};
    auto store_kept = [&](const Kept& k) {
        if (!k.valid) return;
        float* gp = p.gates + k.r * 8 * H + (size_t)dir * 4 * H + unit;
        gp[0] = k.gi; gp[H] = k.gf; gp[2 * H] = k.gg; gp[3 * H] = k.go;
        p.cstate[k.r * 2 * H + (size_t)dir * H + unit] = k.c;
        p.out[k.o * 2 * H + (size_t)dir * H + unit] = k.h;
    };

    // lengths are loop invariant; with a single sequence tile (B <= 8) the input projection of step s+1 is fetched
    // before the barrier of step s, so its DRAM latency never sits on the step's critical path
    const int my_len0 = (lane < 8 && (lane & 7) < p.B) ? min(p.lens[lane & 7], p.Tp) : 0;
    auto fetch_xp = [&](int s, int bb, int len, float* xp) {
        xp[0] = xp[1] = xp[2] = xp[3] = 0.f;
        if (lane < 8 && bb < p.B && s < len) {
            const int t = dir ? len - 1 - s : s;
            const long long r = (long long)bb * p.pitch + t;
#pragma unroll
            for (int g = 0; g < 4; ++g) xp[g] = __ldg(p.xproj + r * 8 * H + (size_t)dir * 4 * H + g * H + unit);
        }
    };
    float xp_next[4];
    if (ONE) fetch_xp(0, lane & 7, my_len0, xp_next);

    for (int s = 0; s < tmax; ++s) {
#pragma unroll 1
        for (int bt = 0; bt < n_tiles; ++bt) {
            const int bb = bt * BT + (lane & 7);
            int len = my_len0, t = 0;
            if (!ONE) len = (lane < 8 && bb < p.B) ? min(p.lens[bb], p.Tp) : 0;
            if (lane < 8 && bb < p.B) t = dir ? len - 1 - s : s;
            const bool active = lane < 8 && bb < p.B && s < len;
            const long long r = (long long)bb * p.pitch + t;
            float xp[4];
            if (ONE) {
#pragma unroll
                for (int g = 0; g < 4; ++g) xp[g] = xp_next[g];
                fetch_xp(s + 1, bb, len, xp_next);           // consumed after the next barrier
            } else {
                fetch_xp(s, bb, len, xp);
            }
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
            const float4* ha = hs4 + (size_t)(2 * bt) * H;      // sequences bt*8 .. +3
            const float4* hb = ha + H;                           // sequences bt*8+4 .. +7
#pragma unroll 2
            for (int k = lane; k < ((p.probe & 1) ? 0 : H); k += 32) {
                const float4 w4 = Wt[w * H + k];
                const float4 h0 = ha[k], h1 = hb[k];
                const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    acc[0 * 8 + b] = fmaf(w4.x, hv[b], acc[0 * 8 + b]);
                    acc[1 * 8 + b] = fmaf(w4.y, hv[b], acc[1 * 8 + b]);
                    acc[2 * 8 + b] = fmaf(w4.z, hv[b], acc[2 * 8 + b]);
                    acc[3 * 8 + b] = fmaf(w4.w, hv[b], acc[3 * 8 + b]);
                }
            }
            const float mine = reduce32(acc, lane);         // lane v: gate v/8, sequence v%8
            const float a_i = __shfl_sync(0xffffffffu, mine, (lane & 7));
            const float a_f = __shfl_sync(0xffffffffu, mine, (lane & 7) + 8);
            const float a_g = __shfl_sync(0xffffffffu, mine, (lane & 7) + 16);
            const float a_o = __shfl_sync(0xffffffffu, mine, (lane & 7) + 24);
            if (lane < 8 && bb < p.Bp) {
                float h = 0.0f;
                if (active) {
                    const float gi = sigmoidf_(a_i + xp[0]), gf = sigmoidf_(a_f + xp[1]);
                    const float gg = tanhf_(a_g + xp[2]), go = sigmoidf_(a_o + xp[3]);
                    const float c = gf * c_state[bt] + gi * gg;
                    h = go * tanhf_(c);
                    c_state[bt] = c;
                    // publish h; the copies kept for the backward pass (6 stores to HBM-homed lines) are deferred past the
                    // barrier so that the release fence only has this one store to wait for
                    xbuf[(((size_t)(s & 1) * (Bp / 4) + (bb >> 2)) * H + unit) * 4 + (bb & 3)] = h;
                    if (bt == n_tiles - 1) {
                        keep = Kept{gi, gf, gg, go, c, h, r, ((long long)bb * p.Tp + t), true};
                    } else {
                        store_kept(Kept{gi, gf, gg, go, c, h, r, ((long long)bb * p.Tp + t), true});
                    }
                } else {
                    xbuf[(((size_t)(s & 1) * (Bp / 4) + (bb >> 2)) * H + unit) * 4 + (bb & 3)] = h;
                }
            }
        }
        if (s + 1 == tmax) { store_kept(keep); break; }
        if (!(p.probe & 4)) dir_barrier(p.counters + dir, (unsigned int)p.G * (s + 1));
        else __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(xbuf + (size_t)(s & 1) * Bp * H);
        if (!(p.probe & 2)) copy_f4_batched<5>(hs4, src, n4, tid, [](int i) { return i; });
        store_kept(keep);
        keep.valid = false;
        __syncthreads();
    }
}

template <bool ONE>
__global__ void __launch_bounds__(NT, 1) lstm_bwd_kernel(const LstmParams p) {
    extern __shared__ float sm[];
    const int H = p.H, Bp = p.Bp, H4 = 4 * p.H;
    float* Wb = sm;                                          // [4H][U]   Wb[rho][uu] = Whh[rho][u0 + uu]
    float* dgs = sm + (size_t)U * H4;                        // [4H][BT]  gate gradients of one sequence tile
    float* red = dgs + (size_t)BT * H4;                      // [8 warps][U * BT] cross-warp partial sums
    const int dir = blockIdx.x / p.G, slice = blockIdx.x % p.G, u0 = slice * U;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* Whh = p.whh[dir];
    for (int i = tid; i < U * H4; i += NT) {
        const int rho = i / U, uu = i % U;
        Wb[i] = Whh[(size_t)rho * H + u0 + uu];
    }
    int tmax = 0;
    for (int b = 0; b < p.B; ++b) tmax = max(tmax, min(p.lens[b], p.Tp));
    __syncthreads();
    const int n_tiles = ONE ? 1 : Bp / BT;
    float dh_rec[8], dc_next[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dh_rec[i] = 0.0f; dc_next[i] = 0.0f; }
    const int unit = u0 + w;
    float* xbuf = p.xchg + (size_t)dir * 2 * Bp * H4;        // [2 parity][4H][Bp]
    // phase-2 thread tile: 2 units x 4 sequences, every 4th gate row
    const int ug = lane >> 3, bg = (lane >> 2) & 1, rq = lane & 3;

    // saved values of (unit, sequence lane&7) for one step; prefetched one step ahead when the batch is a single tile
    struct Saved { float gi, gf, gg, go, c, c_prev, dout; };
    const int my_len0 = ((lane & 7) < p.B) ? min(p.lens[lane & 7], p.Tp) : 0;      // lengths are loop invariant
    auto load_saved = [&](int s, int bb, Saved& v) -> bool {
        int len = my_len0;
        if (!ONE) len = (bb < p.B) ? min(p.lens[bb], p.Tp) : 0;
        const bool active = bb < p.B && s >= 0 && s < len;
        if (active) {
            const int t = dir ? len - 1 - s : s;
            const long long r = (long long)bb * p.pitch + t;
            const float* gp = p.gates + r * 8 * H + (size_t)dir * 4 * H + unit;
            v.gi = gp[0]; v.gf = gp[H]; v.gg = gp[2 * H]; v.go = gp[3 * H];
            v.c = p.cstate[r * 2 * H + (size_t)dir * H + unit];
            const long long rp = dir ? r + 1 : r - 1;
            v.c_prev = (s > 0) ? p.cstate[rp * 2 * H + (size_t)dir * H + unit] : 0.0f;
            v.dout = p.dout[((long long)bb * p.Tp + t) * 2 * H + (size_t)dir * H + unit];
        }
        return active;
    };
    // the copy of the gate gradients kept for the weight-gradient contractions is stored after the barrier (see forward)
    struct KeptG { float d_i, d_f, d_g, d_o; long long r; bool valid; };
    KeptG keep = {};
    auto store_kept = [&](const KeptG& k) {
        if (!k.valid) return;
        float* dg = p.dgates + k.r * 8 * H + (size_t)dir * 4 * H + unit;
        dg[0] = k.d_i; dg[H] = k.d_f; dg[2 * H] = k.d_g; dg[3 * H] = k.d_o;
    };
    Saved pre = {};
    bool pre_active = false;
    if (n_tiles == 1 && lane < 8) pre_active = load_saved(tmax - 1, lane & 7, pre);

    for (int s = tmax - 1; s >= 0; --s) {
        // phase 1: gate gradients of this CTA's units
#pragma unroll 1
        for (int bt = 0; bt < n_tiles; ++bt) {
            const int bb = bt * BT + (lane & 7);
            if (lane < 8 && bb < Bp) {
                int len = my_len0;
                if (!ONE) len = (bb < p.B) ? min(p.lens[bb], p.Tp) : 0;
                Saved v = pre;
                const bool active = (n_tiles == 1) ? pre_active : load_saved(s, bb, v);
                float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f;
                long long r = 0;
                if (active) {
                    const int t = dir ? len - 1 - s : s;
                    r = (long long)bb * p.pitch + t;
                    const float gi = v.gi, gf = v.gf, gg = v.gg, go = v.go, c = v.c, c_prev = v.c_prev;
                    const float dh = v.dout + dh_rec[bt];
                    const float tc = tanhf_(c);
                    const float dc = dh * go * (1.0f - tc * tc) + dc_next[bt];
                    d_o = dh * tc * go * (1.0f - go);
                    d_i = dc * gg * gi * (1.0f - gi);
                    d_g = dc * gi * (1.0f - gg * gg);
                    d_f = dc * c_prev * gf * (1.0f - gf);
                    dc_next[bt] = dc * gf;
                } else {
                    dc_next[bt] = 0.0f;
                }
                // publish first (the other CTAs wait for it), then the copy kept for the weight-gradient contractions
                float* x = xbuf + ((size_t)(s & 1) * H4 + unit) * Bp + bb;          // lanes = consecutive sequences: 32 B segments
                x[0] = d_i; x[(size_t)H * Bp] = d_f; x[(size_t)2 * H * Bp] = d_g; x[(size_t)3 * H * Bp] = d_o;
                if (active) {
                    if (bt == n_tiles - 1) keep = KeptG{d_i, d_f, d_g, d_o, r, true};
                    else store_kept(KeptG{d_i, d_f, d_g, d_o, r, true});
                }
            }
        }
        if (s == 0) { store_kept(keep); break; }
        if (n_tiles == 1 && lane < 8) pre_active = load_saved(s - 1, lane & 7, pre);     // in flight across the barrier
        dir_barrier(p.counters + dir, (unsigned int)p.G * (tmax - s));
        store_kept(keep);
        keep.valid = false;
        // phase 2: dh_rec[b][unit] = sum_rho Whh[rho][unit] * dG[b][rho].  Warp w takes the gate rows rho = 4*(w + 8*i) + rq;
        // a lane accumulates 2 units x 4 sequences (one LDS.64 of weights + one LDS.128 of gate gradients per 8 FMAs, both
        // conflict-free), the 4 rq lanes are summed with two shuffles and the 8 warps through shared memory.
#pragma unroll 1
        for (int bt = 0; bt < n_tiles; ++bt) {
            const float4* src = reinterpret_cast<const float4*>(xbuf + (size_t)(s & 1) * H4 * Bp + (size_t)bt * BT);
            __syncthreads();
            const int q = Bp / 4;
            copy_f4_batched<6>(reinterpret_cast<float4*>(dgs), src, H4 * 2, tid,
                               [q](int i) { return (i >> 1) * q + (i & 1); });      // (rho, half): 4 sequences = one float4
            __syncthreads();
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 6
            for (int rho = 4 * w + rq; rho < H4; rho += 32) {
                const float2 wv = *reinterpret_cast<const float2*>(Wb + rho * U + 2 * ug);
                const float4 g = *reinterpret_cast<const float4*>(dgs + rho * BT + 4 * bg);
                a0[0] = fmaf(wv.x, g.x, a0[0]); a0[1] = fmaf(wv.x, g.y, a0[1]);
                a0[2] = fmaf(wv.x, g.z, a0[2]); a0[3] = fmaf(wv.x, g.w, a0[3]);
                a1[0] = fmaf(wv.y, g.x, a1[0]); a1[1] = fmaf(wv.y, g.y, a1[1]);
                a1[2] = fmaf(wv.y, g.z, a1[2]); a1[3] = fmaf(wv.y, g.w, a1[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], 1);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], 1);
                a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], 2);
                a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], 2);
            }
            if (rq == 0) {
                float* o = red + w * (U * BT) + (2 * ug) * BT + 4 * bg;
                *reinterpret_cast<float4*>(o) = make_float4(a0[0], a0[1], a0[2], a0[3]);
                *reinterpret_cast<float4*>(o + BT) = make_float4(a1[0], a1[1], a1[2], a1[3]);
            }
            __syncthreads();
            if (lane < 8) {
                float v = 0.0f;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) v += red[ww * (U * BT) + w * BT + lane];
                dh_rec[bt] = v;
            }
        }
    }
}

}  // namespace

size_t lstm_workspace_bytes(int B, int H) {
    const int Bp = (int)round_up(B, BT);
    return 256 + sizeof(float) * 2 * 2 * (size_t)Bp * 4 * H;      // counters + the larger (backward) exchange buffers
}

static int lstm_common(LstmParams& p, const int* lens, int B, int Tp, int H, void* workspace) {
    RADMMM_REQUIRE(H % U == 0 && H >= U, "lstm: hidden size %d must be a multiple of %d", H, U);
    RADMMM_REQUIRE(B >= 1 && B <= 64, "lstm: batch %d must be in [1, 64] (chunk larger batches)", B);
    RADMMM_REQUIRE(workspace != nullptr, "lstm: workspace missing");
    p.lens = lens; p.B = B; p.Bp = (int)round_up(B, BT); p.Tp = Tp; p.H = H; p.pitch = Tp + 16; p.G = H / U;
    p.counters = reinterpret_cast<unsigned int*>(workspace);
    p.xchg = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    RADMMM_REQUIRE(2 * p.G <= sms, "lstm: needs %d co-resident CTAs but the device has %d SMs", 2 * p.G, sms);
    { const char* e = getenv("RADMMM_B200_LSTM_PROBE"); p.probe = e ? atoi(e) : 0; }
    return RADMMM_OK;
}

// lstm_cluster.cu
bool lstm_cluster_supported(int B, int H);
int lstm_cluster_forward(const float* xproj, const float* whh_f, const float* whh_r, const int* lens, int B, int Tp, int H,
                         float* out, float* gates, float* cstate, cudaStream_t st);
int lstm_cluster_backward(const float* dout, const float* gates, const float* cstate, const float* whh_f, const float* whh_r,
                          const int* lens, int B, int Tp, int H, float* dgates, cudaStream_t st);
// MODE_BF16 -> the cluster-resident tensor-core recurrence; the fp32-grade modes -> the fp32 cooperative kernel below
// (RADMMM_B200_LSTM_CLUSTER=0 forces the latter: A/B measurements)
static bool use_cluster(int mode, int B, int H) {
    static const bool on = []() { const char* e = getenv("RADMMM_B200_LSTM_CLUSTER"); return !(e && e[0] == '0'); }();
    return on && mode == MODE_BF16 && lstm_cluster_supported(B, H);
}

int lstm_forward(int mode, const float* xproj, const float* whh_f, const float* whh_r, const int* lens, int B, int Tp, int H,
                 float* out, float* gates, float* cstate, void* workspace, cudaStream_t st) {
    if (use_cluster(mode, B, H)) return lstm_cluster_forward(xproj, whh_f, whh_r, lens, B, Tp, H, out, gates, cstate, st);
    LstmParams p;
    memset(&p, 0, sizeof(p));
    RADMMM_TRY(lstm_common(p, lens, B, Tp, H, workspace));
    p.xproj = xproj; p.whh[0] = whh_f; p.whh[1] = whh_r; p.out = out; p.gates = gates; p.cstate = cstate;
    const size_t smem = sizeof(float) * ((size_t)U * H * 4 + (size_t)p.Bp * H);
    RADMMM_REQUIRE(smem <= 220 * 1024, "lstm: shared memory %zu B exceeds the SM (H=%d, B=%d)", smem, H, B);
    static size_t smem_set_fwd_dev[64] = {};        // function attributes are per device
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    size_t& smem_set_fwd = smem_set_fwd_dev[dev_id & 63];
    if (smem > smem_set_fwd) {
        RADMMM_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RADMMM_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_fwd = smem;
    }
    RADMMM_CUDA(cudaMemsetAsync(p.counters, 0, 256, st));
    void* args[] = {&p};
    RADMMM_CUDA(cudaLaunchCooperativeKernel(p.Bp == BT ? (void*)lstm_fwd_kernel<true> : (void*)lstm_fwd_kernel<false>,
                                            dim3(2 * p.G), dim3(NT), args, smem, st));
    count_launch();
    return RADMMM_OK;
}

int lstm_backward(int mode, const float* dout, const float* gates, const float* cstate, const float* whh_f, const float* whh_r,
                  const int* lens, int B, int Tp, int H, float* dgates, void* workspace, cudaStream_t st) {
    if (use_cluster(mode, B, H)) return lstm_cluster_backward(dout, gates, cstate, whh_f, whh_r, lens, B, Tp, H, dgates, st);
    LstmParams p;
    memset(&p, 0, sizeof(p));
    RADMMM_TRY(lstm_common(p, lens, B, Tp, H, workspace));
    p.dout = dout; p.gates = const_cast<float*>(gates); p.cstate = const_cast<float*>(cstate);
    p.whh[0] = whh_f; p.whh[1] = whh_r; p.dgates = dgates;
    const size_t smem = sizeof(float) * ((size_t)U * 4 * H + (size_t)BT * 4 * H + 8 * U * BT);
    RADMMM_REQUIRE(smem <= 220 * 1024, "lstm: shared memory %zu B exceeds the SM (H=%d)", smem, H);
    static size_t smem_set_bwd_dev[64] = {};
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    size_t& smem_set_bwd = smem_set_bwd_dev[dev_id & 63];
    if (smem > smem_set_bwd) {
        RADMMM_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RADMMM_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_bwd = smem;
    }
    RADMMM_CUDA(cudaMemsetAsync(p.counters, 0, 256, st));
    void* args[] = {&p};
    RADMMM_CUDA(cudaLaunchCooperativeKernel(p.Bp == BT ? (void*)lstm_bwd_kernel<true> : (void*)lstm_bwd_kernel<false>,
                                            dim3(2 * p.G), dim3(NT), args, smem, st));
    count_launch();
    return RADMMM_OK;
}

}  // namespace radmmm
