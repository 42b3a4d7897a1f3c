// Contraction interface shared by the FFMA (fp32) and tcgen05 (bf16 / bf16x3) kernels, and the fused epilogues.
//
// "Row GEMM":   D[r][n] = sum_seg sum_k A_seg[r + row_shift_seg][k] * W_seg[n][k]        r in [0,R), n in [0,N)
//     A_seg are act matrices over rows (activations / gradients, [R][K_seg]); rows outside [0,R) read as zero.
//     W_seg are prepared weights [N_pad][K_seg], K-major.  A k=5 dilated conv is 5 segments with row shifts
//     (j-2)*d; the WN start conv is 2 segments (z0 | context); the fused dgrad of a layer is 6 segments.
// "Weight-grad GEMM":  D_tap[m][n] = sum_r dY[r][m] * X[r + shift_tap][n]                 (K = rows; both operands
//     are read from the same row matrices -- MN-major UMMA operands on the tensor-core path, no transposed copies)
//
// The epilogue (one of EPI_*) consumes the fp32 accumulators while they are still on chip.
#pragma once
#include "common.cuh"

namespace radmmm {

enum {
    EPI_START = 0,  // h0 = (acc + bias) masked                               (common.py:820)
    EPI_IN = 1,     // h = softplus(acc * ratio + bias) masked                 (partialconv1d.py:84-94, common.py:190,830)
    EPI_RS = 2,     // s_i = softplus(acc + bias) stored per layer               (common.py:831-832)
    EPI_END = 3,    // params = acc + bias -> channels-first fp32              (common.py:834)
    EPI_DOUT = 4,   // d_out = acc; dq_i = d_out * (1 - exp(-s_i))  for every layer i
    EPI_DH = 5,     // dacc = acc * softplus'(p) * ratio, masked
    EPI_DH0 = 6,    // dh0 = acc masked
    EPI_DZ0 = 7,    // dz0 -> channels-first fp32 (accumulate)
    EPI_DCTX = 8,   // dctx rows fp32
    EPI_WGRAD = 9,  // weight gradient tile -> fp32 (store or atomic add)
    EPI_F32 = 10,   // generic: rows fp32 = acc (+ bias)
};

constexpr int kMaxSeg = 6;
constexpr int kMaxLayers = 8;

struct GemmSeg {
    ActMat a;        // rows operand [R][K]   (weight-grad: dY [R][M])
    ActMat w;        // weights [N_pad][K]    (weight-grad: X [R][N])
    int K;           // contraction length of this segment (multiple of 64)
    int shift;       // row shift applied to A (row GEMM) / to X (weight-grad GEMM)
};

struct EpiParams {
    int kind;
    RowGeom geom;
    int N;                       // logical number of output columns
    int M;                       // weight-grad: logical number of output rows
    const float* bias;           // [N] or null
    int dilation;
    int first, last, accumulate, n_layers;
    ActMat out0, out1;
    float* f32_out;
    long long f32_ld;
    long long f32_tap_stride;    // weight-grad: elements between taps
    const float* padq;
    ActMat sig[kMaxLayers];      // the stored s_i (EPI_DOUT input)
    ActMat dq[kMaxLayers];
    ActMat h;
    float* cf_out;               // channels-first (B, cf_C, Tp) fp32
    int cf_C, cf_c0;
    int atomic;                  // weight-grad split-K: atomicAdd instead of store
    float* group_out[4];         // grouped weight-grad (GemmArgs::wgrad == 3): output of group g (else null: f32_out + tap stride)
    // Bias gradients fused into the gradient epilogues of the tensor-core path (null: not wanted / computed by colsum()):
    // colsum[l][n] += sum over valid rows of what the epilogue produces for layer l BEFORE the partial-conv ratio
    // (EPI_DOUT: dq_l -> res-skip bias l; EPI_DH: acc * softplus' -> dilated-conv bias; EPI_DH0: dh0 -> start bias).
    // The buffers must be zero when the kernel starts.
    float* colsum[kMaxLayers];
};

struct GemmArgs {
    int n_seg;
    GemmSeg seg[kMaxSeg];
    int R;          // rows (multiple of 128)
    int n_tiles_n;  // informational
    int wgrad;      // 0 row GEMM; 1 weight-grad GEMM, one output tile set per segment (= conv tap);
                    // 2 weight-grad GEMM, all segments (same dY, different X) accumulate into ONE output
                    // 3 GROUPED weight-grad GEMMs: segment g is an independent problem dY_g^T X_g -> epi.group_out[g] (same M, N,
                    //   R for all groups): one launch instead of n_seg small ones, full-K tiles, no split-K / atomics / memset
    int split_k;    // weight-grad only
    int zero_output;  // weight-grad: the launcher zeroes `f32_out` itself if (and only if) its split-K reduces with atomics,
                      // and stores plainly otherwise (epi.atomic is then decided by the launcher)
    EpiParams epi;
};

// zero the [taps][M][ld] fp32 weight-gradient output of `a` (columns [0, N))
inline int zero_wgrad_output(const GemmArgs& a, cudaStream_t st) {
    const int taps = a.wgrad == 2 ? 1 : a.n_seg;
    const long long ld = a.epi.f32_ld, ts = a.epi.f32_tap_stride;
    if (ld == a.epi.N && (taps == 1 || ts == (long long)a.epi.M * ld)) {       // one contiguous block
        RADMMM_CUDA(cudaMemsetAsync(a.epi.f32_out, 0, sizeof(float) * (size_t)taps * a.epi.M * ld, st));
    } else {
        for (int j = 0; j < taps; ++j)
            RADMMM_CUDA(cudaMemset2DAsync(a.epi.f32_out + j * ts, sizeof(float) * ld, 0, sizeof(float) * a.epi.N, a.epi.M, st));
    }
    return RADMMM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue for one row r and NV consecutive columns n0..n0+NV-1 (n0 % NV == 0, NV in {4, 8, 16, 32}).
//
// On the tcgen05 path a warp holds 32 consecutive rows (lane = row), so a per-thread 64-byte row store would touch 32
// different lines per instruction.  `Stager` gives each epilogue warp a small shared-memory tile: values are written
// there row-wise, then the warp moves 8 rows x 64 B per instruction (full sectors, 4x fewer line transactions); loads
// of saved activations go the same way in reverse.  The FFMA path passes a null stager (its thread tile is already
// 8 columns x 8 rows per thread with 16 lanes side by side).
// ---------------------------------------------------------------------------------------------------------
struct Stager {
    __nv_bfloat16* buf;   // [32 rows][kStageLd] per warp, or nullptr
    int lane;
    // TMA store of bf16 output tiles (tensor-core path, kinds whose only act-matrix output is `out0`): tensor maps of the
    // matrix `tma_base` (hi plane / lo plane of the split mode) with 32 x 32 boxes and the 64-byte swizzle; nullptr = off
    const void* tmap = nullptr;
    const void* tmap_lo = nullptr;
    const void* tma_base = nullptr;
};
constexpr int kStageLd = 40;   // 32 bf16 + 8 pad (80 B rows: conflict-light 16-byte accesses)

template <int MODE>
__device__ __forceinline__ void staged_store32(const Stager& st, const ActMat& m, int r, int n0, const float* v) {
    // r = row of THIS thread (row_base + lane); all 32 lanes call together
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(m.ptr);
    const int row_base = r - st.lane;
    if (st.tmap != nullptr && m.ptr == st.tma_base) {
        // One TMA store (cp.async.bulk.tensor, UTMASTG) per 32 x 32 tile instead of four 16-byte st.global per lane: every
        // lane writes its row (64 bytes) into the staging tile in the 64-byte-swizzle order the tensor map expects -- 16-byte
        // chunk c of row r sits at chunk c ^ ((r >> 1) & 3), which also makes the 32 concurrent 16-byte shared stores
        // conflict-free -- and lane 0 hands the tile to the copy engine.  The tile is reused by the next chunk, so the
        // previous store must have finished READING it first (wait_group.read).
        uint8_t* tile = reinterpret_cast<uint8_t*>(st.buf);
        const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(tile);
        const int sw = (st.lane >> 1) & 3;
#pragma unroll
        for (int plane = 0; plane < (MODE == MODE_BF16X3 ? 2 : 1); ++plane) {
            __nv_bfloat162 h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = v[2 * j], b = v[2 * j + 1];
                if (plane == 1) {
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
                    a -= __bfloat162float(hi.x); b -= __bfloat162float(hi.y);
                }
                h[j] = __floats2bfloat162_rn(a, b);
            }
            if (st.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(tile + st.lane * 64 + ((j ^ sw) << 4)) = reinterpret_cast<uint4*>(h)[j];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (st.lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(plane == 0 ? st.tmap : st.tmap_lo), "r"(n0), "r"(row_base), "r"(tile_addr) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        return;
    }
#pragma unroll
    for (int plane = 0; plane < (MODE == MODE_BF16X3 ? 2 : 1); ++plane) {
        __nv_bfloat162 h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float a = v[2 * j], b = v[2 * j + 1];
            if (plane == 1) {
                const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
                a -= __bfloat162float(hi.x); b -= __bfloat162float(hi.y);
            }
            h[j] = __floats2bfloat162_rn(a, b);
        }
        uint4* mine = reinterpret_cast<uint4*>(st.buf + st.lane * kStageLd);
#pragma unroll
        for (int j = 0; j < 4; ++j) mine[j] = reinterpret_cast<uint4*>(h)[j];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (st.lane >> 2), seg = st.lane & 3;
            const uint4 val = *reinterpret_cast<const uint4*>(st.buf + rr * kStageLd + seg * 8);
            *reinterpret_cast<uint4*>(base + (long long)plane * m.plane_stride + (long long)(row_base + rr) * m.ld + n0 + seg * 8) = val;
        }
        __syncwarp();
    }
}

// Saved-activation tile of one epilogue warp, 32 rows x 32 columns: coalesced global loads into registers (`raw`, issued
// early so that their latency overlaps the main loop / other loads), later turned into this thread's row via the stage.
template <int MODE>
struct RawTile { uint4 q[(MODE == MODE_BF16X3 ? 2 : 1) * 4]; };

template <int MODE>
__device__ __forceinline__ void staged_fetch32(const Stager& st, const ActMat& m, int r, int n0, RawTile<MODE>& raw) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.ptr);
    const int row_base = r - st.lane;
#pragma unroll
    for (int plane = 0; plane < (MODE == MODE_BF16X3 ? 2 : 1); ++plane) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (st.lane >> 2), seg = st.lane & 3;
            raw.q[plane * 4 + i] = *reinterpret_cast<const uint4*>(base + (long long)plane * m.plane_stride +
                                                                    (long long)(row_base + rr) * m.ld + n0 + seg * 8);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void staged_unpack32(const Stager& st, const RawTile<MODE>& raw, float* v) {
    if (st.tmap != nullptr) {      // the staging tile may still be the source of the previous chunk's TMA store
        if (st.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
#pragma unroll
    for (int plane = 0; plane < (MODE == MODE_BF16X3 ? 2 : 1); ++plane) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (st.lane >> 2), seg = st.lane & 3;
            *reinterpret_cast<uint4*>(st.buf + rr * kStageLd + seg * 8) = raw.q[plane * 4 + i];
        }
        __syncwarp();
        const uint4* mine = reinterpret_cast<const uint4*>(st.buf + st.lane * kStageLd);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 q = mine[j];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(h2[k]);
                if (plane == 0) { v[8 * j + 2 * k] = f.x; v[8 * j + 2 * k + 1] = f.y; }
                else { v[8 * j + 2 * k] += f.x; v[8 * j + 2 * k + 1] += f.y; }
            }
        }
        __syncwarp();
    }
}

template <int MODE>
__device__ __forceinline__ void staged_load32(const Stager& st, const ActMat& m, int r, int n0, float* v) {
    RawTile<MODE> raw;
    staged_fetch32<MODE>(st, m, r, n0, raw);
    staged_unpack32<MODE>(st, raw, v);
}

// fp32 rows [R][ld]: this thread holds 32 consecutive floats of row r.  Through the stage (80-byte rows = 16 floats + pad)
// in two halves, so that a warp instruction writes 8 rows x 64 contiguous bytes instead of 32 rows x 4 bytes.
__device__ __forceinline__ void staged_store32_f32(const Stager& st, float* out, long long ld, int r, int n0, const float* v) {
    float* sbuf = reinterpret_cast<float*>(st.buf);
    constexpr int kLdF = kStageLd / 2;                      // 20 floats per staged row
    const int row_base = r - st.lane;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float4* mine = reinterpret_cast<float4*>(sbuf + st.lane * kLdF);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            mine[j] = make_float4(v[half * 16 + 4 * j], v[half * 16 + 4 * j + 1], v[half * 16 + 4 * j + 2], v[half * 16 + 4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (st.lane >> 2), seg = st.lane & 3;
            const float4 val = *reinterpret_cast<const float4*>(sbuf + rr * kLdF + seg * 4);
            *reinterpret_cast<float4*>(out + (long long)(row_base + rr) * ld + n0 + half * 16 + seg * 4) = val;
        }
        __syncwarp();
    }
}

// Column sums of a warp's 32 x 32 tile (lane = row, v[32] = this row's values): halving butterfly, 31 shuffles; lane c
// ends up with the sum of column c.  Destroys v.
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
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
__device__ __forceinline__ void colsum_add(float* dst, int n0, int N, int lane, float* v) {
    const float sum = warp_colsum32(v, lane);
    if (n0 + lane < N) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + n0 + lane), "f"(sum) : "memory");
}

template <int MODE, int NV>
__device__ __forceinline__ void store_row_vec(const Stager& st, const ActMat& m, int r, int n0, const float* v) {
    if (m.ptr == nullptr) return;
    if constexpr (MODE != MODE_F32 && NV == 32) {
        if (st.buf != nullptr) { staged_store32<MODE>(st, m, r, n0, v); return; }
    }
    long long idx = (long long)r * m.ld + n0;
    if constexpr (MODE == MODE_F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(m.ptr) + idx);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(m.ptr);
        if constexpr (NV >= 8) {
#pragma unroll
            for (int i = 0; i < NV / 8; ++i) {
                __nv_bfloat162 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = v[8 * i + 2 * j], b = v[8 * i + 2 * j + 1];
                    h[j] = __floats2bfloat162_rn(a, b);
                    if constexpr (MODE == MODE_BF16X3)
                        l[j] = __floats2bfloat162_rn(a - __bfloat162float(h[j].x), b - __bfloat162float(h[j].y));
                }
                *reinterpret_cast<uint4*>(base + idx + 8 * i) = *reinterpret_cast<uint4*>(h);
                if constexpr (MODE == MODE_BF16X3)
                    *reinterpret_cast<uint4*>(base + idx + m.plane_stride + 8 * i) = *reinterpret_cast<uint4*>(l);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) act_store<MODE>(m, idx + i, v[i]);
        }
    }
}

template <int MODE, int NV>
__device__ __forceinline__ void load_row_vec(const Stager& st, const ActMat& m, int r, int n0, float* v) {
    if constexpr (MODE != MODE_F32 && NV == 32) {
        if (st.buf != nullptr) { staged_load32<MODE>(st, m, r, n0, v); return; }
    }
    long long idx = (long long)r * m.ld + n0;
    if constexpr (MODE == MODE_F32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(m.ptr) + idx);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            float4 t = p[i];
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.ptr);
#pragma unroll
        for (int i = 0; i < NV / 8; ++i) {
            const uint4 hv = *reinterpret_cast<const uint4*>(base + idx + 8 * i);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y;
            }
            if constexpr (MODE == MODE_BF16X3) {
                const uint4 lv = *reinterpret_cast<const uint4*>(base + idx + m.plane_stride + 8 * i);
                const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&lv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(l2[j]);
                    v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y;
                }
            }
        }
    }
}

template <int NV>
__device__ __forceinline__ void load_vec_f32(const float* __restrict__ src, float* v) {   // 16-byte aligned source
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

template <int MODE, int KIND, int NV>
__device__ __forceinline__ void epi_apply(const EpiParams& p, const Stager& st, int r, int n0, float* acc) {
    constexpr bool FAST = MODE != MODE_F32;
    int b, t, len;
    row_decode(p.geom, r, b, t, len);
    const bool valid = t < len;
    float out[NV];
    if constexpr (KIND == EPI_START) {
        float bias[NV];
        load_vec_f32<NV>(p.bias + n0, bias);
#pragma unroll
        for (int i = 0; i < NV; ++i) out[i] = valid ? acc[i] + bias[i] : 0.0f;
        store_row_vec<MODE, NV>(st, p.out0, r, n0, out);
    } else if constexpr (KIND == EPI_IN) {
        const float ratio = valid ? pconv_ratio(t, len, p.dilation) : 0.0f;
        float bias[NV];
        load_vec_f32<NV>(p.bias + n0, bias);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float y = softplus_f<FAST>(fmaf(acc[i], ratio, bias[i]));
            out[i] = valid ? y : 0.0f;
        }
        store_row_vec<MODE, NV>(st, p.out0, r, n0, out);
    } else if constexpr (KIND == EPI_RS) {
        // s_i = softplus(res_skip_i(h)); frames beyond the length carry the reference's constant softplus(padq).
        // The skip sum is never materialised: the `end` GEMM contracts over the L stored s_i as K-segments.
        float s[NV];
        load_vec_f32<NV>((valid ? p.bias : p.padq) + n0, s);
#pragma unroll
        for (int i = 0; i < NV; ++i) out[i] = softplus_f<FAST>((valid ? acc[i] : 0.0f) + s[i]);
        store_row_vec<MODE, NV>(st, p.out0, r, n0, out);
    } else if constexpr (KIND == EPI_END || KIND == EPI_DZ0) {
        if (b < p.geom.B && t < p.geom.Tp && n0 < p.N) {
            float* dst0 = p.cf_out + ((long long)(b * p.cf_C + p.cf_c0 + n0)) * p.geom.Tp + t;
            const long long cs = p.geom.Tp;                 // channel stride; lanes = consecutive t: coalesced
            if (p.accumulate) {                              // all loads first (they are independent L2 round trips)
#pragma unroll
                for (int i = 0; i < NV; ++i) out[i] = (n0 + i < p.N) ? dst0[i * cs] : 0.0f;
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) out[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (n0 + i < p.N) dst0[i * cs] = out[i] + acc[i] + (KIND == EPI_END ? __ldg(p.bias + n0 + i) : 0.0f);
        }
    } else if constexpr (KIND == EPI_DOUT) {
        if constexpr (MODE != MODE_F32 && NV == 32) {
            if (st.buf != nullptr) {
                // tensor-core path: the epilogue is pure memory traffic (L tiles in, L tiles out); issue the loads of a
                // group of layers back to back so that their latencies overlap instead of adding up
                constexpr int G = (MODE == MODE_BF16X3) ? 2 : 4;
                for (int l0 = 0; l0 < p.n_layers; l0 += G) {
                    RawTile<MODE> raw[G];
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        if (l0 + j < p.n_layers) staged_fetch32<MODE>(st, p.sig[l0 + j], r, n0, raw[j]);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (l0 + j < p.n_layers) {
                            float sv[NV];
                            staged_unpack32<MODE>(st, raw[j], sv);
#pragma unroll
                            for (int i = 0; i < NV; ++i) {
                                const float y = acc[i] * sigmoid_from_softplus<FAST>(sv[i]);
                                out[i] = valid ? y : 0.0f;
                            }
                            store_row_vec<MODE, NV>(st, p.dq[l0 + j], r, n0, out);
                            if (p.colsum[l0 + j] != nullptr) colsum_add(p.colsum[l0 + j], n0, p.N, st.lane, out);
                        }
                    }
                }
                return;
            }
        }
        for (int l = 0; l < p.n_layers; ++l) {
            float sv[NV];
            load_row_vec<MODE, NV>(st, p.sig[l], r, n0, sv);          // s_l = softplus(q_l); d softplus = 1 - exp(-s)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float y = acc[i] * sigmoid_from_softplus<FAST>(sv[i]);
                out[i] = valid ? y : 0.0f;
            }
            store_row_vec<MODE, NV>(st, p.dq[l], r, n0, out);
        }
    } else if constexpr (KIND == EPI_DH) {
        float hv[NV];
        load_row_vec<MODE, NV>(st, p.h, r, n0, hv);             // all lanes take part (staged, warp-cooperative)
        const float ratio = valid ? pconv_ratio(t, len, p.dilation) : 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float y = acc[i] * sigmoid_from_softplus<FAST>(hv[i]) * ratio;
            out[i] = valid ? y : 0.0f;
        }
        store_row_vec<MODE, NV>(st, p.out0, r, n0, out);
    } else if constexpr (KIND == EPI_DH0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) out[i] = valid ? acc[i] : 0.0f;
        store_row_vec<MODE, NV>(st, p.out0, r, n0, out);
        if constexpr (MODE != MODE_F32 && NV == 32) {
            if (st.buf != nullptr && p.colsum[0] != nullptr) colsum_add(p.colsum[0], n0, p.N, st.lane, out);
        }
    } else if constexpr (KIND == EPI_DCTX || KIND == EPI_F32) {
        if constexpr (NV == 32) {
            if (st.buf != nullptr && !p.accumulate && n0 + NV <= p.N && (p.f32_ld & 3) == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) out[i] = acc[i] + ((KIND == EPI_F32 && p.bias) ? __ldg(p.bias + n0 + i) : 0.0f);
                staged_store32_f32(st, p.f32_out, p.f32_ld, r, n0, out);
                return;
            }
        }
        float* o = p.f32_out + (long long)r * p.f32_ld + n0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (n0 + i < p.N) {
                float v = acc[i] + ((KIND == EPI_F32 && p.bias) ? __ldg(p.bias + n0 + i) : 0.0f);
                o[i] = p.accumulate ? o[i] + v : v;
            }
    }
}

// EPI_DH with the saved activation h already in registers (the tcgen05 kernel fetches it while the main loop runs)
template <int MODE>
__device__ __forceinline__ void epi_dh_with_h(const EpiParams& p, const Stager& st, int r, int n0, const float* acc,
                                              const float* hv) {
    int b, t, len;
    row_decode(p.geom, r, b, t, len);
    const bool valid = t < len;
    const float ratio = valid ? pconv_ratio(t, len, p.dilation) : 0.0f;
    float out[32], pre[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float y = acc[i] * sigmoid_from_softplus<true>(hv[i]);       // gradient at the conv's bias (before the ratio)
        pre[i] = valid ? y : 0.0f;
        out[i] = pre[i] * ratio;
    }
    store_row_vec<MODE, 32>(st, p.out0, r, n0, out);
    if (p.colsum[0] != nullptr) colsum_add(p.colsum[0], n0, p.N, st.lane, pre);
}

// weight-grad tile element (m, n0..n0+NV) of tap `tap`
template <int NV>
__device__ __forceinline__ void epi_wgrad(const EpiParams& p, int tap, int m, int n0, const float* acc) {
    if (m >= p.M) return;
    float* o = (p.group_out[tap & 3] != nullptr ? p.group_out[tap & 3] : p.f32_out + (long long)tap * p.f32_tap_stride) +
               (long long)m * p.f32_ld + n0;
    if (n0 + NV <= p.N && (p.f32_ld & 3) == 0) {       // whole vector in range: 16-byte vector reductions / stores
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            if (p.atomic)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(acc[4 * i]),
                             "f"(acc[4 * i + 1]), "f"(acc[4 * i + 2]), "f"(acc[4 * i + 3]) : "memory");
            else
                *reinterpret_cast<float4*>(o + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (n0 + i < p.N) {
            if (p.atomic) atomicAdd(o + i, acc[i]);
            else o[i] = acc[i];
        }
}

// launchers (gemm_ffma.cu / gemm_tc.cu)
int launch_gemm_ffma(const GemmArgs& args, cudaStream_t stream);
int launch_gemm_tc(const GemmArgs& args, int mode, cudaStream_t stream);
int launch_gemm(const GemmArgs& args, int mode, cudaStream_t stream);
void gemm_tc_set_trace(void* buf, int max_ctas, int max_launches);      // diagnostic in-kernel timeline (tools/gemm_probe.py)

}  // namespace radmmm
