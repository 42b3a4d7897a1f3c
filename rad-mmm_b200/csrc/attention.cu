// Fused soft attention (common.py:1259-1276) + text->frame expansion (tts_lightning_modules.py:670).
// The reference materialises the (B, 80, T1, T2) squared-difference tensor (0.3-2 GB); here one warp owns one
// query frame: keys stay resident in shared memory for the CTA's ROWS query frames, distances are accumulated in
// registers, log-softmax / prior / mask / softmax run on warp shuffles, and the attention row is reused from
// shared memory for context = txt_enc . attn^T.  HBM traffic is the algorithmic minimum:
// (Ca*(T1+T2) + [prior] T1*T2 + 2*T1*T2 + Dt*(T2+T1)) * 4 B per utterance.
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {


// Forward: FROWS query frames per CTA (one warp each for the softmax), then ALL threads of the CTA -- one per text-encoding
// channel -- do the context matmul for those frames.  Round 1 ran 8 frames per 256-thread CTA: every CTA re-read all keys
// (42 KB) and the whole text encoding (270 KB at Dt=520, T2=130) from L2 for 8 frames, the channel loop took 3 passes with
// the last one 3 % full, and each inner step needed 8 scalar shared loads: 283 us at B=8, T1=800.  Now 16 frames, the
// attention rows transposed in shared memory so a step is 4 x LDS.128 + 1 LDG per 16 FMAs, and a block of
// round_up(Dt, 32) threads so the channel loop is one pass.
constexpr int FROWS = 16;

__global__ void __launch_bounds__(1024) soft_attention_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ prior,
    const int* __restrict__ in_lens, float* __restrict__ attn, float* __restrict__ attn_logprob,
    const float* __restrict__ txt_enc, float* __restrict__ context, int Ca, int T1, int T2, int Dt, float temp) {
    extern __shared__ float sm[];
    float* ks = sm;                          // [Ca][T2]
    float* qs = ks + (size_t)Ca * T2;        // [FROWS][Ca]
    float* as = qs + FROWS * Ca;             // [FROWS][T2]  attention rows
    float* ast = sm + (((size_t)Ca * T2 + (size_t)FROWS * Ca + (size_t)FROWS * T2 + 3) & ~(size_t)3);   // [T2][FROWS] transposed, 16-byte aligned
    const int b = blockIdx.y, t1_0 = blockIdx.x * FROWS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthr = blockDim.x;
    const float* kb = k + (long long)b * Ca * T2;
    for (int i = tid; i < Ca * T2; i += nthr) ks[i] = kb[i];
    for (int i = tid; i < FROWS * Ca; i += nthr) {
        const int r = i / Ca, c = i % Ca, t1 = t1_0 + r;
        qs[i] = (t1 < T1) ? q[((long long)b * Ca + c) * T1 + t1] : 0.0f;
    }
    __syncthreads();
    const int t1 = t1_0 + wid;
    const int len = min(in_lens[b], T2);
    if (wid < FROWS) {
        float* arow = as + (size_t)wid * T2;
        if (t1 < T1) {
            // logits
            float mx = -INFINITY;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                float d = 0.0f;
                for (int c = 0; c < Ca; ++c) {
                    const float df = qs[wid * Ca + c] - ks[c * T2 + t2];
                    d = fmaf(df, df, d);
                }
                const float lg = -temp * d;
                arow[t2] = lg;
                mx = fmaxf(mx, lg);
            }
            const long long orow = ((long long)b * T1 + t1) * T2;
            if (prior != nullptr) {
                // log_softmax over ALL T2 keys (padded ones included, as the reference does), then + log(prior + 1e-8)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float se = 0.0f;
                for (int t2 = lane; t2 < T2; t2 += 32) se += expf(arow[t2] - mx);
                se = warp_sum(se);
                const float lse = mx + logf(se);
                for (int t2 = lane; t2 < T2; t2 += 32) arow[t2] = arow[t2] - lse + logf(prior[orow + t2] + 1e-8f);
            }
            // attn_logprob = pre-mask copy; softmax over the unmasked keys
            float m2 = -INFINITY;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                const float v = arow[t2];
                attn_logprob[orow + t2] = v;
                if (t2 < len) m2 = fmaxf(m2, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
            float s2 = 0.0f;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                const float e = (t2 < len) ? expf(arow[t2] - m2) : 0.0f;
                arow[t2] = e;
                s2 += e;
            }
            s2 = warp_sum(s2);
            const float inv = 1.0f / s2;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                const float a = arow[t2] * inv;
                arow[t2] = a;
                attn[orow + t2] = a;
            }
        } else {
            for (int t2 = lane; t2 < T2; t2 += 32) arow[t2] = 0.0f;       // frames past T1: zero weights, nothing stored
        }
    }
    if (txt_enc == nullptr) return;
    __syncthreads();
    for (int i = tid; i < FROWS * T2; i += nthr) {
        const int r = i / T2, t2 = i - r * T2;
        ast[t2 * FROWS + r] = as[i];
    }
    __syncthreads();
    // context[b, d, t1_0 .. t1_0+15] = sum_t2 txt_enc[b, d, t2] * attn[t1, t2]: one thread per channel d, 16 accumulators;
    // the thread walks its own txt_enc row (32-byte sectors stay in L1 across 8 steps), the weights are broadcast LDS.128
    const float* tb = txt_enc + (long long)b * Dt * T2;
    for (int d = tid; d < Dt; d += nthr) {
        float acc[FROWS];
#pragma unroll
        for (int r = 0; r < FROWS; ++r) acc[r] = 0.0f;
        const float* trow = tb + (long long)d * T2;
        for (int t2 = 0; t2 < len; ++t2) {
            const float tv = __ldg(trow + t2);
            const float4* w4 = reinterpret_cast<const float4*>(ast + t2 * FROWS);
#pragma unroll
            for (int j = 0; j < FROWS / 4; ++j) {
                const float4 w = w4[j];
                acc[4 * j + 0] = fmaf(tv, w.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(tv, w.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(tv, w.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(tv, w.w, acc[4 * j + 3]);
            }
        }
        float* o = context + ((long long)b * Dt + d) * T1 + t1_0;
        if (t1_0 + FROWS <= T1 && (T1 & 3) == 0) {
#pragma unroll
            for (int j = 0; j < FROWS / 4; ++j)
                reinterpret_cast<float4*>(o)[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < FROWS; ++r)
                if (t1_0 + r < T1) o[r] = acc[r];
        }
    }
}

// Backward of soft_attention_kernel, three launches:
//  1. soft_attention_bwd_kernel -- FROWS query frames per CTA.  The context matmul's contribution to d attn,
//     dA[t1,t2] += sum_d txt[d,t2] dctx[d,t1], is computed by the whole CTA from a shared-memory tile of dcontext
//     ([Dt][FROWS], 64-byte rows) with threads running over (t2, slice of d) so that the text encoding is read coalesced; then one
//     warp per frame recomputes its logits from q / k, applies the softmax (and log-softmax + prior) backward, writes the
//     gradient w.r.t. the squared distances to a workspace (B,T1,T2) and reduces dq for its own frame.
//  2. attn_tn_gemm_kernel<0>: dtxt[d,t2] = sum_t1 dctx[d,t1] attn[t1,t2] -- a tiled fp32 GEMM, no atomics.
//  3. attn_tn_gemm_kernel<1>: dk[c,t2] = -2 sum_t1 q[c,t1] g[t1,t2] + 2 k[c,t2] sum_t1 g[t1,t2] from the workspace.
// Round 1 did all of it in one kernel with 8 frames per CTA: 520 x T2 global atomics per CTA for dtxt (55 M per call at
// B=8, T1=800), shared + global atomics for dk and an uncoalesced 520-step dot product per (t1, t2): 1.85 ms.
__global__ void __launch_bounds__(FROWS * 32) soft_attention_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ prior,
    const int* __restrict__ in_lens, const float* __restrict__ attn, const float* __restrict__ dattn,
    const float* __restrict__ dlogprob, const float* __restrict__ txt_enc, const float* __restrict__ dcontext,
    float* __restrict__ dq, float* __restrict__ gbuf, int Ca, int T1, int T2, int Dt, float temp) {
    extern __shared__ float sm[];
    float* ks = sm;                          // [Ca][T2]
    float* qs = ks + (size_t)Ca * T2;        // [FROWS][Ca]
    float* gs = qs + FROWS * Ca;             // [FROWS][T2]  d attn -> d dist (per frame)
    float* as = gs + (size_t)FROWS * T2;     // [FROWS][T2]  attn rows / logits scratch
    float* dcs = sm + (((size_t)Ca * T2 + (size_t)FROWS * Ca + 2 * (size_t)FROWS * T2 + 3) & ~(size_t)3);   // [Dt][FROWS]
    const int b = blockIdx.y, t1_0 = blockIdx.x * FROWS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NTHR = FROWS * 32;
    const float* kb = k + (long long)b * Ca * T2;
    for (int i = tid; i < Ca * T2; i += NTHR) ks[i] = kb[i];
    for (int i = tid; i < FROWS * Ca; i += NTHR) {
        const int r = i / Ca, c = i % Ca, t1 = t1_0 + r;
        qs[i] = (t1 < T1) ? q[((long long)b * Ca + c) * T1 + t1] : 0.0f;
    }
    for (int i = tid; i < FROWS * T2; i += NTHR) gs[i] = 0.0f;
    const int len = min(in_lens[b], T2);
    const bool with_ctx = txt_enc != nullptr;
    if (with_ctx)
        for (int i = tid; i < Dt * FROWS; i += NTHR) {
            const int d = i / FROWS, r = i - d * FROWS;
            dcs[i] = (t1_0 + r < T1) ? dcontext[((long long)b * Dt + d) * T1 + t1_0 + r] : 0.0f;
        }
    __syncthreads();
    if (with_ctx) {
        // items = (slice of d, t2): consecutive threads -> consecutive t2 (coalesced txt reads), one d slice per group of T2 threads
        const int nsl = max(1, min(NTHR / max(T2, 1), Dt));
        const int per = (Dt + nsl - 1) / nsl;
        for (int item = tid; item < nsl * T2; item += NTHR) {
            const int sl = item / T2, t2 = item - sl * T2;
            if (t2 >= len) continue;
            const int d0 = sl * per, d1 = min(Dt, d0 + per);
            float acc[FROWS];
#pragma unroll
            for (int r = 0; r < FROWS; ++r) acc[r] = 0.0f;
            const float* tp = txt_enc + ((long long)b * Dt + d0) * T2 + t2;
            for (int d = d0; d < d1; ++d, tp += T2) {
                const float tv = __ldg(tp);
                const float4* w4 = reinterpret_cast<const float4*>(dcs + d * FROWS);
#pragma unroll
                for (int j = 0; j < FROWS / 4; ++j) {
                    const float4 w = w4[j];
                    acc[4 * j + 0] = fmaf(tv, w.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(tv, w.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(tv, w.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(tv, w.w, acc[4 * j + 3]);
                }
            }
#pragma unroll
            for (int r = 0; r < FROWS; ++r) atomicAdd(&gs[r * T2 + t2], acc[r]);      // nsl-way contention at most
        }
        __syncthreads();
    }
    const int t1 = t1_0 + wid;
    float* grow = gs + (size_t)wid * T2;
    float* arow = as + (size_t)wid * T2;
    if (t1 < T1) {
        const long long orow = ((long long)b * T1 + t1) * T2;
        // d attn = incoming + context-matmul part; softmax backward needs sum_t2 attn * d attn over the unmasked keys
        float dot = 0.0f;
        for (int t2 = lane; t2 < T2; t2 += 32) {
            const float da = (dattn ? dattn[orow + t2] : 0.0f) + grow[t2];
            const float a = attn[orow + t2];
            arow[t2] = a;
            grow[t2] = da;
            dot += (t2 < len) ? a * da : 0.0f;
        }
        dot = warp_sum(dot);
        // g_lp = dlogprob + softmax backward (masked keys get no softmax gradient)
        float gsum = 0.0f;
        for (int t2 = lane; t2 < T2; t2 += 32) {
            float g = (t2 < len) ? arow[t2] * (grow[t2] - dot) : 0.0f;
            if (dlogprob) g += dlogprob[orow + t2];
            grow[t2] = g;
            gsum += g;
        }
        if (prior != nullptr) {
            // log_softmax backward over ALL keys: g_logit = g_lp - softmax(logit) * sum(g_lp)
            gsum = warp_sum(gsum);
            float mx = -INFINITY;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                float d = 0.0f;
                for (int c = 0; c < Ca; ++c) { const float df = qs[wid * Ca + c] - ks[c * T2 + t2]; d = fmaf(df, df, d); }
                const float lg = -temp * d;
                arow[t2] = lg;
                mx = fmaxf(mx, lg);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float se = 0.0f;
            for (int t2 = lane; t2 < T2; t2 += 32) se += expf(arow[t2] - mx);
            se = warp_sum(se);
            for (int t2 = lane; t2 < T2; t2 += 32) grow[t2] -= expf(arow[t2] - mx) / se * gsum;
        }
        for (int t2 = lane; t2 < T2; t2 += 32) {
            const float g = grow[t2] * -temp;                                    // d dist
            grow[t2] = g;
            gbuf[orow + t2] = g;
        }
        __syncwarp();
        // dq[c] = sum_t2 2 (q - k) ddist
        for (int c = 0; c < Ca; ++c) {
            const float qc = qs[wid * Ca + c];
            float acc = 0.0f;
            for (int t2 = lane; t2 < T2; t2 += 32) acc = fmaf(2.0f * (qc - ks[c * T2 + t2]), grow[t2], acc);
            acc = warp_sum(acc);
            if (lane == 0) dq[((long long)b * Ca + c) * T1 + t1] = acc;
        }
    }
}

// out[b][m][t2] = sum_t1 A[b][m][t1] * Bm[b][t1][t2]   (A: M x T1 row-major, Bm: T1 x T2 row-major)
// MODE 0: plain (dtxt; columns t2 >= in_len are written as 0).  MODE 1 (dk): out = -2 * acc + 2 * kmat[b][m][t2] * colsum,
// colsum[t2] = sum_t1 Bm[t1][t2] accumulated by every thread for its own columns.
// 64 (m) x 128 (t2) tile per CTA, 256 threads of 4 x 8 outputs, K (= T1) in chunks of 32 through shared memory.
constexpr int AG_TM = 64, AG_TN = 128, AG_TK = 32;
template <int MODE>
__global__ void __launch_bounds__(256) attn_tn_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                           const float* __restrict__ kmat, const int* __restrict__ in_lens,
                                                           float* __restrict__ out, int M, int T1, int T2) {
    __shared__ __align__(16) float As[AG_TK][AG_TM + 4];     // [k][m]
    __shared__ __align__(16) float Bs[AG_TK][AG_TN];         // [k][t2]
    const int m0 = blockIdx.x * AG_TM, n0 = blockIdx.y * AG_TN, b = blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float* Ab = A + (long long)b * M * T1;
    const float* Bb = Bm + (long long)b * T1 * T2;
    float acc[4][8], cs[8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.0f;
    for (int k0 = 0; k0 < T1; k0 += AG_TK) {
        for (int i = tid; i < AG_TM * AG_TK; i += 256) {         // A tile: rows of 32 consecutive t1
            const int m = i >> 5, kk = i & 31;
            As[kk][m] = (m0 + m < M && k0 + kk < T1) ? Ab[(long long)(m0 + m) * T1 + k0 + kk] : 0.0f;
        }
        for (int i = tid; i < AG_TK * AG_TN; i += 256) {         // B tile: rows of 128 consecutive t2
            const int kk = i >> 7, n = i & 127;
            Bs[kk][n] = (k0 + kk < T1 && n0 + n < T2) ? Bb[(long long)(k0 + kk) * T2 + n0 + n] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < AG_TK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][8 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][8 * tx + 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) cs[j] += bv[j];
            }
        }
        __syncthreads();
    }
    const int len = min(in_lens[b], T2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + 8 * tx + j;
            if (n >= T2) continue;
            const long long o = ((long long)b * M + m) * T2 + n;
            if (MODE == 0) out[o] = (n < len) ? acc[i][j] : 0.0f;
            else out[o] = -2.0f * acc[i][j] + 2.0f * kmat[o] * cs[j];
        }
    }
}

}  // namespace

long long soft_attention_bwd_workspace_bytes(int B, int T1, int T2) { return (long long)B * T1 * T2 * (long long)sizeof(float); }

int soft_attention_bwd(const float* q, const float* k, const float* prior, const int* in_lens, const float* attn,
                       const float* dattn, const float* dlogprob, const float* txt_enc, const float* dcontext, float* dq,
                       float* dk, float* dtxt, int B, int Ca, int T1, int T2, int Dt, float temperature, void* workspace,
                       long long workspace_bytes, cudaStream_t st) {
    RADMMM_REQUIRE(B > 0 && Ca > 0 && T1 > 0 && T2 > 0, "soft_attention_bwd: bad sizes");
    RADMMM_REQUIRE((txt_enc == nullptr) == (dcontext == nullptr), "soft_attention_bwd: txt_enc and dcontext go together");
    RADMMM_REQUIRE(workspace != nullptr && workspace_bytes >= soft_attention_bwd_workspace_bytes(B, T1, T2),
                   "soft_attention_bwd: workspace of %lld bytes needed", soft_attention_bwd_workspace_bytes(B, T1, T2));
    const size_t head = ((size_t)Ca * T2 + (size_t)FROWS * Ca + 2 * (size_t)FROWS * T2 + 3) & ~(size_t)3;
    const size_t smem = sizeof(float) * (head + (txt_enc ? (size_t)Dt * FROWS : 0));
    RADMMM_REQUIRE(smem <= 220 * 1024, "soft_attention_bwd: T2=%d keys do not fit in shared memory", T2);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        RADMMM_CUDA(cudaFuncSetAttribute(soft_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set[dev & 63] = true;
    }
    float* gbuf = reinterpret_cast<float*>(workspace);
    dim3 grid(cdiv(T1, FROWS), B);
    soft_attention_bwd_kernel<<<grid, FROWS * 32, smem, st>>>(q, k, prior, in_lens, attn, dattn, dlogprob, txt_enc, dcontext, dq,
                                                              gbuf, Ca, T1, T2, Dt, temperature);
    RADMMM_LAUNCH_CHECK();
    dim3 gk(cdiv(Ca, AG_TM), cdiv(T2, AG_TN), B);
    attn_tn_gemm_kernel<1><<<gk, 256, 0, st>>>(q, gbuf, k, in_lens, dk, Ca, T1, T2);
    RADMMM_LAUNCH_CHECK();
    if (dtxt != nullptr) {
        dim3 gt(cdiv(Dt, AG_TM), cdiv(T2, AG_TN), B);
        attn_tn_gemm_kernel<0><<<gt, 256, 0, st>>>(dcontext, attn, nullptr, in_lens, dtxt, Dt, T1, T2);
        RADMMM_LAUNCH_CHECK();
    }
    return RADMMM_OK;
}

int soft_attention(const float* q, const float* k, const float* prior, const int* in_lens, float* attn,
                   float* attn_logprob, const float* txt_enc, float* context, int B, int Ca, int T1, int T2, int Dt,
                   float temperature, cudaStream_t st) {
    RADMMM_REQUIRE(B > 0 && Ca > 0 && T1 > 0 && T2 > 0, "soft_attention: bad sizes");
    RADMMM_REQUIRE(txt_enc == nullptr || context != nullptr, "soft_attention: context output missing");
    // ks + qs + as, rounded to 16 bytes so that the transposed copy can be read with LDS.128, + ast
    const size_t head = ((size_t)Ca * T2 + (size_t)FROWS * Ca + (size_t)FROWS * T2 + 3) & ~(size_t)3;
    const size_t smem = sizeof(float) * (head + (size_t)FROWS * T2);
    RADMMM_REQUIRE(smem <= 220 * 1024, "soft_attention: T2=%d keys do not fit in shared memory", T2);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        RADMMM_CUDA(cudaFuncSetAttribute(soft_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set[dev & 63] = true;
    }
    int threads = FROWS * 32;                                   // one warp per frame; one thread per text channel if more
    if (txt_enc != nullptr && Dt > threads) threads = (int)round_up(Dt < 1024 ? Dt : 1024, 32);
    dim3 grid(cdiv(T1, FROWS), B);
    soft_attention_kernel<<<grid, threads, smem, st>>>(q, k, prior, in_lens, attn, attn_logprob, txt_enc, context, Ca,
                                                       T1, T2, Dt, temperature);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
