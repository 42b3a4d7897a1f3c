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

constexpr int ROWS = 8;   // query frames (warps) per CTA

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

// Backward of soft_attention_kernel.  One warp per query frame recomputes its logits from q/k (nothing but attn and the
// optional prior-normalised log-probabilities' inputs are needed), forms d logit from d attn (+ the context matmul's
// contribution txt_enc^T d context) and d attn_logprob, and accumulates dq (own frame, plain store) and dk (shared
// across query frames: shared-memory accumulation per CTA, then one atomic per element).
__global__ void __launch_bounds__(ROWS * 32) soft_attention_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ prior,
    const int* __restrict__ in_lens, const float* __restrict__ attn, const float* __restrict__ dattn,
    const float* __restrict__ dlogprob, const float* __restrict__ txt_enc, const float* __restrict__ dcontext,
    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dtxt, int Ca, int T1, int T2, int Dt, float temp) {
    extern __shared__ float sm[];
    float* ks = sm;                          // [Ca][T2]
    float* dks = ks + (size_t)Ca * T2;       // [Ca][T2] accumulated over this CTA's query frames
    float* qs = dks + (size_t)Ca * T2;       // [ROWS][Ca]
    float* gs = qs + ROWS * Ca;              // [ROWS][T2]  d dist (per frame)
    float* as = gs + ROWS * T2;              // [ROWS][T2]  attn rows (for dtxt)
    const int b = blockIdx.y, t1_0 = blockIdx.x * ROWS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* kb = k + (long long)b * Ca * T2;
    for (int i = tid; i < Ca * T2; i += ROWS * 32) { ks[i] = kb[i]; dks[i] = 0.0f; }
    for (int i = tid; i < ROWS * Ca; i += ROWS * 32) {
        const int r = i / Ca, c = i % Ca, t1 = t1_0 + r;
        qs[i] = (t1 < T1) ? q[((long long)b * Ca + c) * T1 + t1] : 0.0f;
    }
    __syncthreads();
    const int t1 = t1_0 + wid;
    const int len = min(in_lens[b], T2);
    float* grow = gs + (size_t)wid * T2;
    float* arow = as + (size_t)wid * T2;
    if (t1 < T1) {
        const long long orow = ((long long)b * T1 + t1) * T2;
        // d attn including the context matmul: dA[t2] = dattn[t2] + sum_d txt[d,t2] dctx[d,t1]
        float dot = 0.0f;
        for (int t2 = lane; t2 < T2; t2 += 32) {
            float da = dattn ? dattn[orow + t2] : 0.0f;
            if (txt_enc != nullptr && t2 < len) {
                float acc = 0.0f;
                for (int d = 0; d < Dt; ++d) acc = fmaf(txt_enc[((long long)b * Dt + d) * T2 + t2], dcontext[((long long)b * Dt + d) * T1 + t1], acc);
                da += acc;
            }
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
                arow[t2] = lg;                       // reuse as logits (attn no longer needed for this frame's dq/dk)
                mx = fmaxf(mx, lg);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float se = 0.0f;
            for (int t2 = lane; t2 < T2; t2 += 32) se += expf(arow[t2] - mx);
            se = warp_sum(se);
            for (int t2 = lane; t2 < T2; t2 += 32) grow[t2] -= expf(arow[t2] - mx) / se * gsum;
            for (int t2 = lane; t2 < T2; t2 += 32) arow[t2] = attn[orow + t2];     // restore for dtxt
        }
        for (int t2 = lane; t2 < T2; t2 += 32) grow[t2] *= -temp;                   // d dist
        __syncwarp();
        // dq[c] = sum_t2 2 (q - k) ddist ; dk[c,t2] -= 2 (q - k) ddist
        for (int c = 0; c < Ca; ++c) {
            const float qc = qs[wid * Ca + c];
            float acc = 0.0f;
            for (int t2 = lane; t2 < T2; t2 += 32) {
                const float v = 2.0f * (qc - ks[c * T2 + t2]) * grow[t2];
                acc += v;
                atomicAdd(&dks[c * T2 + t2], -v);
            }
            acc = warp_sum(acc);
            if (lane == 0) dq[((long long)b * Ca + c) * T1 + t1] = acc;
        }
    } else {
        for (int t2 = lane; t2 < T2; t2 += 32) arow[t2] = 0.0f;
    }
    __syncthreads();
    float* dkb = dk + (long long)b * Ca * T2;
    for (int i = tid; i < Ca * T2; i += ROWS * 32)
        if (dks[i] != 0.0f) atomicAdd(dkb + i, dks[i]);
    if (dtxt != nullptr) {
        // dtxt[d,t2] += sum over this CTA's frames of dctx[d,t1] attn[t1,t2]
        for (int i = tid; i < Dt * T2; i += ROWS * 32) {
            const int d = i / T2, t2 = i % T2;
            if (t2 >= len) continue;
            float acc = 0.0f;
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (t1_0 + r < T1) acc = fmaf(dcontext[((long long)b * Dt + d) * T1 + t1_0 + r], as[r * T2 + t2], acc);
            atomicAdd(dtxt + ((long long)b * Dt + d) * T2 + t2, acc);
        }
    }
}

}  // namespace

int soft_attention_bwd(const float* q, const float* k, const float* prior, const int* in_lens, const float* attn,
                       const float* dattn, const float* dlogprob, const float* txt_enc, const float* dcontext, float* dq,
                       float* dk, float* dtxt, int B, int Ca, int T1, int T2, int Dt, float temperature, cudaStream_t st) {
    RADMMM_REQUIRE(B > 0 && Ca > 0 && T1 > 0 && T2 > 0, "soft_attention_bwd: bad sizes");
    RADMMM_REQUIRE((txt_enc == nullptr) == (dcontext == nullptr), "soft_attention_bwd: txt_enc and dcontext go together");
    const size_t smem = sizeof(float) * (2 * (size_t)Ca * T2 + (size_t)ROWS * Ca + 2 * (size_t)ROWS * T2);
    RADMMM_REQUIRE(smem <= 220 * 1024, "soft_attention_bwd: T2=%d keys do not fit in shared memory", T2);
    if (smem > 48 * 1024)
        RADMMM_CUDA(cudaFuncSetAttribute(soft_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RADMMM_CUDA(cudaMemsetAsync(dk, 0, sizeof(float) * (size_t)B * Ca * T2, st));
    if (dtxt) RADMMM_CUDA(cudaMemsetAsync(dtxt, 0, sizeof(float) * (size_t)B * Dt * T2, st));
    dim3 grid(cdiv(T1, ROWS), B);
    soft_attention_bwd_kernel<<<grid, ROWS * 32, smem, st>>>(q, k, prior, in_lens, attn, dattn, dlogprob, txt_enc, dcontext,
                                                             dq, dk, dtxt, Ca, T1, T2, Dt, temperature);
    RADMMM_LAUNCH_CHECK();
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
