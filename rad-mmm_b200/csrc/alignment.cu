// Hard alignment on the GPU.
//
//  * mas_kernel: monotonic alignment search with width 1 (alignment.py:31-59 `mas_width1`), batched over the utterances of
//    a step (the reference copies the (B,1,T1,T2) attention to the host and runs B numba loops one after the other,
//    tts_lightning_modules.py:270-284 `binarize_attention`).  One CTA per utterance, one thread per text position: the
//    Viterbi row update is a wavefront over the text axis with one CTA barrier per mel frame, the back pointers are one BIT
//    per cell (moved left / stayed) kept in shared memory, the backtrack is a 1-thread pointer chase through those bits and
//    the 0/1 map is written with coalesced stores by the whole CTA.  Arithmetic is the reference's: float32 log-probabilities,
//    one float32 add per cell, `>=` tie-break towards the left neighbour, row 0 forced to text position 0, and the
//    reference's extra `opt[0, 0] = 1` after the backtrack loop.
//  * attention_ctc_kernel: AttentionCTCLoss (loss.py:112-140): per utterance a CTC negative log-likelihood of the target
//    1..K over classes [blank | K keys] with the blank logit padded in front, `reduction='mean'`, `zero_infinity=True`,
//    averaged over the batch.  One CTA per utterance computes the per-frame log-sum-exp, the alpha pass, the beta pass and
//    the gradient w.r.t. attn_logprob in one launch (the label-state alphas are parked in the gradient buffer between the
//    passes, so there is no scratch); the reference runs B separate log_softmax + CTCLoss calls.
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int kMasDepth = 8;        // rows of attention prefetched ahead of the row update

__device__ __forceinline__ float mas_log(float x, int is_log) {
    // the reference takes np.log of float32 probabilities (libm logf, correctly rounded in all but rare cases): a double
    // logarithm rounded to float reproduces that; device logf (1 ulp) would flip near-ties of the Viterbi comparison
    return is_log ? x : (float)log((double)x);
}

template <int COLS>
__global__ void __launch_bounds__(1024) mas_kernel(const float* __restrict__ attn, const int* __restrict__ in_lens,
                                                   const int* __restrict__ out_lens, float* __restrict__ out, int T1max, int T2max,
                                                   int is_log, uint32_t* __restrict__ gbits, int words) {
    extern __shared__ uint8_t mas_smem[];
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.x;
    const int T1 = max(0, min(out_lens[b], T1max)), T2 = max(0, min(in_lens[b], T2max));
    float* rowbuf = reinterpret_cast<float*>(mas_smem);                 // [2][COLS*NT + 1], entry 0 of each = -inf sentinel
    const int rowlen = COLS * NT + 1;
    int* path = reinterpret_cast<int*>(rowbuf + 2 * rowlen);            // [T1max]
    uint32_t* bits = gbits != nullptr ? gbits + (size_t)b * T1max * words : reinterpret_cast<uint32_t*>(path + T1max);
    const float* A = attn + (size_t)b * T1max * T2max;
    float* O = out + (size_t)b * T1max * T2max;
    const float NEG = -INFINITY;

    if (T1 > 0 && T2 > 0) {
        if (tid == 0) { rowbuf[0] = NEG; rowbuf[rowlen] = NEG; }
        // row 0: log_p[0, 0] = log attn[0, 0], -inf elsewhere; no back pointers
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const int j = tid + c * NT;
            rowbuf[1 + j] = (j == 0) ? mas_log(A[0], is_log) : NEG;
            if (lane == 0 && (j >> 5) < words) bits[j >> 5] = 0u;
        }
        float nxt[COLS][kMasDepth];
        auto fetch = [&](int i0) {
#pragma unroll
            for (int d = 0; d < kMasDepth; ++d)
#pragma unroll
                for (int c = 0; c < COLS; ++c) {
                    const int i = i0 + d, j = tid + c * NT;
                    nxt[c][d] = (i < T1 && j < T2) ? A[(size_t)i * T2max + j] : 1.0f;
                }
        };
        fetch(1);
        __syncthreads();
        for (int i0 = 1; i0 < T1; i0 += kMasDepth) {
            float cur[COLS][kMasDepth];
#pragma unroll
            for (int d = 0; d < kMasDepth; ++d)
#pragma unroll
                for (int c = 0; c < COLS; ++c) cur[c][d] = mas_log(nxt[c][d], is_log);
            fetch(i0 + kMasDepth);
#pragma unroll
            for (int d = 0; d < kMasDepth; ++d) {
                const int i = i0 + d;
                if (i >= T1) break;                       // uniform
                const float* prev = rowbuf + ((i - 1) & 1) * rowlen + 1;
                float* now = rowbuf + (i & 1) * rowlen + 1;
#pragma unroll
                for (int c = 0; c < COLS; ++c) {
                    const int j = tid + c * NT;
                    const float stay = prev[j], left = prev[j - 1];       // prev[-1] is the -inf sentinel, never taken at j = 0
                    const bool move = (j >= 1) && (left >= stay);
                    now[j] = cur[c][d] + (move ? left : stay);
                    const uint32_t w = __ballot_sync(0xffffffffu, move && j < T2);
                    if (lane == 0 && (j >> 5) < words) bits[(size_t)i * words + (j >> 5)] = w;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (tid == 0) {
            int curj = T2 - 1;
            for (int i = T1 - 1; i >= 0; --i) {
                path[i] = curj;
                curj -= (int)((bits[(size_t)i * words + (curj >> 5)] >> (curj & 31)) & 1u);
            }
        }
        __syncthreads();
    }
    // the 0/1 map of the whole padded slab
    const int total = T1max * T2max;
    for (int e = tid; e < total; e += NT) {
        const int i = e / T2max, j = e - i * T2max;
        float v = 0.0f;
        if (i < T1 && j < T2 && (j == path[i] || (i == 0 && j == 0))) v = 1.0f;
        O[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------- attention CTC loss
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

constexpr int kCtcDepth = 4;

// Thread i < K owns the blank state 2i and the label state 2i+1 (label i+1 = key i); thread K owns the final blank 2K.
__global__ void __launch_bounds__(1024) attention_ctc_kernel(const float* __restrict__ logprob, const int* __restrict__ in_lens,
                                                             const int* __restrict__ out_lens, float* __restrict__ cost,
                                                             float* __restrict__ grad, int B, int T1max, int T2max, float blank_logprob) {
    extern __shared__ float ctc_smem[];
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
    const int b = blockIdx.x;
    const int K = max(0, min(in_lens[b], T2max)), T = max(0, min(out_lens[b], T1max));
    float* lse = ctc_smem;                          // [T1max] log-sum-exp over [blank | K keys] per frame
    float* xb = lse + T1max;                        // [2][NT + 2]: label-state exchange (alpha: i-1, beta: i+1), both kinds for beta
    float* xl = xb + 2 * (NT + 2);                  // [2][NT + 2]
    __shared__ float s_nll;
    const float* X = logprob + (size_t)b * T1max * T2max;
    float* G = grad + (size_t)b * T1max * T2max;
    const float NEG = -INFINITY;
    const int xs = NT + 2;

    // per-frame normaliser (nn.LogSoftmax over the K + 1 classes of this utterance)
    for (int t = wid; t < T; t += nw) {
        float m = blank_logprob;
        for (int k = lane; k < K; k += 32) m = fmaxf(m, X[(size_t)t * T2max + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = (lane == 0) ? expf(blank_logprob - m) : 0.0f;
        for (int k = lane; k < K; k += 32) s += expf(X[(size_t)t * T2max + k] - m);
        s = warp_sum(s);
        if (lane == 0) lse[t] = m + logf(s);
    }
    // exchange buffers start at -inf (slot 0 = "thread -1", slot NT+1 = "thread NT")
    for (int e = tid; e < 2 * xs; e += NT) { xb[e] = NEG; xl[e] = NEG; }
    __syncthreads();

    const bool has_label = tid < K, has_blank = tid <= K;
    float nll = INFINITY;
    if (T > 0 && K > 0) {
        // ---------------------------------------------------------------- alpha pass
        float aB = NEG, aL = NEG;
        float xn[kCtcDepth];
        auto fetch_x = [&](int t0, int dir) {
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) {
                const int t = t0 + dir * d;
                xn[d] = (has_label && t >= 0 && t < T) ? X[(size_t)t * T2max + tid] : 0.0f;
            }
        };
        fetch_x(0, 1);
        for (int t0 = 0; t0 < T; t0 += kCtcDepth) {
            float xc[kCtcDepth];
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) xc[d] = xn[d];
            fetch_x(t0 + kCtcDepth, 1);
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) {
                const int t = t0 + d;
                if (t >= T) break;
                const float l = lse[t];
                const float lpB = blank_logprob - l, lpL = xc[d] - l;
                float nB, nL;
                if (t == 0) {
                    nB = (tid == 0) ? lpB : NEG;
                    nL = (tid == 0 && has_label) ? lpL : NEG;
                } else {
                    const float pL = xl[((t - 1) & 1) * xs + tid];           // label state of thread tid - 1 at t - 1
                    nB = has_blank ? lpB + lse2(aB, pL) : NEG;
                    nL = has_label ? lpL + lse3(aL, aB, pL) : NEG;
                }
                aB = nB; aL = nL;
                xl[(t & 1) * xs + tid + 1] = aL;
                if (has_label) G[(size_t)t * T2max + tid] = aL;              // parked until the beta pass
                __syncthreads();
            }
        }
        if (tid == K) s_nll = -lse2(aB, xl[((T - 1) & 1) * xs + tid]);       // final blank and last label
        __syncthreads();
        nll = s_nll;
        // ---------------------------------------------------------------- beta pass + gradient
        const bool finite = nll < INFINITY && nll == nll;
        const float scale = 1.0f / ((float)max(K, 1) * (float)B);
        for (int e = tid; e < 2 * xs; e += NT) { xb[e] = NEG; xl[e] = NEG; }
        __syncthreads();
        float bB = NEG, bL = NEG;
        float an[kCtcDepth];
        auto fetch_a = [&](int t0) {
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) {
                const int t = t0 - d;
                an[d] = (has_label && t >= 0) ? G[(size_t)t * T2max + tid] : 0.0f;
            }
        };
        fetch_x(T - 1, -1);
        fetch_a(T - 1);
        for (int t0 = T - 1; t0 >= 0; t0 -= kCtcDepth) {
            float xc[kCtcDepth], ac[kCtcDepth];
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) { xc[d] = xn[d]; ac[d] = an[d]; }
            fetch_x(t0 - kCtcDepth, -1);
            fetch_a(t0 - kCtcDepth);
#pragma unroll
            for (int d = 0; d < kCtcDepth; ++d) {
                const int t = t0 - d;
                if (t < 0) break;
                const float l = lse[t];
                const float lpB = blank_logprob - l, lpL = xc[d] - l;
                float nB, nL;
                if (t == T - 1) {
                    nB = (tid == K) ? lpB : NEG;
                    nL = (tid == K - 1) ? lpL : NEG;
                } else {
                    const int src = ((t + 1) & 1) * xs + tid + 2;            // thread tid + 1 at t + 1
                    const float qB = xb[src], qL = xl[src];
                    nB = has_blank ? lpB + lse2(bB, bL) : NEG;               // blank 2i -> itself or its label 2i+1
                    nL = has_label ? lpL + lse3(bL, qB, (tid + 1 < K) ? qL : NEG) : NEG;
                    if (tid == K) nB = lpB + bB;                             // the final blank only continues into itself
                }
                bB = nB; bL = nL;
                xb[(t & 1) * xs + tid + 1] = bB;
                xl[(t & 1) * xs + tid + 1] = bL;
                if (has_label) {
                    float g = 0.0f;
                    if (finite) g = (expf(lpL) - expf(ac[d] + bL - lpL + nll)) * scale;
                    G[(size_t)t * T2max + tid] = g;
                }
                __syncthreads();
            }
        }
    }
    // zero the rest of the padded slab: frames >= T, keys >= K
    const int total = T1max * T2max;
    for (int e = tid; e < total; e += NT) {
        const int t = e / T2max, k = e - t * T2max;
        if (t >= T || k >= K || T == 0 || K == 0) G[e] = 0.0f;
    }
    if (tid == 0) {
        float c = 0.0f;
        if (T > 0 && K > 0 && nll < INFINITY && nll == nll) c = nll / (float)K;      // reduction='mean' divides by the target length
        cost[b] = c;
    }
}

}  // namespace

long long mas_workspace_bytes(int B, int T1, int T2) {
    const int nt = (int)round_up(T2 < 32 ? 32 : T2, 32);
    const int NT = nt > 1024 ? 1024 : nt;
    const int cols = cdiv(T2, NT);
    const int words = cdiv(T2, 32);
    const size_t smem = (size_t)2 * (cols * NT + 1) * 4 + (size_t)T1 * 4 + (size_t)T1 * words * 4;
    return smem <= 200 * 1024 ? 0 : (long long)B * T1 * words * 4;
}

int mas_width1(const float* attn, const int* in_lens, const int* out_lens, float* out, int B, int T1, int T2, int is_log,
               void* workspace, long long workspace_bytes, cudaStream_t st) {
    RADMMM_REQUIRE(B >= 0 && T1 >= 0 && T2 >= 0, "mas: bad shape B=%d T1=%d T2=%d", B, T1, T2);
    if (B == 0 || T1 == 0 || T2 == 0) return RADMMM_OK;
    RADMMM_REQUIRE(T2 <= 4096, "mas: at most 4096 text positions (got %d)", T2);
    RADMMM_REQUIRE(attn && in_lens && out_lens && out, "mas: null argument");
    const int nt = (int)round_up(T2 < 32 ? 32 : T2, 32);
    const int NT = nt > 1024 ? 1024 : nt;
    const int cols = cdiv(T2, NT);
    const int words = cdiv(T2, 32);
    const long long need = mas_workspace_bytes(B, T1, T2);
    RADMMM_REQUIRE(need == 0 || (workspace != nullptr && workspace_bytes >= need), "mas: workspace of %lld bytes needed (radmmm_mas_workspace_bytes)", need);
    const int C = cols <= 1 ? 1 : (cols <= 2 ? 2 : 4);
    size_t smem = (size_t)2 * (C * NT + 1) * 4 + (size_t)T1 * 4 + (need == 0 ? (size_t)T1 * words * 4 : 0);
    RADMMM_REQUIRE(smem <= 220 * 1024, "mas: %d mel frames do not fit the backtrack buffer", T1);
    uint32_t* gb = need == 0 ? nullptr : reinterpret_cast<uint32_t*>(workspace);
#define RADMMM_MAS_LAUNCH(CC)                                                                                                   \
    do {                                                                                                                        \
        static bool attr_set[64] = {};                                                                                          \
        int dev = 0;                                                                                                            \
        cudaGetDevice(&dev);                                                                                                    \
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {                                                                           \
            RADMMM_CUDA(cudaFuncSetAttribute(mas_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));         \
            attr_set[dev] = true;                                                                                               \
        }                                                                                                                       \
        mas_kernel<CC><<<B, NT, smem, st>>>(attn, in_lens, out_lens, out, T1, T2, is_log, gb, words);                           \
    } while (0)
    if (C == 1) RADMMM_MAS_LAUNCH(1);
    else if (C == 2) RADMMM_MAS_LAUNCH(2);
    else RADMMM_MAS_LAUNCH(4);
#undef RADMMM_MAS_LAUNCH
    RADMMM_CUDA(cudaGetLastError());
    count_launch();
    return RADMMM_OK;
}

int attention_ctc(const float* logprob, const int* in_lens, const int* out_lens, float* cost, float* grad, int B, int T1, int T2,
                  float blank_logprob, cudaStream_t st) {
    RADMMM_REQUIRE(B >= 0 && T1 >= 0 && T2 >= 0, "attention_ctc: bad shape B=%d T1=%d T2=%d", B, T1, T2);
    if (B == 0) return RADMMM_OK;
    RADMMM_REQUIRE(logprob && in_lens && out_lens && cost && grad, "attention_ctc: null argument");
    RADMMM_REQUIRE(T2 <= 1023, "attention_ctc: at most 1023 keys (got %d)", T2);
    const int NT = (int)round_up(T2 + 1 < 32 ? 32 : T2 + 1, 32);
    const size_t smem = ((size_t)T1 + 4 * (NT + 2)) * 4;
    RADMMM_REQUIRE(smem <= 200 * 1024, "attention_ctc: %d mel frames do not fit", T1);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        RADMMM_CUDA(cudaFuncSetAttribute(attention_ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[dev] = true;
    }
    attention_ctc_kernel<<<B, NT, smem, st>>>(logprob, in_lens, out_lens, cost, grad, B, T1, T2, blank_logprob);
    RADMMM_CUDA(cudaGetLastError());
    count_launch();
    return RADMMM_OK;
}

}  // namespace radmmm
