// Fused multi-tensor RAdam + global-norm gradient clipping (reference: radam.py:63-142 -- a Python loop over ~150 tensors with
// fp32 copies and ~12 ATen kernels each -- and Lightning's gradient_clip_val: 1.0 / gradient_clip_algorithm: norm,
// configs/RADMMM_train_config.yaml:7-8, i.e. torch.nn.utils.clip_grad_norm_ before optimizer.step).
//
// Three launches per step for ALL parameters, no host synchronisation, CUDA-graph capturable:
//   1. grad_sumsq:   sum of squares of every gradient (one pass over the gradients, fp64 accumulation)
//   2. radam_hyper:  one thread: step += 1, RAdam's step size / variance-rectification switch, the clip coefficient
//   3. radam_update: exp_avg, exp_avg_sq, weight decay and the parameter update in one pass
// HBM-bound: 4 B (pass 1) + 16 B read + 12 B written (pass 3) per parameter.
// Work is cut into chunks of kChunk elements of ONE tensor; a device table maps chunk -> (tensor, offset).
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int kChunk = 16384;      // elements per CTA
constexpr int kThreads = 256;

struct TensorRec { float* p; const float* g; float* m; float* v; long long n; };

// state (device, doubles): [0] sum of squares, [1] step count, [2] step_size, [3] adaptive (N_sma >= 5), [4] clip coefficient
// cfg (device, doubles):   [0] lr, [1] beta1, [2] beta2, [3] eps, [4] weight_decay, [5] max_grad_norm (<= 0: no clipping)

__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const TensorRec* __restrict__ recs, const int* __restrict__ chunk_tensor,
                                                              const long long* __restrict__ chunk_off, double* __restrict__ state) {
    const TensorRec r = recs[chunk_tensor[blockIdx.x]];
    const long long off = chunk_off[blockIdx.x];
    const int n = (int)min((long long)kChunk, r.n - off);
    const float* g = r.g + off;
    float acc = 0.0f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const int n4 = n >> 2;
        for (int base = 0; base < n4; base += kThreads * 4) {
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = base + j * kThreads + threadIdx.x; t[j] = i < n4 ? __ldg(g4 + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += t[j].x * t[j].x + t[j].y * t[j].y + t[j].z * t[j].z + t[j].w * t[j].w;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) acc += g[i] * g[i];
    }
    double d = warp_sum((double)acc);          // <= 64 products per thread in fp32, everything above in fp64
    __shared__ double red[kThreads / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < kThreads / 32; ++i) a += red[i];
        atomicAdd(state, a);
    }
}

// radam.py:100-122 in double precision (the reference does this arithmetic in Python floats)
__global__ void radam_hyper_kernel(double* __restrict__ state, const double* __restrict__ cfg) {
    const double lr = cfg[0], beta1 = cfg[1], beta2 = cfg[2], max_norm = cfg[5];
    const double step = state[1] + 1.0;
    state[1] = step;
    const double beta2_t = pow(beta2, step);
    const double n_sma_max = 2.0 / (1.0 - beta2) - 1.0;
    const double n_sma = n_sma_max - 2.0 * step * beta2_t / (1.0 - beta2_t);
    double step_size;
    if (n_sma >= 5.0)
        step_size = lr * sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_sma_max - 4.0) * (n_sma - 2.0) / n_sma * n_sma_max / (n_sma_max - 2.0)) /
                    (1.0 - pow(beta1, step));
    else
        step_size = lr / (1.0 - pow(beta1, step));
    state[2] = step_size;
    state[3] = n_sma >= 5.0 ? 1.0 : 0.0;
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    double coef = 1.0;
    if (max_norm > 0.0) coef = fmin(1.0, max_norm / (sqrt(state[0]) + 1e-6));
    state[4] = coef;
    state[5] = sqrt(state[0]);                 // total gradient norm of this step (what Lightning logs)
    state[0] = 0.0;                            // ready for the next step's accumulation
}

__global__ void __launch_bounds__(kThreads) radam_update_kernel(const TensorRec* __restrict__ recs, const int* __restrict__ chunk_tensor,
                                                                const long long* __restrict__ chunk_off, const double* __restrict__ state,
                                                                const double* __restrict__ cfg) {
    const TensorRec r = recs[chunk_tensor[blockIdx.x]];
    const long long off = chunk_off[blockIdx.x];
    const int n = (int)min((long long)kChunk, r.n - off);
    const float lr = (float)cfg[0], beta1 = (float)cfg[1], beta2 = (float)cfg[2], eps = (float)cfg[3], wd = (float)cfg[4];
    const float step_size = (float)state[2], clip = (float)state[4];
    const bool adaptive = state[3] != 0.0;
    const float om1 = 1.0f - beta1, om2 = 1.0f - beta2, decay = -wd * lr;
    float* p = r.p + off;
    const float* g = r.g + off;
    float* m = r.m + off;
    float* v = r.v + off;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= clip;                                            // Lightning clips the gradients in place before step()
        vv = vv * beta2 + om2 * gg * gg;                       // exp_avg_sq.mul_(beta2).addcmul_(1 - beta2, grad, grad)
        mm = mm * beta1 + om1 * gg;                            // exp_avg.mul_(beta1).add_(1 - beta1, grad)
        if (wd != 0.0f) pp = pp + decay * pp;                  // p.add_(-weight_decay * lr, p)
        if (adaptive) pp = pp - step_size * (mm / (sqrtf(vv) + eps));      // p.addcdiv_(-step_size, exp_avg, denom)
        else pp = pp - step_size * mm;                         // p.add_(-step_size, exp_avg)
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
        const int n4 = n >> 2;
        for (int base = 0; base < n4; base += kThreads * 2) {
            float4 tp[2], tg[2], tm[2], tv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = base + j * kThreads + threadIdx.x;
                if (i < n4) {
                    tp[j] = reinterpret_cast<const float4*>(p)[i]; tg[j] = __ldg(reinterpret_cast<const float4*>(g) + i);
                    tm[j] = reinterpret_cast<const float4*>(m)[i]; tv[j] = reinterpret_cast<const float4*>(v)[i];
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = base + j * kThreads + threadIdx.x;
                if (i < n4) {
                    upd(tp[j].x, tg[j].x, tm[j].x, tv[j].x); upd(tp[j].y, tg[j].y, tm[j].y, tv[j].y);
                    upd(tp[j].z, tg[j].z, tm[j].z, tv[j].z); upd(tp[j].w, tg[j].w, tm[j].w, tv[j].w);
                    reinterpret_cast<float4*>(p)[i] = tp[j]; reinterpret_cast<float4*>(m)[i] = tm[j]; reinterpret_cast<float4*>(v)[i] = tv[j];
                }
            }
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) upd(p[i], g[i], m[i], v[i]);
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) upd(p[i], g[i], m[i], v[i]);
    }
}

}  // namespace

int radam_chunk_elems() { return kChunk; }

int radam_step(const void* recs, const int* chunk_tensor, const long long* chunk_off, int n_chunks, double* state, const double* cfg,
               cudaStream_t st) {
    RADMMM_REQUIRE(recs && chunk_tensor && chunk_off && state && cfg && n_chunks > 0, "radam_step: missing tables (n_chunks=%d)", n_chunks);
    const TensorRec* r = reinterpret_cast<const TensorRec*>(recs);
    grad_sumsq_kernel<<<n_chunks, kThreads, 0, st>>>(r, chunk_tensor, chunk_off, state);
    RADMMM_LAUNCH_CHECK();
    radam_hyper_kernel<<<1, 1, 0, st>>>(state, cfg);
    RADMMM_LAUNCH_CHECK();
    radam_update_kernel<<<n_chunks, kThreads, 0, st>>>(r, chunk_tensor, chunk_off, state, cfg);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
