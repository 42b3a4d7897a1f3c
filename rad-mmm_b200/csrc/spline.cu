// Piecewise-quadratic spline coupling transform (splines.py:241-339 as driven by common.py:1040-1090).
// One thread evaluates one (b, channel, frame) element with its 65 parameters (32 widths + 33 heights) held in
// registers; lanes run along time so every parameter load is a coalesced 128-byte line.  The per-frame log-Jacobian
// sum over channels is reduced through shared memory (deterministic order).  HBM-bound on the parameter tensor:
// 65*4 B per element in, 4 B out (forward); the backward writes the 65 parameter gradients back.
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int K = 32;            // bins (decoders.py:56)
constexpr int NB = 2 * K + 1;    // parameters per channel
constexpr float EPS = 1.1920928955078125e-07f;   // torch.finfo(torch.float32).eps

struct SplineState {
    float w[K];       // softmax widths
    float v[K + 1];   // normalised heights
    float e[K + 1];   // exp(v~ - max)
    float S;          // normaliser sum_k (u_k+u_{k+1})/2 w_k
    int vmax_idx;
};

__device__ __forceinline__ void spline_setup(const float* __restrict__ q, long long stride, SplineState& s) {
    float wt[K];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) { wt[k] = q[k * stride]; m = fmaxf(m, wt[k]); }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) { wt[k] = expf(wt[k] - m); sum += wt[k]; }
#pragma unroll
    for (int k = 0; k < K; ++k) s.w[k] = wt[k] / sum;
    float vm = -INFINITY;
    s.vmax_idx = 0;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        s.v[k] = q[(K + k) * stride];
        if (s.v[k] > vm) { vm = s.v[k]; s.vmax_idx = k; }
    }
#pragma unroll
    for (int k = 0; k <= K; ++k) { s.e[k] = expf(s.v[k] - vm); s.v[k] = s.e[k] + 1e-8f; }
    float S = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) S += (s.v[k] + s.v[k + 1]) * 0.5f * s.w[k];
    s.S = S;
#pragma unroll
    for (int k = 0; k <= K; ++k) s.v[k] = s.v[k] / S;
}

// returns bin index; fills the left edges (W_{b-1}, F_{b-1})
__device__ __forceinline__ int spline_bin(const SplineState& s, float x, bool inverse, float& w_left, float& f_left) {
    float wc = 0.0f, fc = 0.0f;
    int bin = K - 1;
    bool found = false;
    w_left = 0.0f; f_left = 0.0f;
    float wl = 0.0f, fl = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        wl = wc; fl = fc;
        wc += s.w[k];
        fc += (s.v[k + 1] + s.v[k]) * 0.5f * s.w[k];
        const float edge = (k == K - 1) ? 1.0f : (inverse ? fc : wc);     // last cumulative value forced to 1
        if (!found && edge >= x) { found = true; bin = k; w_left = wl; f_left = fl; }
    }
    if (!found) { w_left = wl; f_left = fl; }
    return bin;
}

template <typename T>
__device__ __forceinline__ T pick(const T* a, int n, int idx) {      // register-array gather without local memory
    T r = a[0];
#pragma unroll
    for (int k = 1; k < n; ++k) r = (k == idx) ? a[k] : r;
    return r;
}

// grid (ceil(Tp/32), B), block (32, 8)
__global__ void __launch_bounds__(256) spline_fwd_kernel(const float* __restrict__ z1, const float* __restrict__ q,
                                                         float* __restrict__ z1_out, float* __restrict__ log_s, int Ch,
                                                         int Tp, float lo, float hi, int inverse) {
    __shared__ float red[8][33];
    const int t = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    const float range = hi - lo;
    float lj_sum = 0.0f;
    if (t < Tp) {
        for (int c = threadIdx.y; c < Ch; c += 8) {
            const long long zi = ((long long)b * Ch + c) * Tp + t;
            const float x = (z1[zi] - lo) / range;
            float y = x, lj = 0.0f;
            if (x >= 0.0f && x < 1.0f) {
                SplineState s;
                spline_setup(q + ((long long)b * Ch + c) * NB * Tp + t, Tp, s);
                float w_left, f_left;
                const int bin = spline_bin(s, x, inverse != 0, w_left, f_left);
                const float w_b = pick(s.w, K, bin), v_b = pick(s.v, K + 1, bin), v_b1 = pick(s.v, K + 1, bin + 1);
                if (!inverse) {
                    const float alpha = (x - w_left) / fmaxf(w_b, EPS);
                    float cc = alpha * alpha * 0.5f * (v_b1 - v_b) * w_b + alpha * v_b * w_b + f_left;
                    lj = logf(fmaxf(v_b + alpha * (v_b1 - v_b), EPS));
                    y = fminf(fmaxf(cc, EPS), 1.0f - EPS);
                } else {
                    const float a = (v_b1 - v_b) * w_b * 0.5f, bb = v_b * w_b, cc = f_left - x;
                    const float alpha = (-bb + sqrtf(bb * bb - 4.0f * a * cc)) / (2.0f * a);
                    y = fminf(fmaxf(alpha * w_b + w_left, EPS), 1.0f - EPS);
                }
            }
            z1_out[zi] = y * range + lo;
            lj_sum += lj;
        }
    }
    if (inverse || log_s == nullptr) return;
    red[threadIdx.y][threadIdx.x] = lj_sum;
    __syncthreads();
    if (threadIdx.y == 0 && t < Tp) {
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
        log_s[(long long)b * Tp + t] = tot;     // + Ch*(log(top-bottom) - log(right-left)) = 0 for equal ranges
    }
}

// backward of the forward transform: one thread per element, grid (ceil(Tp/128), Ch, B)
__global__ void __launch_bounds__(128) spline_bwd_kernel(const float* __restrict__ z1, const float* __restrict__ q,
                                                         const int* __restrict__ lens, const float* __restrict__ dz1_out,
                                                         const float* __restrict__ dlog_s, float* __restrict__ dz1,
                                                         float* __restrict__ dq, int Ch, int Tp, float lo, float hi) {
    const int t = blockIdx.x * 128 + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
    if (t >= Tp) return;
    const float range = hi - lo;
    const long long zi = ((long long)b * Ch + c) * Tp + t;
    float* dqp = dq + ((long long)b * Ch + c) * NB * Tp + t;
    const bool valid = t < lens[b];
    const float gz = valid ? dz1_out[zi] : 0.0f;
    const float gj = (valid && dlog_s) ? dlog_s[(long long)b * Tp + t] : 0.0f;
    const float x = (z1[zi] - lo) / range;
    if (!(x >= 0.0f && x < 1.0f) || !valid) {
        dz1[zi] = gz;
#pragma unroll 1
        for (int k = 0; k < NB; ++k) dqp[(long long)k * Tp] = 0.0f;
        return;
    }
    SplineState s;
    spline_setup(q + ((long long)b * Ch + c) * NB * Tp + t, Tp, s);
    float w_left, f_left;
    const int bin = spline_bin(s, x, false, w_left, f_left);
    const float w_b = pick(s.w, K, bin), v_b = pick(s.v, K + 1, bin), v_b1 = pick(s.v, K + 1, bin + 1);
    const float w_bc = fmaxf(w_b, EPS);
    const float alpha = (x - w_left) / w_bc;
    const float cc = alpha * alpha * 0.5f * (v_b1 - v_b) * w_b + alpha * v_b * w_b + f_left;
    const float lerp = v_b + alpha * (v_b1 - v_b);
    const float gc = (cc >= EPS && cc <= 1.0f - EPS) ? gz * range : 0.0f;     // clamp passes gradient only inside
    const float ds = (lerp >= EPS) ? gj / lerp : 0.0f;
    const float dalpha = gc * (alpha * (v_b1 - v_b) * w_b + v_b * w_b) + ds * (v_b1 - v_b);
    const float dvb = gc * (alpha * w_b - 0.5f * alpha * alpha * w_b) + ds * (1.0f - alpha);
    const float dvb1 = gc * (0.5f * alpha * alpha * w_b) + ds * alpha;
    const float dwb = gc * (0.5f * alpha * alpha * (v_b1 - v_b) + alpha * v_b) - ((w_b >= EPS) ? dalpha * alpha / w_bc : 0.0f);
    const float dleft = -dalpha / w_bc;             // d / dW_{b-1}
    dz1[zi] = (dalpha / w_bc) / range;
    float dw[K], dv[K + 1];
#pragma unroll
    for (int k = 0; k <= K; ++k) dv[k] = (k == bin) ? dvb : ((k == bin + 1) ? dvb1 : 0.0f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float g = (k == bin) ? dwb : 0.0f;
        if (k < bin) {                          // W_{b-1} and F_{b-1} are sums over bins left of b
            g += dleft + gc * 0.5f * (s.v[k] + s.v[k + 1]);
            dv[k] += gc * 0.5f * s.w[k];
            dv[k + 1] += gc * 0.5f * s.w[k];
        }
        dw[k] = g;
    }
    // v = u / S
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k <= K; ++k) dot += dv[k] * s.v[k];
    const float dS = -dot / s.S;
    float dmax = 0.0f;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        const float wl = (k > 0) ? s.w[k - 1] : 0.0f, wr = (k < K) ? s.w[k] : 0.0f;
        const float du = dv[k] / s.S + dS * 0.5f * (wl + wr);
        const float dvt = du * s.e[k];
        dmax += dvt;
        dv[k] = dvt;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) dw[k] += dS * 0.5f * (s.v[k] + s.v[k + 1]) * s.S;   // u_k = v_k * S
    float wdot = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) wdot += dw[k] * s.w[k];
#pragma unroll
    for (int k = 0; k < K; ++k) dqp[(long long)k * Tp] = s.w[k] * (dw[k] - wdot);
#pragma unroll
    for (int k = 0; k <= K; ++k) dqp[(long long)(K + k) * Tp] = dv[k] - ((k == s.vmax_idx) ? dmax : 0.0f);
}

// ---------------------------------------------------------------------------------------------------------------
// Piecewise-LINEAR transform (splines.py:57-142 forward, 145-238 inverse; SplineTransformationLayer(use_quadratic=False),
// common.py:1019-1020,1069-1075).  NBL un-normalised bin heights per element; q = NBL * softmax(q~) are the slopes.
// forward:  y = (x - m/NBL) q_m + sum_{k<m} q_k / NBL,  m = clamp(floor(NBL x)),  log J = log q_m
// inverse:  m = last bin whose left integral is <= y,  x = (y - left_m) / q_m + m / NBL
// Elements outside [0, 1] pass through (slope 1).  One thread per element, parameters in registers, lanes along time.
// ---------------------------------------------------------------------------------------------------------------
template <int NBL>
__device__ __forceinline__ void linear_setup(const float* __restrict__ q, long long stride, float* p) {
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NBL; ++k) { p[k] = q[k * stride]; m = fmaxf(m, p[k]); }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < NBL; ++k) { p[k] = expf(p[k] - m); sum += p[k]; }
#pragma unroll
    for (int k = 0; k < NBL; ++k) p[k] = p[k] / sum;            // softmax; slope q_k = NBL * p_k
}

// grid (ceil(Tp/32), B), block (32, 8)
template <int NBL>
__global__ void __launch_bounds__(256) spline_linear_kernel(const float* __restrict__ z1, const float* __restrict__ q,
                                                            float* __restrict__ z1_out, float* __restrict__ log_s, int Ch,
                                                            int Tp, float lo, float hi, int inverse) {
    __shared__ float red[8][33];
    const int t = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    const float range = hi - lo, w = 1.0f / NBL;
    float lj_sum = 0.0f;
    if (t < Tp) {
        for (int c = threadIdx.y; c < Ch; c += 8) {
            const long long zi = ((long long)b * Ch + c) * Tp + t;
            const float x = (z1[zi] - lo) / range;
            float y = x, slope = 1.0f;
            if (x >= 0.0f && x <= 1.0f) {
                float p[NBL];
                linear_setup<NBL>(q + ((long long)b * Ch + c) * NBL * Tp + t, Tp, p);
                if (!inverse) {
                    const int m = min(max((int)floorf(NBL * x), 0), NBL - 1);
                    float left = 0.0f;
#pragma unroll
                    for (int k = 0; k < NBL; ++k) left += (k < m) ? p[k] : 0.0f;       // sum_{k<m} q_k w = sum p_k
                    slope = NBL * pick(p, NBL, m);
                    y = (x - m * w) * slope + left;
                } else {
                    // argmin over bins of (y - left_k) among the non-negative ones = the last bin with left_k <= y
                    float left = 0.0f, left_m = 0.0f;
                    int m = 0;
#pragma unroll
                    for (int k = 0; k < NBL; ++k) {
                        if (x - left >= 0.0f) { m = k; left_m = left; }
                        left += p[k];
                    }
                    slope = NBL * pick(p, NBL, m);
                    y = (x - left_m) / slope + m * w;
                }
                y = fminf(fmaxf(y, EPS), 1.0f - EPS);
            }
            z1_out[zi] = y * range + lo;
            lj_sum += logf(slope);
        }
    }
    if (log_s == nullptr) return;
    red[threadIdx.y][threadIdx.x] = lj_sum;
    __syncthreads();
    if (threadIdx.y == 0 && t < Tp) {
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
        log_s[(long long)b * Tp + t] = inverse ? -tot : tot;
    }
}

// backward of the forward linear transform: grid (ceil(Tp/128), Ch, B)
template <int NBL>
__global__ void __launch_bounds__(128) spline_linear_bwd_kernel(const float* __restrict__ z1, const float* __restrict__ q,
                                                                const int* __restrict__ lens, const float* __restrict__ dz1_out,
                                                                const float* __restrict__ dlog_s, float* __restrict__ dz1,
                                                                float* __restrict__ dq, int Ch, int Tp, float lo, float hi) {
    const int t = blockIdx.x * 128 + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
    if (t >= Tp) return;
    const float range = hi - lo, w = 1.0f / NBL;
    const long long zi = ((long long)b * Ch + c) * Tp + t;
    float* dqp = dq + ((long long)b * Ch + c) * NBL * Tp + t;
    const bool valid = t < lens[b];
    const float gz = valid ? dz1_out[zi] : 0.0f;
    const float gj = (valid && dlog_s) ? dlog_s[(long long)b * Tp + t] : 0.0f;
    const float x = (z1[zi] - lo) / range;
    if (!(x >= 0.0f && x <= 1.0f) || !valid) {
        dz1[zi] = gz;
#pragma unroll 1
        for (int k = 0; k < NBL; ++k) dqp[(long long)k * Tp] = 0.0f;
        return;
    }
    float p[NBL];
    linear_setup<NBL>(q + ((long long)b * Ch + c) * NBL * Tp + t, Tp, p);
    const int m = min(max((int)floorf(NBL * x), 0), NBL - 1);
    float left = 0.0f;
#pragma unroll
    for (int k = 0; k < NBL; ++k) left += (k < m) ? p[k] : 0.0f;
    const float pm = pick(p, NBL, m), slope = NBL * pm, alpha = x - m * w;
    const float y = alpha * slope + left;
    const float gy = (y >= EPS && y <= 1.0f - EPS) ? gz * range : 0.0f;         // the clamp passes gradient only inside
    dz1[zi] = gy * slope / range;
    // dL/dp_k: y = NBL alpha p_m + sum_{k<m} p_k ; log J = log(NBL p_m)
    float dp[NBL], dot = 0.0f;
#pragma unroll
    for (int k = 0; k < NBL; ++k) {
        dp[k] = (k < m) ? gy : ((k == m) ? gy * NBL * alpha + gj / pm : 0.0f);
        dot += dp[k] * p[k];
    }
#pragma unroll
    for (int k = 0; k < NBL; ++k) dqp[(long long)k * Tp] = p[k] * (dp[k] - dot);     // softmax Jacobian
}

}  // namespace

template <int NBL>
static void launch_linear(const float* z1, const float* q, float* z1_out, float* log_s, int B, int Ch, int Tp, float lo, float hi,
                          int inverse, cudaStream_t st) {
    dim3 grid(cdiv(Tp, 32), B), block(32, 8);
    spline_linear_kernel<NBL><<<grid, block, 0, st>>>(z1, q, z1_out, log_s, Ch, Tp, lo, hi, inverse);
}
template <int NBL>
static void launch_linear_bwd(const float* z1, const float* q, const int* lens, const float* dz1_out, const float* dlog_s, float* dz1,
                              float* dq, int B, int Ch, int Tp, float lo, float hi, cudaStream_t st) {
    dim3 grid(cdiv(Tp, 128), Ch, B);
    spline_linear_bwd_kernel<NBL><<<grid, 128, 0, st>>>(z1, q, lens, dz1_out, dlog_s, dz1, dq, Ch, Tp, lo, hi);
}

int spline_linear(const float* z1, const float* q, const int* lens, float* z1_out, float* log_s, int B, int Ch, int Tp, int n_bins,
                  float lo, float hi, int inverse, cudaStream_t st) {
    (void)lens;
    RADMMM_REQUIRE(hi > lo, "spline_linear: bad bounds");
    switch (n_bins) {
        case 8: launch_linear<8>(z1, q, z1_out, log_s, B, Ch, Tp, lo, hi, inverse, st); break;
        case 16: launch_linear<16>(z1, q, z1_out, log_s, B, Ch, Tp, lo, hi, inverse, st); break;
        case 32: launch_linear<32>(z1, q, z1_out, log_s, B, Ch, Tp, lo, hi, inverse, st); break;
        default: RADMMM_REQUIRE(false, "spline_linear: n_bins=%d (8, 16 or 32)", n_bins);
    }
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int spline_linear_bwd(const float* z1, const float* q, const int* lens, const float* dz1_out, const float* dlog_s, float* dz1, float* dq,
                      int B, int Ch, int Tp, int n_bins, float lo, float hi, cudaStream_t st) {
    switch (n_bins) {
        case 8: launch_linear_bwd<8>(z1, q, lens, dz1_out, dlog_s, dz1, dq, B, Ch, Tp, lo, hi, st); break;
        case 16: launch_linear_bwd<16>(z1, q, lens, dz1_out, dlog_s, dz1, dq, B, Ch, Tp, lo, hi, st); break;
        case 32: launch_linear_bwd<32>(z1, q, lens, dz1_out, dlog_s, dz1, dq, B, Ch, Tp, lo, hi, st); break;
        default: RADMMM_REQUIRE(false, "spline_linear_bwd: n_bins=%d (8, 16 or 32)", n_bins);
    }
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int spline_fwd(const float* z1, const float* q, const int* lens, float* z1_out, float* log_s, int B, int Ch, int Tp,
               int n_bins, float lo, float hi, int inverse, cudaStream_t st) {
    (void)lens;
    RADMMM_REQUIRE(n_bins == K, "spline: only %d bins are supported (got %d)", K, n_bins);
    RADMMM_REQUIRE(hi > lo, "spline: bad bounds");
    dim3 grid(cdiv(Tp, 32), B), block(32, 8);
    spline_fwd_kernel<<<grid, block, 0, st>>>(z1, q, z1_out, log_s, Ch, Tp, lo, hi, inverse);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int spline_bwd(const float* z1, const float* q, const int* lens, const float* dz1_out, const float* dlog_s, float* dz1,
               float* dq, int B, int Ch, int Tp, int n_bins, float lo, float hi, cudaStream_t st) {
    RADMMM_REQUIRE(n_bins == K, "spline: only %d bins are supported (got %d)", K, n_bins);
    dim3 grid(cdiv(Tp, 128), Ch, B);
    spline_bwd_kernel<<<grid, 128, 0, st>>>(z1, q, lens, dz1_out, dlog_s, dz1, dq, Ch, Tp, lo, hi);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
