// Shared device/host helpers for the radmmm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace radmmm {

// ---------------------------------------------------------------------------------------------------------
// error reporting: thread-local message, negative return codes (see include/radmmm_b200.h)
// ---------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();
void count_launch();          // bumps the process-wide kernel-launch counter (radmmm_launch_count)

#define RADMMM_OK 0
#define RADMMM_ERR_ARG -1
#define RADMMM_ERR_CUDA -2
#define RADMMM_ERR_UNSUPPORTED -3

#define RADMMM_REQUIRE(cond, ...)                                                                         \
    do {                                                                                                  \
        if (!(cond)) { ::radmmm::set_error(__VA_ARGS__); return RADMMM_ERR_ARG; }                         \
    } while (0)

#define RADMMM_CUDA(expr)                                                                                 \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            ::radmmm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return RADMMM_ERR_CUDA;                                                                       \
        }                                                                                                 \
    } while (0)

#define RADMMM_LAUNCH_CHECK()                                                                             \
    do {                                                                                                  \
        ::radmmm::count_launch();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                              \
        if (_e != cudaSuccess) {                                                                          \
            ::radmmm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return RADMMM_ERR_CUDA;                                                                       \
        }                                                                                                 \
    } while (0)

#define RADMMM_TRY(expr)                                                                                  \
    do { int _rc = (expr); if (_rc != 0) return _rc; } while (0)

__host__ __device__ static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------------------
// Precision modes.  An "act" matrix is how activations / weights are stored for the contraction kernels:
//   MODE_F32    1 plane  float            FFMA kernels, exact-parity path
//   MODE_BF16   1 plane  __nv_bfloat16    tcgen05, throughput path
//   MODE_BF16X3 2 planes __nv_bfloat16    tcgen05, hi + lo split; the GEMM issues hi*hi + lo*hi + hi*lo
// ---------------------------------------------------------------------------------------------------------
enum { MODE_F32 = 0, MODE_BF16 = 1, MODE_BF16X3 = 2 };

__host__ __device__ inline int mode_elem_bytes(int mode) { return mode == MODE_F32 ? 4 : 2; }
__host__ __device__ inline int mode_planes(int mode) { return mode == MODE_BF16X3 ? 2 : 1; }

// Row-major matrix [rows][ld] in act format; plane p starts at base + p*plane_stride elements.
struct ActMat {
    void* ptr;
    long long ld;            // elements per row
    long long plane_stride;  // elements between hi and lo planes (MODE_BF16X3)
};

template <int MODE>
__device__ __forceinline__ void act_store(const ActMat& m, long long idx, float v) {
    if constexpr (MODE == MODE_F32) {
        reinterpret_cast<float*>(m.ptr)[idx] = v;
    } else if constexpr (MODE == MODE_BF16) {
        reinterpret_cast<__nv_bfloat16*>(m.ptr)[idx] = __float2bfloat16_rn(v);
    } else {
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        reinterpret_cast<__nv_bfloat16*>(m.ptr)[idx] = hi;
        reinterpret_cast<__nv_bfloat16*>(m.ptr)[idx + m.plane_stride] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

template <int MODE>
__device__ __forceinline__ float act_load(const ActMat& m, long long idx) {
    if constexpr (MODE == MODE_F32) {
        return reinterpret_cast<const float*>(m.ptr)[idx];
    } else if constexpr (MODE == MODE_BF16) {
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(m.ptr)[idx]);
    } else {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(m.ptr);
        return __bfloat162float(p[idx]) + __bfloat162float(p[idx + m.plane_stride]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Row geometry.  All WN-internal matrices are [R][channels] over "rows": row r = b*pitch + t, t in [0,pitch),
// pitch = T' + gap (gap >= 16 zero rows separate utterances so dilated taps never cross a batch boundary),
// R = round_up(B*pitch, 128).  Row r is valid iff b < B and t < len_b.
// ---------------------------------------------------------------------------------------------------------
struct RowGeom {
    const int* lens;   // device, B grouped lengths
    int B, Tp, pitch, R;
};

__device__ __forceinline__ void row_decode(const RowGeom& g, int r, int& b, int& t, int& len) {
    b = r / g.pitch;
    t = r - b * g.pitch;
    len = (b < g.B) ? min(g.lens[b], g.Tp) : 0;
}

// 16-byte load that stays where it is written: a batch of these followed by a barrier is in flight all at once.  (A batch of
// __ldg / ld.global.nc loads is not: ptxas may move non-coherent loads below a bar.sync and it does, interleaving them with
// their consumers a few at a time to save registers.)
__device__ __forceinline__ float4 ldg_f4_issue(const float4* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// number of in-sequence taps of a k=5 conv with dilation d centred at t (partialconv1d.py:74-77 in closed form)
__device__ __forceinline__ int tap_count(int t, int len, int d) {
    int u = 0;
#pragma unroll
    for (int j = -2; j <= 2; ++j) {
        int s = t + j * d;
        u += (s >= 0 && s < len) ? 1 : 0;
    }
    return u;
}
// PartialConv1d ratio for a valid frame: slide_winsize / (u + 1e-6)   (partialconv1d.py:80)
__device__ __forceinline__ float pconv_ratio(int t, int len, int d) {
    return 5.0f / ((float)tap_count(t, len, d) + 1e-6f);
}

// torch.nn.Softplus(beta=1, threshold=20).  FAST (tensor-core modes) uses the MUFU ex2/lg2 intrinsics: absolute error
// < 2e-7, far below the bf16 / bf16x3 operand rounding; the fp32 checker path keeps the precise libm forms.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// FAST is BRANCH-FREE: softplus(x) = max(x, 0) + ln2 * lg2(1 + 2^(-|x| log2 e)).  For x > 20 the correction is below half an
// ulp of x, i.e. the result IS x, as with torch's threshold.  (The earlier `x > 20 ? x : ...` form compiled to one divergent
// branch per element -- 32 of them per epilogue chunk -- and made the contraction epilogues 3x slower than their stores.)
template <bool FAST>
__device__ __forceinline__ float softplus_f(float x) {
    if constexpr (FAST) return fmaxf(x, 0.0f) + 0.6931471805599453f * lg2_approx(1.0f + ex2_approx(-1.4426950408889634f * fabsf(x)));
    else return x > 20.0f ? x : log1pf(expf(x));
}
// d softplus / dx written in terms of the OUTPUT h = softplus(x):  sigmoid(x) = 1 - exp(-h)   (h > 20: 1 - 2e-9 rounds to 1)
template <bool FAST>
__device__ __forceinline__ float sigmoid_from_softplus(float h) {
    if constexpr (FAST) return 1.0f - ex2_approx(-1.4426950408889634f * h);
    else return h > 20.0f ? 1.0f : -expm1f(-h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace radmmm
