// HBM-bound kernels of the flow step: layout packing, affine coupling (fwd / inverse / bwd), invertible 1x1
// convolution, flow-NLL reduction, weight-norm preparation and its backward, bias-gradient column sums.
// All are sized so that a warp touches consecutive addresses along the contiguous (time or channel) axis.
#include <stdlib.h>
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

// =========================================================================================================
// cf (B, C, Tp) fp32  ->  act rows [R][ld]; invalid rows and pad columns are zero.
// 32x32 tile transpose through shared memory: reads coalesced along t, row writes coalesced along c.
// =========================================================================================================
template <int MODE>
__global__ void rows_from_cf_kernel(const float* __restrict__ src, long long batch_stride, int n_ch, RowGeom g,
                                    ActMat dst, int n_cols, int mask_invalid) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        int b, t, len;
        row_decode(g, r, b, t, len);
        float v = 0.0f;
        const bool ok = mask_invalid ? (t < len) : (b < g.B && t < g.Tp);
        if (c < n_ch && ok) v = src[(long long)b * batch_stride + (long long)c * g.Tp + t];
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        if (c < n_cols) act_store<MODE>(dst, (long long)r * dst.ld + c, tile[tx][i]);
    }
}

int rows_from_cf(int mode, const float* src, long long batch_stride, int n_ch, const RowGeom& g, ActMat dst,
                 int n_cols, int mask_invalid, cudaStream_t st) {
    dim3 grid(g.R / 32, cdiv(n_cols, 32)), block(32, 8);
    if (mode == MODE_F32) rows_from_cf_kernel<MODE_F32><<<grid, block, 0, st>>>(src, batch_stride, n_ch, g, dst, n_cols, mask_invalid);
    else if (mode == MODE_BF16) rows_from_cf_kernel<MODE_BF16><<<grid, block, 0, st>>>(src, batch_stride, n_ch, g, dst, n_cols, mask_invalid);
    else rows_from_cf_kernel<MODE_BF16X3><<<grid, block, 0, st>>>(src, batch_stride, n_ch, g, dst, n_cols, mask_invalid);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// fp32 (B, Tp, D) (the context bi-LSTM output, channels-last)  ->  act rows [R][ld]
// =========================================================================================================
template <int MODE>
__global__ void rows_from_btd_kernel(const float* __restrict__ src, int D, RowGeom g, ActMat dst, int n_cols) {
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        int b, t, len;
        row_decode(g, r, b, t, len);
        float v = 0.0f;
        if (c < D && t < len) v = src[((long long)b * g.Tp + t) * D + c];
        if (c < n_cols) act_store<MODE>(dst, (long long)r * dst.ld + c, v);
    }
}

int rows_from_btd(int mode, const float* src, int D, const RowGeom& g, ActMat dst, int n_cols, cudaStream_t st) {
    dim3 grid(g.R / 32, cdiv(n_cols, 32)), block(32, 8);
    if (mode == MODE_F32) rows_from_btd_kernel<MODE_F32><<<grid, block, 0, st>>>(src, D, g, dst, n_cols);
    else if (mode == MODE_BF16) rows_from_btd_kernel<MODE_BF16><<<grid, block, 0, st>>>(src, D, g, dst, n_cols);
    else rows_from_btd_kernel<MODE_BF16X3><<<grid, block, 0, st>>>(src, D, g, dst, n_cols);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// fp32 rows [R][ld] -> fp32 (B, Tp, D) accumulate (gradient of the context back to the LSTM output layout)
__global__ void btd_from_rows_kernel(const float* __restrict__ rows, long long ld, int D, RowGeom g,
                                     float* __restrict__ dst, int accumulate) {
    const int r = blockIdx.x;
    int b, t, len;
    row_decode(g, r, b, t, len);
    if (b >= g.B || t >= g.Tp) return;
    float* o = dst + ((long long)b * g.Tp + t) * D;
    const float* s = rows + (long long)r * ld;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float v = (t < len) ? s[c] : 0.0f;
        o[c] = accumulate ? o[c] + v : v;
    }
}

int btd_from_rows(const float* rows, long long ld, int D, const RowGeom& g, float* dst, int accumulate,
                  cudaStream_t st) {
    btd_from_rows_kernel<<<g.B * g.pitch, 256, 0, st>>>(rows, ld, D, g, dst, accumulate);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// Affine coupling (common.py:1127-1185).  params (B, C, Tp): [:Ch] = scale pre-activation a, [Ch:] = shift b.
// =========================================================================================================
__device__ __forceinline__ void scale_fn(int fn, float a, float& s, float& log_s, float& ds_da) {
    if (fn == SCALE_TANH) {
        float th = tanhf(a);
        s = th + 1.0f + 1e-6f;
        log_s = logf(s);
        ds_da = 1.0f - th * th;
    } else if (fn == SCALE_EXP) {
        s = expf(a);
        log_s = a;
        ds_da = s;
    } else if (fn == SCALE_SIGMOID) {
        float sg = 1.0f / (1.0f + expf(-(a + 10.0f)));
        s = sg + 1e-6f;
        log_s = logf(s);
        ds_da = sg * (1.0f - sg);
    } else {  // translate
        s = 1.0f;
        log_s = 0.0f;
        ds_da = 0.0f;
    }
}

// grid: (ceil(Tp/256), Ch, B)
__global__ void coupling_fwd_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                    float* __restrict__ z_out, float* __restrict__ log_s, int C, int Tp, int fn,
                                    int inverse) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    const int c = blockIdx.y, b = blockIdx.z, Ch = C / 2;
    const long long i0 = ((long long)b * C + c) * Tp + t, i1 = ((long long)b * C + Ch + c) * Tp + t;
    float s, ls, d;
    scale_fn(fn, params[i0], s, ls, d);
    const float shift = params[i1];
    if (z_out != z) z_out[i0] = z[i0];
    if (inverse) {
        z_out[i1] = (z[i1] - shift) / s;
    } else {
        z_out[i1] = s * z[i1] + shift;
        log_s[((long long)b * Ch + c) * Tp + t] = ls;
    }
}

int coupling_fwd(const float* z, const float* params, float* z_out, float* log_s, int B, int C, int Tp, int fn,
                 int inverse, cudaStream_t st) {
    dim3 grid(cdiv(Tp, 256), C / 2, B);
    coupling_fwd_kernel<<<grid, 256, 0, st>>>(z, params, z_out, log_s, C, Tp, fn, inverse);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// backward of the forward coupling.  Incoming gradients are treated as zero beyond each length (the flow loss
// masks them, loss.py:91,102).  Writes dz (B,C,Tp) [first half = pass-through part only] and dparams (B,C,Tp).
__global__ void coupling_bwd_kernel(const float* __restrict__ dz_out, const float* __restrict__ dlog_s,
                                    const float* __restrict__ z, const float* __restrict__ params,
                                    const int* __restrict__ lens, float* __restrict__ dz,
                                    float* __restrict__ dparams, int C, int Tp, int fn) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    const int c = blockIdx.y, b = blockIdx.z, Ch = C / 2;
    const long long i0 = ((long long)b * C + c) * Tp + t, i1 = ((long long)b * C + Ch + c) * Tp + t;
    const bool valid = t < lens[b];
    float s, ls, d;
    scale_fn(fn, params[i0], s, ls, d);
    const float g1 = valid ? dz_out[i1] : 0.0f;
    const float gl = (valid && dlog_s) ? dlog_s[((long long)b * Ch + c) * Tp + t] : 0.0f;
    dz[i0] = valid ? dz_out[i0] : 0.0f;
    dz[i1] = g1 * s;
    float da;
    if (fn == SCALE_EXP) da = g1 * z[i1] * s + gl;
    else if (fn == SCALE_TRANSLATE) da = 0.0f;
    else da = (g1 * z[i1] + gl / s) * d;
    dparams[i0] = da;
    dparams[i1] = g1;
}

int coupling_bwd(const float* dz_out, const float* dlog_s, const float* z, const float* params, const int* lens,
                 float* dz, float* dparams, int B, int C, int Tp, int fn, cudaStream_t st) {
    dim3 grid(cdiv(Tp, 256), C / 2, B);
    coupling_bwd_kernel<<<grid, 256, 0, st>>>(dz_out, dlog_s, z, params, lens, dz, dparams, C, Tp, fn);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// Invertible 1x1 convolution (common.py:540-548, 605-617):  out[b,co,t] = sum_ci W[co,ci]*(in[b,ci,t]-pre[ci]) + post[co]
// A small fp32 GEMM per utterance (M = Cout <= 256, N = T', K = Cin), kept in fp32 FFMA because the flow's invertibility and
// log-determinant ride on it.  HBM-bound in principle (2*C*4 bytes per grouped frame; W, 100 KB, comes from L2), latency-bound
// in practice: the whole problem is 0.5 M outputs, so the tile shape is chosen for the number of resident warps, not for
// register reuse.  One CTA (128 threads) per 32 (co) x 64 (t) tile, a thread owns 4 channels x 4 consecutive frames (two
// 16-byte shared loads per 16 FMAs), K in chunks of 32 through shared memory with the next chunk prefetched into registers.
// At B=8, C=160, T'=400 that is 280 CTAs, all resident at once (two per SM, two warps per scheduler): 32 x 128 tiles made
// 160 CTAs of 8 warps -- 1.08 waves, so the 12 SMs that got a second CTA doubled the kernel's duration (16.6 us).
// History (profiles/r2_ncu_inv1x1.md): a 4 x 8 micro-tile with 128-thread CTAs left ONE warp per scheduler -- 23 us per
// call at B=8, T'=400, issue slots 25 % busy, every shared load's latency exposed (short-scoreboard 2.1 cycles per issue).
// =========================================================================================================
constexpr int INV_TC = 32, INV_TN = 64, INV_TK = 32, INV_THREADS = 128;
__global__ void __launch_bounds__(INV_THREADS) inv1x1_kernel(const float* __restrict__ in, long long in_bs,
                                                     const float* __restrict__ W, const float* __restrict__ pre,
                                                     const float* __restrict__ post, float* __restrict__ out,
                                                     long long out_bs, int Cin, int Cout, int Tp) {
    __shared__ __align__(16) float ws[INV_TK][INV_TC + 4];      // [k][co]
    __shared__ __align__(16) float xs[INV_TK][INV_TN];          // [k][t]
    const int t0 = blockIdx.x * INV_TN, co0 = blockIdx.y * INV_TC, b = blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // 16 (t quads) x 8 (co quads)
    const float* inb = in + (long long)b * in_bs;
    const bool vec_ok = (Tp & 3) == 0 && (in_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    float wreg[8];
    float4 xreg[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {                              // W tile: 32 co x 32 ci, rows of 32 consecutive ci
            const int i = j * INV_THREADS + tid, c = i >> 5, k = i & 31;
            wreg[j] = (co0 + c < Cout && k0 + k < Cin) ? __ldg(W + (long long)(co0 + c) * Cin + k0 + k) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                              // x tile: 32 ci x 64 t, 16-byte pieces
            const int i = j * INV_THREADS + tid, k = i >> 4, q = i & 15;
            const int ci = k0 + k, t = t0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ci < Cin && t < Tp) {
                const float* src = inb + (long long)ci * Tp + t;
                if (vec_ok && t + 3 < Tp) {
                    v = ldg_f4_issue(reinterpret_cast<const float4*>(src));      // stays ahead of the FMAs of the current chunk
                } else {
                    v.x = __ldg(src);
                    if (t + 1 < Tp) v.y = __ldg(src + 1);
                    if (t + 2 < Tp) v.z = __ldg(src + 2);
                    if (t + 3 < Tp) v.w = __ldg(src + 3);
                }
                if (pre) {
                    const float m = __ldg(pre + ci);
                    v.x -= m; v.y -= m; v.z -= m; v.w -= m;      // frames past Tp are never stored
                }
            }
            xreg[j] = v;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < Cin; k0 += INV_TK) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = j * INV_THREADS + tid; ws[i & 31][i >> 5] = wreg[j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int i = j * INV_THREADS + tid; *reinterpret_cast<float4*>(&xs[i >> 4][4 * (i & 15)]) = xreg[j]; }
        __syncthreads();
        if (k0 + INV_TK < Cin) fetch(k0 + INV_TK);
        float4 w4 = *reinterpret_cast<const float4*>(&ws[0][4 * ty]);
        float4 x4 = *reinterpret_cast<const float4*>(&xs[0][4 * tx]);
#pragma unroll
        for (int k = 0; k < INV_TK; ++k) {
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
            if (k + 1 < INV_TK) {                                  // next step's operands are in flight during this step's FMAs
                w4 = *reinterpret_cast<const float4*>(&ws[k + 1][4 * ty]);
                x4 = *reinterpret_cast<const float4*>(&xs[k + 1][4 * tx]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int t = t0 + 4 * tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + 4 * ty + i;
        if (co < Cout && t < Tp) {
            const float pb = post ? __ldg(post + co) : 0.0f;
            float* o = out + (long long)b * out_bs + (long long)co * Tp + t;
            if (t + 3 < Tp && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                reinterpret_cast<float4*>(o)[0] = make_float4(acc[i][0] + pb, acc[i][1] + pb, acc[i][2] + pb, acc[i][3] + pb);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (t + j < Tp) o[j] = acc[i][j] + pb;
            }
        }
    }
}

int inv1x1(const float* in, long long in_bs, const float* W, const float* pre, const float* post, float* out,
           long long out_bs, int B, int Cin, int Cout, int Tp, cudaStream_t st) {
    RADMMM_REQUIRE(Cout >= 1 && Cin >= 1, "inv1x1: channel count out of range (Cin=%d, Cout=%d)", Cin, Cout);
    dim3 grid(cdiv(Tp, INV_TN), cdiv(Cout, INV_TC), B);
    inv1x1_kernel<<<grid, INV_THREADS, 0, st>>>(in, in_bs, W, pre, post, out, out_bs, Cin, Cout, Tp);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// dW[co][ci] += sum over valid (b,t) of dz[b,co,t] * (x[b,ci,t] - pre[ci]).  One CTA per 64 x 64 output tile and
// (utterance, time chunk); a thread owns 4 x 4 outputs (two LDS.128 per 16 FMAs); partial tiles are reduced with fp32
// atomics into the zeroed output.
constexpr int WG_T = 32;
__global__ void __launch_bounds__(256) inv1x1_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ x,
                                                           const float* __restrict__ pre, const int* __restrict__ lens,
                                                           float* __restrict__ dW, int C, int Tp, int t_chunk) {
    __shared__ __align__(16) float a[WG_T][64 + 4], bx[WG_T][64 + 4];        // [t][channel]
    const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
    const int n_tc = cdiv(Tp, t_chunk);
    const int b = blockIdx.z / n_tc, tc = blockIdx.z % n_tc;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // 16 (ci quads) x 16 (co quads)
    const int len = min(lens[b], Tp);
    const int t_begin = tc * t_chunk, t_end = min(len, t_begin + t_chunk);
    if (t_begin >= t_end) return;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int t0 = t_begin; t0 < t_end; t0 += WG_T) {
        for (int i = tid; i < 64 * WG_T; i += 256) {               // rows = channels, 32 consecutive frames each
            const int c = i >> 5, tt = i & 31;
            const int t = t0 + tt;
            const bool ok = t < t_end;
            a[tt][c] = (ok && co0 + c < C) ? dz[((long long)b * C + co0 + c) * Tp + t] : 0.0f;
            bx[tt][c] = (ok && ci0 + c < C) ? x[((long long)b * C + ci0 + c) * Tp + t] - (pre ? pre[ci0 + c] : 0.0f) : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < WG_T; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&a[k][4 * ty]);
            const float4 bv = *reinterpret_cast<const float4*>(&bx[k][4 * tx]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + 4 * tx + j;
            if (co < C && ci < C) atomicAdd(dW + (long long)co * C + ci, acc[i][j]);
        }
    }
}

int inv1x1_wgrad(const float* dz, const float* x, const float* pre, const int* lens, float* dW, int B, int C, int Tp,
                 cudaStream_t st) {
    const int t_chunk = 224;
    RADMMM_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * C * C, st));
    dim3 grid(cdiv(C, 64), cdiv(C, 64), B * cdiv(Tp, t_chunk));
    inv1x1_wgrad_kernel<<<grid, 256, 0, st>>>(dz, x, pre, lens, dW, C, Tp, t_chunk);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// Flow NLL reduction (loss.py:85-110): sums[0] = sum (z*m)^2, sums[1+i] = sum log_s_i * m.  fp64 accumulation,
// warp-shuffle then one atomic per warp.
// =========================================================================================================
__global__ void masked_sum_kernel(const float* __restrict__ x, const int* __restrict__ lens, int C, int Tp,
                                  int square, double* __restrict__ out) {
    const int b = blockIdx.y;
    const int len = min(lens[b], Tp);
    const long long n = (long long)C * Tp;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % Tp);
        if (t < len) {
            const float v = x[(long long)b * n + i];
            acc += square ? (double)v * (double)v : (double)v;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
}

int masked_sum(const float* x, const int* lens, int B, int C, int Tp, int square, double* out, cudaStream_t st) {
    dim3 grid(max(1, min(64, cdiv((long long)C * Tp, 1024))), B);
    masked_sum_kernel<<<grid, 256, 0, st>>>(x, lens, C, Tp, square, out);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// d/dx of  coef * sum (x*m)^2   (square=1: 2*coef*x*m)   or  coef * sum x*m  (square=0: coef*m)
__global__ void masked_sum_bwd_kernel(const float* __restrict__ x, const int* __restrict__ lens, int C, int Tp,
                                      int square, const float* __restrict__ coef_ptr, float coef_mul,
                                      float* __restrict__ dx) {
    const int b = blockIdx.y;
    const int len = min(lens[b], Tp);
    const long long n = (long long)C * Tp;
    const float coef = coef_ptr[0] * coef_mul;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % Tp);
        float v = 0.0f;
        if (t < len) v = square ? 2.0f * coef * x[(long long)b * n + i] : coef;
        dx[(long long)b * n + i] = v;
    }
}

int masked_sum_bwd(const float* x, const int* lens, int B, int C, int Tp, int square, const float* coef_ptr,
                   float coef_mul, float* dx, cudaStream_t st) {
    dim3 grid(max(1, min(256, cdiv((long long)C * Tp, 1024))), B);
    masked_sum_bwd_kernel<<<grid, 256, 0, st>>>(x, lens, C, Tp, square, coef_ptr, coef_mul, dx);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// Weight preparation: weight-norm (w = g * v / ||v||, per output channel) and re-layout into the K-major
// per-tap matrices the contraction kernels read, plus the transposed copies the dgrad GEMMs read.
// =========================================================================================================
// one CTA per output channel: norm[co] = ||v[co]||, rowsum[co] = sum v[co].  16-byte loads, all of a thread's loads in
// flight before the first use; fp32 partials over <= 32 products, then fp64 (a row has up to 5280 elements).
__global__ void __launch_bounds__(256) wn_norm_kernel(const float* __restrict__ v, int per_co, float* __restrict__ norm,
                                                      float* __restrict__ rowsum) {
    const int co = blockIdx.x;
    const float* p = v + (long long)co * per_co;
    double s2 = 0.0, s1 = 0.0;
    if ((per_co & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int n4 = per_co >> 2;
        for (int base = 0; base < n4; base += 256 * 8) {
            float4 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + j * 256 + threadIdx.x;
                t[j] = i < n4 ? __ldg(p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float a2 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a2 = fmaf(t[j].x, t[j].x, a2); a2 = fmaf(t[j].y, t[j].y, a2);
                a2 = fmaf(t[j].z, t[j].z, a2); a2 = fmaf(t[j].w, t[j].w, a2);
                a1 += (t[j].x + t[j].y) + (t[j].z + t[j].w);
            }
            s2 += (double)a2;
            s1 += (double)a1;
        }
    } else {
        for (int i = threadIdx.x; i < per_co; i += blockDim.x) {
            const double x = p[i];
            s2 += x * x;
            s1 += x;
        }
    }
    __shared__ double r2[8], r1[8];
    s2 = warp_sum(s2);
    s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { r2[threadIdx.x >> 5] = s2; r1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += r2[i]; b += r1[i]; }
        norm[co] = (float)sqrt(a);
        if (rowsum) rowsum[co] = (float)b;
    }
}

int wn_norm(const float* v, int n_co, int per_co, float* norm, float* rowsum, cudaStream_t st) {
    wn_norm_kernel<<<n_co, 256, 0, st>>>(v, per_co, norm, rowsum);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// v (co, ci_total, k) fp32; columns [ci_begin, ci_begin+n_ci) go to dst[tap][co][ci - ci_begin] (ld_dst, tap stride)
// and dstT[tap][ci - ci_begin][co].  scale[co] = g/||v|| (or 1 when g == null).
// One CTA per 32 (co) x 64 (ci) block, ALL taps: each output channel contributes one contiguous run of 64*KS floats,
// read once with full-line loads into shared memory; both layouts are then written with 16-byte stores (8 bf16 per
// lane: 128-byte runs along ci for the K-major copy, 64-byte runs along co for the transposed copy).
constexpr int WS_CO = 32, WS_CI = 64;
template <int MODE, int KS>
__global__ void __launch_bounds__(256) wn_scatter_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                         const float* __restrict__ norm, int n_co, int ci_total, int ci_begin,
                                                         int n_ci, ActMat dst, long long dst_tap, ActMat dstT, long long dstT_tap) {
    constexpr int RUN = WS_CI * KS;               // floats per output channel in this block
    constexpr int LDS_ = RUN + 1;                 // +1: conflict-light column reads
    extern __shared__ float tile[];               // [WS_CO][LDS_]
    const int co0 = blockIdx.x * WS_CO, ci0 = blockIdx.y * WS_CI;
    const int tid = threadIdx.x;
    const int ci_n = min(WS_CI, n_ci - ci0);      // valid input channels of this block
    const int run_n = ci_n * KS;
    __shared__ float sc_s[WS_CO];                 // g / ||v|| per output channel of this block
    if (tid < WS_CO) sc_s[tid] = (co0 + tid < n_co) ? (g ? g[co0 + tid] / norm[co0 + tid] : 1.0f) : 0.0f;
    // full tiles whose rows start on a 16-byte boundary: every thread issues ALL of its 16-byte loads before the first use,
    // so a CTA has its whole 40 KB tile in flight at once (the scalar loop below kept ~4 loads per thread in flight and ran
    // at 1.1 TB/s: profiles/r2_launches_bf16.md)
    constexpr int V4 = RUN / 4, NIT = WS_CO * V4 / 256;
    static_assert(RUN % 4 == 0 && (WS_CO * V4) % 256 == 0, "tile must split into whole float4 batches");
    const long long row0 = ((long long)co0 * ci_total + ci_begin + ci0) * KS;
    const bool vec = ci_n == WS_CI && co0 + WS_CO <= n_co && (row0 & 3) == 0 && (((long long)ci_total * KS) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(v) & 15) == 0;
    if (vec) {
        float4 buf[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = it * 256 + tid, c = i / V4, j4 = i - c * V4;
            buf[it] = ldg_f4_issue(reinterpret_cast<const float4*>(v + row0 + (long long)c * ci_total * KS) + j4);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int i = it * 256 + tid, c = i / V4, j4 = i - c * V4;
            const float sc = sc_s[c];
            float* o = tile + c * LDS_ + 4 * j4;
            o[0] = sc * buf[it].x; o[1] = sc * buf[it].y; o[2] = sc * buf[it].z; o[3] = sc * buf[it].w;
        }
    } else {
        __syncthreads();
        for (int i = tid; i < WS_CO * RUN; i += 256) {
            const int c = i / RUN, j = i - c * RUN;
            float val = 0.0f;
            if (co0 + c < n_co && j < run_n) val = sc_s[c] * __ldg(v + ((long long)(co0 + c) * ci_total + ci_begin + ci0) * KS + j);
            tile[c * LDS_ + j] = val;
        }
    }
    __syncthreads();
    constexpr int P = (MODE == MODE_BF16X3) ? 2 : 1;
    // K-major copy: (tap, co) rows of 64 ci; a thread packs 8 consecutive ci
    if (dst.ptr != nullptr) {
        for (int i = tid; i < KS * WS_CO * (WS_CI / 8); i += 256) {
            const int seg = i % (WS_CI / 8), c = (i / (WS_CI / 8)) % WS_CO, tap = i / ((WS_CI / 8) * WS_CO);
            const int ci = seg * 8;
            if (co0 + c >= n_co || ci >= ci_n) continue;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = (ci + e < ci_n) ? tile[c * LDS_ + (ci + e) * KS + tap] : 0.0f;
            const long long off = tap * dst_tap + (long long)(co0 + c) * dst.ld + ci0 + ci;
            if constexpr (MODE == MODE_F32) {
                float* o = reinterpret_cast<float*>(dst.ptr) + off;
                if (ci + 8 <= ci_n && ((dst.ld | (ci0 + ci)) & 3) == 0) {
                    reinterpret_cast<float4*>(o)[0] = make_float4(x[0], x[1], x[2], x[3]);
                    reinterpret_cast<float4*>(o)[1] = make_float4(x[4], x[5], x[6], x[7]);
                } else {
                    for (int e = 0; e < 8 && ci + e < ci_n; ++e) o[e] = x[e];
                }
            } else {
#pragma unroll
                for (int pl = 0; pl < P; ++pl) {
                    __nv_bfloat162 h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        h[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
                        if (P == 2 && pl == 0) { x[2 * e] -= __bfloat162float(h[e].x); x[2 * e + 1] -= __bfloat162float(h[e].y); }
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst.ptr) + pl * dst.plane_stride + off;
                    if (ci + 8 <= ci_n && ((dst.ld | (ci0 + ci)) & 7) == 0) {
                        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
                    } else {
                        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(h);
                        for (int e = 0; e < 8 && ci + e < ci_n; ++e) o[e] = hs[e];
                    }
                }
            }
        }
    }
    // transposed copy: (tap, ci) rows of 32 co; a thread packs 8 consecutive co
    if (dstT.ptr != nullptr) {
        for (int i = tid; i < KS * WS_CI * (WS_CO / 8); i += 256) {
            const int seg = i % (WS_CO / 8), ci = (i / (WS_CO / 8)) % WS_CI, tap = i / ((WS_CO / 8) * WS_CI);
            const int c = seg * 8;
            if (ci >= ci_n || co0 + c >= n_co) continue;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = (co0 + c + e < n_co) ? tile[(c + e) * LDS_ + ci * KS + tap] : 0.0f;
            const long long off = tap * dstT_tap + (long long)(ci0 + ci) * dstT.ld + co0 + c;
            const bool full = co0 + c + 8 <= n_co;
            if constexpr (MODE == MODE_F32) {
                float* o = reinterpret_cast<float*>(dstT.ptr) + off;
                if (full && ((dstT.ld | (co0 + c)) & 3) == 0) {
                    reinterpret_cast<float4*>(o)[0] = make_float4(x[0], x[1], x[2], x[3]);
                    reinterpret_cast<float4*>(o)[1] = make_float4(x[4], x[5], x[6], x[7]);
                } else {
                    for (int e = 0; e < 8 && co0 + c + e < n_co; ++e) o[e] = x[e];
                }
            } else {
#pragma unroll
                for (int pl = 0; pl < P; ++pl) {
                    __nv_bfloat162 h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        h[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
                        if (P == 2 && pl == 0) { x[2 * e] -= __bfloat162float(h[e].x); x[2 * e + 1] -= __bfloat162float(h[e].y); }
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dstT.ptr) + pl * dstT.plane_stride + off;
                    if (full && ((dstT.ld | (co0 + c)) & 7) == 0) {
                        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
                    } else {
                        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(h);
                        for (int e = 0; e < 8 && co0 + c + e < n_co; ++e) o[e] = hs[e];
                    }
                }
            }
        }
    }
}

template <int MODE, int KS>
static int wn_scatter_launch(const float* v, const float* g, const float* norm, int n_co, int ci_total, int ci_begin,
                             int n_ci, ActMat dst, long long dst_tap, ActMat dstT, long long dstT_tap, cudaStream_t st) {
    const size_t smem = sizeof(float) * WS_CO * (WS_CI * KS + 1);
    auto kern = wn_scatter_kernel<MODE, KS>;
    if (smem > 48 * 1024) {
        static bool set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!set[dev & 63]) {
            RADMMM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set[dev & 63] = true;
        }
    }
    dim3 grid(cdiv(n_co, WS_CO), cdiv(n_ci, WS_CI));
    kern<<<grid, 256, smem, st>>>(v, g, norm, n_co, ci_total, ci_begin, n_ci, dst, dst_tap, dstT, dstT_tap);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int wn_scatter(int mode, const float* v, const float* g, const float* norm, int n_co, int ci_total, int ksize,
               int ci_begin, int n_ci, ActMat dst, long long dst_tap, ActMat dstT, long long dstT_tap, cudaStream_t st) {
    RADMMM_REQUIRE(ksize == 1 || ksize == 5, "wn_scatter: kernel size %d (1 or 5)", ksize);
#define RADMMM_WS(M)                                                                                                   \
    return ksize == 5 ? wn_scatter_launch<M, 5>(v, g, norm, n_co, ci_total, ci_begin, n_ci, dst, dst_tap, dstT, dstT_tap, st) \
                      : wn_scatter_launch<M, 1>(v, g, norm, n_co, ci_total, ci_begin, n_ci, dst, dst_tap, dstT, dstT_tap, st)
    if (mode == MODE_F32) { RADMMM_WS(MODE_F32); }
    if (mode == MODE_BF16) { RADMMM_WS(MODE_BF16); }
    RADMMM_WS(MODE_BF16X3);
#undef RADMMM_WS
}

// padq[co] = log(2) * (g/||v||) * rowsum(v[co]) + bias[co]: the res-skip pre-activation on frames beyond the
// sequence length, where the reference feeds softplus(0) = log 2 on every channel (common.py:830-831).
__global__ void padq_kernel(const float* g, const float* norm, const float* rowsum, const float* bias, float* padq, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) padq[i] = 0.6931471805599453f * (g[i] / norm[i]) * rowsum[i] + bias[i];
}

int padq_compute(const float* g, const float* norm, const float* rowsum, const float* bias, float* padq, int n,
                 cudaStream_t st) {
    padq_kernel<<<cdiv(n, 256), 256, 0, st>>>(g, norm, rowsum, bias, padq, n);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// backward of weight norm.  dW comes from the weight-grad GEMM as up to two column blocks of [tap][co][ld]:
//   ci in [0, n_ci0) from src0, ci in [n_ci0, ci_total) from src1.   One CTA per output channel.
//   dg[co] = <dW, v>/||v|| ;  dv = (g/||v||) * (dW - v * <dW,v>/||v||^2)
__global__ void wn_bwd_kernel(const float* __restrict__ src0, long long ld0, long long tap0, int n_ci0,
                              const float* __restrict__ src1, long long ld1, long long tap1,
                              const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ norm,
                              int ci_total, int ksize, float* __restrict__ dv, float* __restrict__ dg) {
    const int co = blockIdx.x;
    const int per_co = ci_total * ksize;
    const float* vp = v + (long long)co * per_co;
    auto dw_at = [&](int ci, int k) -> float {
        return ci < n_ci0 ? src0[k * tap0 + (long long)co * ld0 + ci] : src1[k * tap1 + (long long)co * ld1 + (ci - n_ci0)];
    };
    double dot = 0.0;
    for (int i = threadIdx.x; i < per_co; i += blockDim.x) dot += (double)dw_at(i / ksize, i % ksize) * (double)vp[i];
    __shared__ double red[8];
    __shared__ float sdot;
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += red[i];
        sdot = (float)a;
    }
    __syncthreads();
    const float nrm = norm[co], gg = g[co];
    const float d = sdot;
    if (threadIdx.x == 0) dg[co] = d / nrm;
    const float sc = gg / nrm, coef = d / (nrm * nrm);
    float* o = dv + (long long)co * per_co;
    for (int i = threadIdx.x; i < per_co; i += blockDim.x) o[i] = sc * (dw_at(i / ksize, i % ksize) - vp[i] * coef);
}

// Same computation with the output channel's v row and its dW entries staged in shared memory: every global access is
// coalesced (dW is read tap plane by tap plane, v and dv as contiguous rows) and each byte moves exactly once
// (algorithmic traffic 3 x 4 bytes per weight; the generic kernel above gathers dW across the tap planes twice).
__global__ void __launch_bounds__(256) wn_bwd_staged_kernel(const float* __restrict__ src0, long long ld0, long long tap0,
                                                            int n_ci0, const float* __restrict__ src1, long long ld1,
                                                            long long tap1, const float* __restrict__ v,
                                                            const float* __restrict__ g, const float* __restrict__ norm,
                                                            int ci_total, int ksize, float* __restrict__ dv,
                                                            float* __restrict__ dg) {
    extern __shared__ float wsm[];
    const int co = blockIdx.x;
    const int per_co = ci_total * ksize;
    float* sv = wsm;                 // [ci][k]  (the layout of v)
    float* sd = wsm + per_co;        // [ci][k]
    const float* vp = v + (long long)co * per_co;
    // loads are issued in batches of 8 per thread before the first store: enough bytes in flight to cover the DRAM latency
    if ((per_co & 3) == 0 && (reinterpret_cast<uintptr_t>(vp) & 15) == 0) {
        const float4* v4 = reinterpret_cast<const float4*>(vp);
        const int n4 = per_co >> 2;
        for (int base = 0; base < n4; base += 256 * 4) {
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = base + j * 256 + threadIdx.x; if (i < n4) t[j] = __ldg(v4 + i); }
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int i = base + j * 256 + threadIdx.x; if (i < n4) reinterpret_cast<float4*>(sv)[i] = t[j]; }
        }
    } else {
        for (int base = 0; base < per_co; base += 256 * 8) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int i = base + j * 256 + threadIdx.x; if (i < per_co) t[j] = __ldg(vp + i); }
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int i = base + j * 256 + threadIdx.x; if (i < per_co) sv[i] = t[j]; }
        }
    }
    for (int k0 = 0; k0 < ksize; k0 += 5) {                       // all taps of a chunk of input channels in flight together
        for (int c0 = 0; c0 < ci_total; c0 += 256 * 2) {
            float t[5][2];
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
                const int k = k0 + kk;
                if (k >= ksize) break;
                const float* p0 = src0 + k * tap0 + (long long)co * ld0;
                const float* p1 = src1 ? src1 + k * tap1 + (long long)co * ld1 : nullptr;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ci = c0 + j * 256 + threadIdx.x;
                    if (ci < ci_total) t[kk][j] = ci < n_ci0 ? __ldg(p0 + ci) : __ldg(p1 + ci - n_ci0);
                }
            }
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
                const int k = k0 + kk;
                if (k >= ksize) break;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ci = c0 + j * 256 + threadIdx.x;
                    if (ci < ci_total) sd[ci * ksize + k] = t[kk][j];
                }
            }
        }
    }
    __syncthreads();
    double dot = 0.0;
    {
        float part = 0.0f;               // <= 8 products per fp32 partial, then fp64
        int n = 0;
        for (int i = threadIdx.x; i < per_co; i += 256) {
            part = fmaf(sd[i], sv[i], part);
            if (++n == 8) { dot += (double)part; part = 0.0f; n = 0; }
        }
        dot += (double)part;
    }
    __shared__ double red[8];
    __shared__ float sdot;
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += red[i];
        sdot = (float)a;
    }
    __syncthreads();
    const float nrm = norm[co], gg = g[co];
    const float d = sdot;
    if (threadIdx.x == 0) dg[co] = d / nrm;
    const float sc = gg / nrm, coef = d / (nrm * nrm);
    float* o = dv + (long long)co * per_co;
    if ((per_co & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
        for (int i = threadIdx.x; i < (per_co >> 2); i += 256) {
            const float4 a = reinterpret_cast<const float4*>(sd)[i], b = reinterpret_cast<const float4*>(sv)[i];
            reinterpret_cast<float4*>(o)[i] = make_float4(sc * (a.x - b.x * coef), sc * (a.y - b.y * coef),
                                                          sc * (a.z - b.z * coef), sc * (a.w - b.w * coef));
        }
    } else {
        for (int i = threadIdx.x; i < per_co; i += 256) o[i] = sc * (sd[i] - sv[i] * coef);
    }
}

// The same computation fed by the bulk-copy engine: the v row (contiguous) and the ksize dW rows of the output channel are
// brought into shared memory by cp.async.bulk (one elected thread, one mbarrier) -- no register staging, all of a CTA's
// bytes in flight at once, several CTAs per SM overlapping their load / compute / store phases.  dW stays in its
// [tap][ci] layout; the dot product and dv read it transposed out of shared memory (stride ksize: conflict-free for 5).
// Needs 16-byte aligned rows (ci_total % 4 == 0 and a single source block); other shapes use the staged kernel above.
// KS (taps) is a template parameter so that i -> (ci, k) is a multiplication, and the dW rows ([k][ci], as the weight-grad GEMM
// writes them) are staged with a row pitch of ci_total + 8 floats: with the natural pitch (a multiple of 32 banks) the KS
// consecutive threads that share one ci hit ONE bank in both passes -- a 5-way conflict that made this kernel LSU-bound at
// ~2.7 TB/s; the padded pitch keeps the bulk copies 16-byte aligned and spreads them over 8k + ci banks.
constexpr int WNB_PAD = 8;
template <int KS>
__global__ void __launch_bounds__(256) wn_bwd_bulk_kernel(const float* __restrict__ src, long long ld, long long tap_stride,
                                                          const float* __restrict__ v, const float* __restrict__ g,
                                                          const float* __restrict__ norm, int ci_total,
                                                          float* __restrict__ dv, float* __restrict__ dg) {
    extern __shared__ __align__(128) float wsm[];
    const int co = blockIdx.x;
    const int per_co = ci_total * KS;
    const int pitch = ci_total + WNB_PAD;
    float* sv = wsm;                 // [ci][k]  (the layout of v)
    float* sd = wsm + per_co;        // [k][pitch]  (the layout of dW, padded rows)
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes_v = (unsigned)per_co * 4u, bytes_row = (unsigned)ci_total * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes_v + bytes_row * KS) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(sv)), "l"(v + (long long)co * per_co), "r"(bytes_v), "r"(bar_addr) : "memory");
        for (int k = 0; k < KS; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((unsigned)__cvta_generic_to_shared(sd + k * pitch)), "l"(src + k * tap_stride + (long long)co * ld),
                           "r"(bytes_row), "r"(bar_addr) : "memory");
    }
    __syncthreads();                                   // the barrier is initialised before anyone polls it
    {
        unsigned done = 0;
        long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_addr) : "memory");
            if (++spins > (1ll << 22)) { printf("radmmm wn_bwd_bulk: copy timed out (block %d)\n", blockIdx.x); __trap(); }
        }
    }
    double dot = 0.0;
    {
        float part = 0.0f;               // <= 8 products per fp32 partial, then fp64
        int n = 0;
        for (int i = threadIdx.x; i < per_co; i += 256) {
            const int ci = i / KS, k = i - ci * KS;
            part = fmaf(sd[k * pitch + ci], sv[i], part);
            if (++n == 8) { dot += (double)part; part = 0.0f; n = 0; }
        }
        dot += (double)part;
    }
    __shared__ double red[8];
    __shared__ float sdot;
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < 8; ++i) a += red[i];
        sdot = (float)a;
    }
    __syncthreads();
    const float nrm = norm[co], gg = g[co];
    const float d = sdot;
    if (threadIdx.x == 0) dg[co] = d / nrm;
    const float sc = gg / nrm, coef = d / (nrm * nrm);
    float4* o = reinterpret_cast<float4*>(dv + (long long)co * per_co);
    for (int i4 = threadIdx.x; i4 < (per_co >> 2); i4 += 256) {
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e, ci = i / KS, k = i - ci * KS;
            r[e] = sc * (sd[k * pitch + ci] - sv[i] * coef);
        }
        o[i4] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

int wn_bwd(const float* src0, long long ld0, long long tap0, int n_ci0, const float* src1, long long ld1,
           long long tap1, const float* v, const float* g, const float* norm, int n_co, int ci_total, int ksize,
           float* dv, float* dg, cudaStream_t st) {
    const size_t smem = sizeof(float) * 2 * (size_t)ci_total * ksize;
    static const bool bulk_on = []() { const char* e = getenv("RADMMM_B200_WNBWD_BULK"); return !(e && e[0] == '0'); }();
    const bool aligned = src1 == nullptr && n_ci0 == ci_total && (ci_total & 3) == 0 && (ld0 & 3) == 0 && (tap0 & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0;
    const size_t smem_bulk = smem + sizeof(float) * (size_t)ksize * WNB_PAD;
    if (bulk_on && aligned && smem_bulk <= 48 * 1024 && (ksize == 5 || ksize == 1)) {
        if (ksize == 5) wn_bwd_bulk_kernel<5><<<n_co, 256, smem_bulk, st>>>(src0, ld0, tap0, v, g, norm, ci_total, dv, dg);
        else wn_bwd_bulk_kernel<1><<<n_co, 256, smem_bulk, st>>>(src0, ld0, tap0, v, g, norm, ci_total, dv, dg);
    } else if (smem <= 48 * 1024) {
        wn_bwd_staged_kernel<<<n_co, 256, smem, st>>>(src0, ld0, tap0, n_ci0, src1, ld1, tap1, v, g, norm, ci_total, ksize, dv, dg);
    } else {
        wn_bwd_kernel<<<n_co, 256, 0, st>>>(src0, ld0, tap0, n_ci0, src1, ld1, tap1, v, g, norm, ci_total, ksize, dv, dg);
    }
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

// =========================================================================================================
// Bias gradients: out[n] = sum over valid rows of X[r][n] * (unratio ? 1/ratio(r) : 1).  X is an act matrix.
// =========================================================================================================
// block (32, 8): x = 8 consecutive columns per thread (one 16-byte load in the bf16 modes), y = row lane; each block
// reduces 64 rows x 256 columns through shared memory and issues one atomic per column.
template <int MODE>
__global__ void __launch_bounds__(256) colsum_kernel(ActMat x, RowGeom g, int n_cols, int dilation, int unratio,
                                                     float* __restrict__ out) {
    __shared__ float red[8][256 + 8];
    const int c0 = blockIdx.x * 256 + threadIdx.x * 8;
    const int r_begin = blockIdx.y * 64;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < n_cols) {
        for (int r = r_begin + threadIdx.y; r < min(g.R, r_begin + 64); r += 8) {
            int b, t, len;
            row_decode(g, r, b, t, len);
            if (t >= len) continue;
            const float w = unratio ? 1.0f / pconv_ratio(t, len, dilation) : 1.0f;
            const long long idx = (long long)r * x.ld + c0;
            float v[8];
            if constexpr (MODE == MODE_F32) {
                const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x.ptr) + idx);
                const float4 c = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x.ptr) + idx + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
            } else {
                const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
                const uint4 hv = *reinterpret_cast<const uint4*>(base + idx);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h2[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
                if constexpr (MODE == MODE_BF16X3) {
                    const uint4 lv = *reinterpret_cast<const uint4*>(base + idx + x.plane_stride);
                    const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&lv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(l2[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], w, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][tid];
    const int c = blockIdx.x * 256 + tid;
    if (c < n_cols && tot != 0.0f) atomicAdd(out + c, tot);
}

__global__ void zero_list_kernel(ZeroList z) {
    float* p = z.ptr[blockIdx.x];
    for (int i = threadIdx.x; i < z.count[blockIdx.x]; i += blockDim.x) p[i] = 0.0f;
}
int zero_list(const ZeroList& z, cudaStream_t st) {
    if (z.n <= 0) return RADMMM_OK;
    zero_list_kernel<<<z.n, 256, 0, st>>>(z);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int colsum(int mode, ActMat x, const RowGeom& g, int n_cols, int dilation, int unratio, float* out, cudaStream_t st) {
    RADMMM_REQUIRE(x.ld % 8 == 0, "colsum: row pitch must be a multiple of 8");
    RADMMM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_cols, st));
    dim3 grid(cdiv(n_cols, 256), cdiv(g.R, 64)), block(32, 8);
    if (mode == MODE_F32) colsum_kernel<MODE_F32><<<grid, block, 0, st>>>(x, g, n_cols, dilation, unratio, out);
    else if (mode == MODE_BF16) colsum_kernel<MODE_BF16><<<grid, block, 0, st>>>(x, g, n_cols, dilation, unratio, out);
    else colsum_kernel<MODE_BF16X3><<<grid, block, 0, st>>>(x, g, n_cols, dilation, unratio, out);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
