// extern "C" surface of libradmmm_b200.so (see include/radmmm_b200.h).
#include <stdarg.h>
#include "../../include/radmmm_b200.h"
#include "gemm.cuh"
#include "ops.cuh"

namespace radmmm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1, __ATOMIC_RELAXED); }

// ---- optional per-launch profiler (bench.py): CUDA events around every contraction launch, on the launching stream.
// Off by default; when off the launch path does not touch it.
struct ProfSlot { cudaEvent_t e0, e1; int tag; double flops; };
static const int kProfMax = 4096;
static ProfSlot g_prof[kProfMax];
static int g_prof_n = 0;
static bool g_prof_on = false;

int launch_gemm(const GemmArgs& args, int mode, cudaStream_t stream) {
    ProfSlot* slot = nullptr;
    if (g_prof_on && g_prof_n < kProfMax) {
        slot = &g_prof[g_prof_n++];
        if (!slot->e0) { cudaEventCreate(&slot->e0); cudaEventCreate(&slot->e1); }
        slot->tag = args.wgrad == 3 ? 24 + args.n_seg : args.wgrad ? 16 + args.n_seg : args.epi.kind;
        double k = 0;
        for (int s = 0; s < args.n_seg; ++s) k += args.seg[s].K;
        slot->flops = args.wgrad ? 2.0 * args.epi.M * args.epi.N * (double)args.R * args.n_seg
                                 : 2.0 * args.R * (double)args.epi.N * k;
        cudaEventRecord(slot->e0, stream);
    }
    int rc = (mode == MODE_F32) ? launch_gemm_ffma(args, stream) : launch_gemm_tc(args, mode, stream);
    if (slot) cudaEventRecord(slot->e1, stream);
    return rc;
}

void lstm_cluster_set_trace(void* buf);      // lstm_cluster.cu (diagnostic phase timers)
// wn.cu
int flow_prepare(const radmmm_flow_desc* f, cudaStream_t st);
int flow_forward(const radmmm_flow_desc* f, const float* z_in, float* z_mid, float* params, float* z_out, float* log_s,
                 cudaStream_t st);
int flow_inverse(const radmmm_flow_desc* f, const float* z_in, float* params, float* z_tmp, float* z_out, cudaStream_t st);
int flow_backward(const radmmm_flow_desc* f, const float* z_in, const float* z_mid, const float* params,
                  const float* dz_out, const float* dlog_s, float* dz_mid, float* dparams, float* dz_in,
                  float* dctx_rows, const radmmm_flow_grads* gr, void* scratch, cudaStream_t st);
size_t flow_prepared_bytes(int mode, int C, int D, int H, int L);
size_t flow_workspace_bytes(int mode, int training, int B, int Tp, int C, int D, int H, int L);
size_t flow_scratch_bytes(int mode, int B, int Tp, int C, int D, int H, int L);
size_t context_rows_bytes(int mode, int B, int Tp, int D);
int context_rows(int mode, const float* ctx_btd, const int* lens, int B, int Tp, int D, void* rows, cudaStream_t st);
int context_rows_backward(const float* drows, const int* lens, int B, int Tp, int D, float* dctx, int accumulate, cudaStream_t st);

template <int MODE>
__global__ void cast_rows_kernel(const float* __restrict__ src, long long n, ActMat dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        act_store<MODE>(dst, i, src[i]);
}

int cast_rows(int mode, const float* src, long long n, void* dst, long long plane_stride, cudaStream_t st) {
    ActMat m; m.ptr = dst; m.ld = 0; m.plane_stride = plane_stride;
    int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    if (grid < 1) grid = 1;
    if (mode == MODE_F32) cast_rows_kernel<MODE_F32><<<grid, 256, 0, st>>>(src, n, m);
    else if (mode == MODE_BF16) cast_rows_kernel<MODE_BF16><<<grid, 256, 0, st>>>(src, n, m);
    else cast_rows_kernel<MODE_BF16X3><<<grid, 256, 0, st>>>(src, n, m);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm

using namespace radmmm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int radmmm_abi_version(void) { return RADMMM_ABI_VERSION; }
long long radmmm_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

void radmmm_debug_trace(void* device_buf, int max_ctas, int max_launches) {
    if (max_launches < 0) lstm_cluster_set_trace(device_buf);      // max_launches < 0: the buffer goes to the cluster LSTM kernels
    else gemm_tc_set_trace(device_buf, max_ctas, max_launches);
}
void radmmm_profile_enable(int on) { g_prof_on = on != 0; if (on) g_prof_n = 0; }
int radmmm_profile_collect(int max_tags, int* counts, double* ms, double* flops) {
    cudaDeviceSynchronize();
    for (int i = 0; i < max_tags; ++i) { counts[i] = 0; ms[i] = 0; flops[i] = 0; }
    for (int i = 0; i < g_prof_n; ++i) {
        float t = 0;
        if (cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1) != cudaSuccess) continue;
        const int tag = g_prof[i].tag;
        if (tag < 0 || tag >= max_tags) continue;
        counts[tag] += 1; ms[tag] += t; flops[tag] += g_prof[i].flops;
    }
    const int n = g_prof_n;
    g_prof_n = 0;
    return n;
}
const char* radmmm_last_error(void) { return last_error(); }
size_t radmmm_sizeof_flow_desc(void) { return sizeof(radmmm_flow_desc); }
size_t radmmm_sizeof_flow_grads(void) { return sizeof(radmmm_flow_grads); }
int radmmm_pitch(int Tp) { return Tp + RADMMM_ROW_GAP; }
int radmmm_rows(int B, int Tp) { return (int)round_up((long long)B * (Tp + RADMMM_ROW_GAP), 256); }

size_t radmmm_flow_prepared_bytes(int mode, int C, int D, int H, int L) { return flow_prepared_bytes(mode, C, D, H, L); }
size_t radmmm_flow_workspace_bytes(int mode, int training, int B, int Tp, int C, int D, int H, int L) {
    return flow_workspace_bytes(mode, training, B, Tp, C, D, H, L);
}
size_t radmmm_flow_backward_scratch_bytes(int mode, int B, int Tp, int C, int D, int H, int L) {
    return flow_scratch_bytes(mode, B, Tp, C, D, H, L);
}
size_t radmmm_context_rows_bytes(int mode, int B, int Tp, int D) { return context_rows_bytes(mode, B, Tp, D); }
int radmmm_flow_prepare(const radmmm_flow_desc* d, void* stream) { return flow_prepare(d, ST(stream)); }
int radmmm_context_rows(int mode, const float* ctx_btd, const int32_t* lens, int B, int Tp, int D, void* rows,
                        void* stream) {
    return context_rows(mode, ctx_btd, lens, B, Tp, D, rows, ST(stream));
}
int radmmm_context_rows_backward(const float* drows, const int32_t* lens, int B, int Tp, int D, float* dctx_btd,
                                 int accumulate, void* stream) {
    return context_rows_backward(drows, lens, B, Tp, D, dctx_btd, accumulate, ST(stream));
}
int radmmm_flow_forward(const radmmm_flow_desc* d, const float* z_in, float* z_mid, float* params, float* z_out,
                        float* log_s, void* stream) {
    return flow_forward(d, z_in, z_mid, params, z_out, log_s, ST(stream));
}
int radmmm_flow_inverse(const radmmm_flow_desc* d, const float* z_in, float* params, float* z_tmp, float* z_out,
                        void* stream) {
    return flow_inverse(d, z_in, params, z_tmp, z_out, ST(stream));
}
int radmmm_flow_backward(const radmmm_flow_desc* d, const float* z_in, const float* z_mid, const float* params,
                         const float* dz_out, const float* dlog_s, float* dz_mid, float* dparams, float* dz_in,
                         float* dctx_rows, const radmmm_flow_grads* grads, void* scratch, void* stream) {
    return flow_backward(d, z_in, z_mid, params, dz_out, dlog_s, dz_mid, dparams, dz_in, dctx_rows, grads, scratch,
                         ST(stream));
}

int radmmm_radam_chunk_elems(void) { return radam_chunk_elems(); }
int radmmm_radam_step(const void* recs, const int32_t* chunk_tensor, const long long* chunk_off, int n_chunks, double* state,
                      const double* cfg, void* stream) {
    return radam_step(recs, chunk_tensor, chunk_off, n_chunks, state, cfg, ST(stream));
}

size_t radmmm_lstm_workspace_bytes(int B, int H) { return lstm_workspace_bytes(B, H); }
int radmmm_lstm_forward(int mode, const float* xproj, const float* whh_f, const float* whh_r, const int32_t* lens, int B, int Tp,
                        int H, float* out, float* gates, float* cstate, void* workspace, void* stream) {
    return lstm_forward(mode, xproj, whh_f, whh_r, lens, B, Tp, H, out, gates, cstate, workspace, ST(stream));
}
int radmmm_lstm_backward(int mode, const float* dout, const float* gates, const float* cstate, const float* whh_f,
                         const float* whh_r, const int32_t* lens, int B, int Tp, int H, float* dgates, void* workspace,
                         void* stream) {
    return lstm_backward(mode, dout, gates, cstate, whh_f, whh_r, lens, B, Tp, H, dgates, workspace, ST(stream));
}

int radmmm_inv1x1(const float* in, const float* W, const float* pre, const float* post, float* out, int B, int Cin,
                  int Cout, int Tp, void* stream) {
    return inv1x1(in, (long long)Cin * Tp, W, pre, post, out, (long long)Cout * Tp, B, Cin, Cout, Tp, ST(stream));
}
int radmmm_inv1x1_wgrad(const float* dz, const float* x, const float* pre, const int32_t* lens, float* dW, int B, int C,
                        int Tp, void* stream) {
    return inv1x1_wgrad(dz, x, pre, lens, dW, B, C, Tp, ST(stream));
}
int radmmm_coupling_forward(const float* z, const float* params, float* z_out, float* log_s, int B, int C, int Tp,
                            int scaling_fn, int inverse, void* stream) {
    return coupling_fwd(z, params, z_out, log_s, B, C, Tp, scaling_fn, inverse, ST(stream));
}
int radmmm_coupling_backward(const float* dz_out, const float* dlog_s, const float* z, const float* params,
                             const int32_t* lens, float* dz, float* dparams, int B, int C, int Tp, int scaling_fn,
                             void* stream) {
    return coupling_bwd(dz_out, dlog_s, z, params, lens, dz, dparams, B, C, Tp, scaling_fn, ST(stream));
}
int radmmm_masked_sum(const float* x, const int32_t* lens, int B, int C, int Tp, int square, double* out, void* stream) {
    return masked_sum(x, lens, B, C, Tp, square, out, ST(stream));
}
int radmmm_masked_sum_backward(const float* x, const int32_t* lens, int B, int C, int Tp, int square, const float* coef,
                               float coef_mul, float* dx, void* stream) {
    return masked_sum_bwd(x, lens, B, C, Tp, square, coef, coef_mul, dx, ST(stream));
}

int radmmm_conv_rows(int mode, const void* x_rows, long long x_ld, long long x_plane, const void* w, long long w_ld,
                     long long w_plane, long long w_tap_stride, const float* bias, float* y, long long y_ld, int R,
                     int K, int N, int taps, int dilation, void* stream) {
    RADMMM_REQUIRE(mode >= 0 && mode <= 2, "conv_rows: bad mode %d", mode);
    RADMMM_REQUIRE(taps >= 1 && taps <= kMaxSeg && taps % 2 == 1, "conv_rows: taps=%d must be odd and <= %d", taps, kMaxSeg);
    RADMMM_REQUIRE(R % 128 == 0 && K % 64 == 0, "conv_rows: R=%d must be a multiple of 128 and K=%d of 64", R, K);
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R;
    a.epi.kind = EPI_F32;
    a.epi.N = N;
    a.epi.bias = bias;
    a.epi.f32_out = y;
    a.epi.f32_ld = y_ld;
    static const int one_len = 0;
    a.epi.geom.lens = nullptr; a.epi.geom.B = 0; a.epi.geom.Tp = R; a.epi.geom.pitch = R; a.epi.geom.R = R;
    (void)one_len;
    const size_t es = mode_elem_bytes(mode);
    for (int j = 0; j < taps; ++j) {
        GemmSeg& s = a.seg[a.n_seg++];
        s.a.ptr = const_cast<void*>(x_rows); s.a.ld = x_ld; s.a.plane_stride = x_plane;
        s.w.ptr = (char*)const_cast<void*>(w) + (size_t)j * w_tap_stride * es; s.w.ld = w_ld; s.w.plane_stride = w_plane;
        s.K = K;
        s.shift = (j - taps / 2) * dilation;
    }
    return launch_gemm(a, mode, ST(stream));
}
int radmmm_wgrad_rows(int mode, const void* dy, long long dy_ld, long long dy_plane, const void* x,
                      long long x_ld, long long x_plane, float* out, long long out_ld,
                      long long out_tap_stride, int R, int M, int N, int taps, int dilation, int shift_offset,
                      void* stream) {
    RADMMM_REQUIRE(mode >= 0 && mode <= 2, "wgrad_rows: bad mode %d", mode);
    RADMMM_REQUIRE(taps >= 1 && taps <= kMaxSeg && taps % 2 == 1, "wgrad_rows: taps=%d must be odd and <= %d", taps, kMaxSeg);
    RADMMM_REQUIRE(R % 128 == 0, "wgrad_rows: R=%d must be a multiple of 128", R);
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.R = R;
    a.wgrad = 1;
    a.split_k = 0;
    a.epi.kind = EPI_WGRAD;
    a.epi.N = N;
    a.epi.M = M;
    a.epi.f32_out = out;
    a.epi.f32_ld = out_ld;
    a.epi.f32_tap_stride = out_tap_stride;
    a.epi.atomic = 1;
    for (int j = 0; j < taps; ++j) {
        GemmSeg& s = a.seg[a.n_seg++];
        s.a.ptr = const_cast<void*>(dy); s.a.ld = dy_ld; s.a.plane_stride = dy_plane;
        s.w.ptr = const_cast<void*>(x); s.w.ld = x_ld; s.w.plane_stride = x_plane;
        s.K = R;
        s.shift = (j - taps / 2) * dilation + shift_offset;
    }
    a.zero_output = 1;          // the launcher zero-fills `out` itself when its tiling reduces with red.add (same as the train step)
    return launch_gemm(a, mode, ST(stream));
}
int radmmm_cast_rows(int mode, const float* src, long long n, void* dst, long long plane_stride, void* stream) {
    return cast_rows(mode, src, n, dst, plane_stride, ST(stream));
}

int radmmm_spline_forward(const float* z1, const float* q, const int32_t* lens, float* z1_out, float* log_s, int B,
                          int Ch, int Tp, int n_bins, float lo, float hi, int inverse, void* stream) {
    return spline_fwd(z1, q, lens, z1_out, log_s, B, Ch, Tp, n_bins, lo, hi, inverse, ST(stream));
}
int radmmm_spline_backward(const float* z1, const float* q, const int32_t* lens, const float* dz1_out,
                           const float* dlog_s, float* dz1, float* dq, int B, int Ch, int Tp, int n_bins, float lo,
                           float hi, void* stream) {
    return spline_bwd(z1, q, lens, dz1_out, dlog_s, dz1, dq, B, Ch, Tp, n_bins, lo, hi, ST(stream));
}
int radmmm_spline_linear_forward(const float* z1, const float* q, const int32_t* lens, float* z1_out, float* log_s, int B,
                                 int Ch, int Tp, int n_bins, float lo, float hi, int inverse, void* stream) {
    return spline_linear(z1, q, lens, z1_out, log_s, B, Ch, Tp, n_bins, lo, hi, inverse, ST(stream));
}
int radmmm_spline_linear_backward(const float* z1, const float* q, const int32_t* lens, const float* dz1_out,
                                  const float* dlog_s, float* dz1, float* dq, int B, int Ch, int Tp, int n_bins, float lo,
                                  float hi, void* stream) {
    return spline_linear_bwd(z1, q, lens, dz1_out, dlog_s, dz1, dq, B, Ch, Tp, n_bins, lo, hi, ST(stream));
}
int radmmm_stft_mel(const float* audio, const float* mel_basis, float* mel, float* magnitude_or_null, int B, int S,
                    int n_fft, int hop, int n_mel, float clip, void* stream) {
    return stft_mel(audio, mel_basis, nullptr, mel, magnitude_or_null, B, S, n_fft, hop, n_mel, clip, ST(stream));
}
int radmmm_mel_support(const float* mel_basis, int n_mel, int n_bins, int32_t* support, void* stream) {
    return mel_support(mel_basis, n_mel, n_bins, support, ST(stream));
}
int radmmm_stft_mel_sparse(const float* audio, const float* mel_basis, const int32_t* support, float* mel,
                           float* magnitude_or_null, int B, int S, int n_fft, int hop, int n_mel, float clip, void* stream) {
    return stft_mel(audio, mel_basis, support, mel, magnitude_or_null, B, S, n_fft, hop, n_mel, clip, ST(stream));
}
long long radmmm_mas_workspace_bytes(int B, int T1, int T2) { return mas_workspace_bytes(B, T1, T2); }
int radmmm_mas_width1(const float* attn, const int32_t* in_lens, const int32_t* out_lens, float* out, int B, int T1, int T2,
                      int is_log, void* workspace, long long workspace_bytes, void* stream) {
    return mas_width1(attn, in_lens, out_lens, out, B, T1, T2, is_log, workspace, workspace_bytes, ST(stream));
}
int radmmm_attention_ctc(const float* attn_logprob, const int32_t* in_lens, const int32_t* out_lens, float* cost, float* grad,
                         int B, int T1, int T2, float blank_logprob, void* stream) {
    return attention_ctc(attn_logprob, in_lens, out_lens, cost, grad, B, T1, T2, blank_logprob, ST(stream));
}
int radmmm_soft_attention(const float* q, const float* k, const float* prior, const int32_t* in_lens, float* attn,
                          float* attn_logprob, const float* txt_enc, float* context, int B, int Ca, int T1, int T2,
                          int Dt, float temperature, void* stream) {
    return soft_attention(q, k, prior, in_lens, attn, attn_logprob, txt_enc, context, B, Ca, T1, T2, Dt, temperature,
                          ST(stream));
}

long long radmmm_soft_attention_backward_workspace_bytes(int B, int T1, int T2) {
    return soft_attention_bwd_workspace_bytes(B, T1, T2);
}
int radmmm_soft_attention_backward(const float* q, const float* k, const float* prior, const int32_t* in_lens,
                                   const float* attn, const float* dattn, const float* dlogprob, const float* txt_enc,
                                   const float* dcontext, float* dq, float* dk, float* dtxt, int B, int Ca, int T1, int T2,
                                   int Dt, float temperature, void* workspace, long long workspace_bytes, void* stream) {
    return soft_attention_bwd(q, k, prior, in_lens, attn, dattn, dlogprob, txt_enc, dcontext, dq, dk, dtxt, B, Ca, T1, T2,
                              Dt, temperature, workspace, workspace_bytes, ST(stream));
}

}  // extern "C"
