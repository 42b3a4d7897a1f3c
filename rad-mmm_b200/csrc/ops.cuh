// Internal declarations of the non-GEMM kernels' host launchers (elementwise.cu, spline.cu, stft.cu, attention.cu).
#pragma once
#include "common.cuh"

namespace radmmm {

enum { SCALE_TANH = 0, SCALE_EXP = 1, SCALE_SIGMOID = 2, SCALE_TRANSLATE = 3 };

int rows_from_cf(int mode, const float* src, long long batch_stride, int n_ch, const RowGeom& g, ActMat dst,
                 int n_cols, int mask_invalid, cudaStream_t st);
int rows_from_btd(int mode, const float* src, int D, const RowGeom& g, ActMat dst, int n_cols, cudaStream_t st);
int btd_from_rows(const float* rows, long long ld, int D, const RowGeom& g, float* dst, int accumulate, cudaStream_t st);
int coupling_fwd(const float* z, const float* params, float* z_out, float* log_s, int B, int C, int Tp, int fn,
                 int inverse, cudaStream_t st);
int coupling_bwd(const float* dz_out, const float* dlog_s, const float* z, const float* params, const int* lens,
                 float* dz, float* dparams, int B, int C, int Tp, int fn, cudaStream_t st);
int inv1x1(const float* in, long long in_bs, const float* W, const float* pre, const float* post, float* out,
           long long out_bs, int B, int Cin, int Cout, int Tp, cudaStream_t st);
int inv1x1_wgrad(const float* dz, const float* x, const float* pre, const int* lens, float* dW, int B, int C, int Tp,
                 cudaStream_t st);
int masked_sum(const float* x, const int* lens, int B, int C, int Tp, int square, double* out, cudaStream_t st);
int masked_sum_bwd(const float* x, const int* lens, int B, int C, int Tp, int square, const float* coef_ptr,
                   float coef_mul, float* dx, cudaStream_t st);
int wn_norm(const float* v, int n_co, int per_co, float* norm, float* rowsum, cudaStream_t st);
int wn_scatter(int mode, const float* v, const float* g, const float* norm, int n_co, int ci_total, int ksize,
               int ci_begin, int n_ci, ActMat dst, long long dst_tap, ActMat dstT, long long dstT_tap, cudaStream_t st);
int padq_compute(const float* g, const float* norm, const float* rowsum, const float* bias, float* padq, int n,
                 cudaStream_t st);
int wn_bwd(const float* src0, long long ld0, long long tap0, int n_ci0, const float* src1, long long ld1,
           long long tap1, const float* v, const float* g, const float* norm, int n_co, int ci_total, int ksize,
           float* dv, float* dg, cudaStream_t st);
int colsum(int mode, ActMat x, const RowGeom& g, int n_cols, int dilation, int unratio, float* out, cudaStream_t st);
int cast_rows(int mode, const float* src, long long n, void* dst, long long plane_stride, cudaStream_t st);
// zero up to 16 small fp32 buffers with one launch
struct ZeroList { float* ptr[16]; int count[16]; int n; };
int zero_list(const ZeroList& z, cudaStream_t st);

int radam_chunk_elems();
int radam_step(const void* recs, const int* chunk_tensor, const long long* chunk_off, int n_chunks, double* state, const double* cfg,
               cudaStream_t st);

size_t lstm_workspace_bytes(int B, int H);
int lstm_forward(int mode, const float* xproj, const float* whh_f, const float* whh_r, const int* lens, int B, int Tp, int H,
                 float* out, float* gates, float* cstate, void* workspace, cudaStream_t st);
int lstm_backward(int mode, const float* dout, const float* gates, const float* cstate, const float* whh_f, const float* whh_r,
                  const int* lens, int B, int Tp, int H, float* dgates, void* workspace, cudaStream_t st);

int spline_fwd(const float* z1, const float* q, const int* lens, float* z1_out, float* log_s, int B, int Ch, int Tp,
               int n_bins, float lo, float hi, int inverse, cudaStream_t st);
int spline_bwd(const float* z1, const float* q, const int* lens, const float* dz1_out, const float* dlog_s, float* dz1,
               float* dq, int B, int Ch, int Tp, int n_bins, float lo, float hi, cudaStream_t st);
int spline_linear(const float* z1, const float* q, const int* lens, float* z1_out, float* log_s, int B, int Ch, int Tp, int n_bins,
                  float lo, float hi, int inverse, cudaStream_t st);
int spline_linear_bwd(const float* z1, const float* q, const int* lens, const float* dz1_out, const float* dlog_s, float* dz1, float* dq,
                      int B, int Ch, int Tp, int n_bins, float lo, float hi, cudaStream_t st);
int mel_support(const float* mel_basis, int n_mel, int n_bins, int* support, cudaStream_t st);
int stft_mel(const float* audio, const float* mel_basis, const int* support, float* mel, float* mag, int B, int S, int n_fft,
             int hop, int n_mel, float clip, cudaStream_t st);
long long mas_workspace_bytes(int B, int T1, int T2);
int mas_width1(const float* attn, const int* in_lens, const int* out_lens, float* out, int B, int T1, int T2, int is_log,
               void* workspace, long long workspace_bytes, cudaStream_t st);
int attention_ctc(const float* logprob, const int* in_lens, const int* out_lens, float* cost, float* grad, int B, int T1, int T2,
                  float blank_logprob, cudaStream_t st);
int soft_attention(const float* q, const float* k, const float* prior, const int* in_lens, float* attn,
                   float* attn_logprob, const float* txt_enc, float* context, int B, int Ca, int T1, int T2, int Dt,
                   float temperature, cudaStream_t st);

long long soft_attention_bwd_workspace_bytes(int B, int T1, int T2);
int soft_attention_bwd(const float* q, const float* k, const float* prior, const int* in_lens, const float* attn,
                       const float* dattn, const float* dlogprob, const float* txt_enc, const float* dcontext, float* dq,
                       float* dk, float* dtxt, int B, int Ca, int T1, int T2, int Dt, float temperature, void* workspace,
                       long long workspace_bytes, cudaStream_t st);

}  // namespace radmmm
