/* radmmm_b200 -- C ABI of the B200-native RAD-MMM flow-decoder kernels (libradmmm_b200.so).
 *
 * The reference (NVIDIA/RAD-MMM) is pure Python/PyTorch and has no FFI of its own; the interface each entry point
 * replaces is therefore a Python call site of the reference, cited per function (paths relative to the reference
 * repository).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - caller-allocated outputs and workspaces (query the *_bytes functions); no device allocation and no
 *     synchronisation inside; work is enqueued on the `stream` argument (a cudaStream_t).  The only library-owned
 *     state: per-host-thread auxiliary streams / events (see radmmm_flow_desc.side_stream), the launch counter and
 *     the optional profiler slots;
 *   - return 0 on success, a negative code on failure; radmmm_last_error() gives a thread-local message;
 *   - re-entrant across host threads and CUDA streams.
 *
 * Tensor layouts at the boundary are the reference's: fp32, (batch, channel, time) contiguous, time fastest
 * ("channels-first", cf).  Internally the WN stack works on "rows": row r = b*pitch + t with pitch = Tp + gap
 * zero rows between utterances, R = round_up(B*pitch, 256) rows in total (radmmm_rows()).
 */
#ifndef RADMMM_B200_H
#define RADMMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADMMM_ABI_VERSION 3

/* precision of the WN contractions */
#define RADMMM_MODE_F32 0     /* fp32 FFMA, exact-parity path                                   */
#define RADMMM_MODE_BF16 1    /* tcgen05, bf16 operands, fp32 accumulate (throughput path)       */
#define RADMMM_MODE_BF16X3 2  /* tcgen05, bf16 hi/lo split, 3 MMAs per product (fp32-grade parity) */

/* AffineTransformationLayer.scaling_fn, common.py:1127-1141 */
#define RADMMM_SCALE_TANH 0
#define RADMMM_SCALE_EXP 1
#define RADMMM_SCALE_SIGMOID 2
#define RADMMM_SCALE_TRANSLATE 3

#define RADMMM_MAX_LAYERS 8
#define RADMMM_ROW_GAP 16     /* >= 2 * max dilation (k=5, dilation 2^i, i < 4) */

int radmmm_abi_version(void);
const char* radmmm_last_error(void);
/* Measurement aid (bench.py): when enabled, every contraction launch is bracketed by CUDA events on its own stream.
 * radmmm_profile_collect synchronises the device and returns, per tag (epilogue kind 0..10; 16+taps for the
 * weight-grad GEMM), the launch count, summed device milliseconds and summed executed FLOPs.  Not thread-safe. */
void radmmm_profile_enable(int on);
int radmmm_profile_collect(int max_tags, int* counts, double* ms, double* flops);
/* Diagnostic (tools/gemm_probe.py): when device_buf != NULL every tensor-core contraction launched afterwards writes an
 * in-kernel timeline into region i of the buffer (i = launches since this call, i < max_launches) -- 10 uint64 per CTA for
 * the first max_ctas CTAs (SM clock at entry / set-up / first TMA / first operands / last MMA / accumulator ready /
 * epilogue done / exit, then globaltimer at entry and exit).  NULL disables.  max_launches < 0: the buffer (32 x 8
 * uint64) instead receives the accumulated per-phase SM clocks of the cluster LSTM kernels (csrc/lstm_cluster.cu). */
void radmmm_debug_trace(void* device_buf, int max_ctas, int max_launches);
/* Number of kernels this library has launched in the process so far (every launch site counts itself; memsets and
 * library calls are not included).  bench.py reports the difference over its timed region as "gpu_launches". */
long long radmmm_launch_count(void);
/* sizeof(radmmm_flow_desc) / sizeof(radmmm_flow_grads) as compiled -- lets a binding verify its struct layout */
size_t radmmm_sizeof_flow_desc(void);
size_t radmmm_sizeof_flow_grads(void);

/* rows R for a batch of B utterances padded to Tp grouped frames */
int radmmm_rows(int B, int Tp);
int radmmm_pitch(int Tp);

/* One flow step = Invertible 1x1 conv + WN-parameterised affine coupling.
 * Replaces decoders.FlowStep.forward (decoders.py:72-80) -> common.Invertible1x1ConvLUS.forward (common.py:527-548) /
 * DataInitializedInvertible1x1Conv.forward (common.py:593-617), common.AffineTransformationLayer.forward
 * (common.py:1163-1185) and common.WN.forward (common.py:816-835) with ConvNorm/PartialConv1d (common.py:179-191,
 * partialconv1d.py:57-94) and nn.utils.weight_norm underneath. */
typedef struct radmmm_flow_desc {
    int32_t mode;        /* RADMMM_MODE_* */
    int32_t B;           /* batch */
    int32_t C;           /* channels of this flow step (even) */
    int32_t Tp;          /* grouped frames (time length of every cf tensor) */
    int32_t D;           /* conditioning channels (decoder_cond_dims) */
    int32_t H;           /* WN width (n_channels, 1024 in the reference); multiple of 128 */
    int32_t L;           /* WN layers (n_conv_layers_per_step), dilation 2^i */
    int32_t scaling_fn;  /* RADMMM_SCALE_* */
    int32_t training;    /* 1: keep activations for radmmm_flow_backward */
    int32_t reserved;
    const int32_t* lens; /* (B) grouped lengths */
    /* raw parameters, fp32, reference layouts (state_dict tensors) */
    const float* start_g; const float* start_v; const float* start_b;      /* (H,1,1) (H,C/2+D,1) (H) */
    const float* in_g[RADMMM_MAX_LAYERS]; const float* in_v[RADMMM_MAX_LAYERS]; const float* in_b[RADMMM_MAX_LAYERS];   /* (H,1,1) (H,H,5) (H) */
    const float* rs_g[RADMMM_MAX_LAYERS]; const float* rs_v[RADMMM_MAX_LAYERS]; const float* rs_b[RADMMM_MAX_LAYERS];   /* (H,1,1) (H,H,1) (H) */
    const float* end_w; const float* end_b;                                 /* (C,H,1) (C) */
    /* invertible 1x1 conv: W (C,C) assembled by the caller from P,L,U (common.py:529-531) or U (common.py:598);
     * W_inv (C,C) for the inverse pass; W_T (C,C) = W transposed for the backward pass; mean = input_mean (C) for the
     * whitening layer or NULL.  Only the pointers a call needs have to be non-NULL. */
    const float* W; const float* W_inv; const float* W_T; const float* mean;
    /* prepared (weight-normed, re-laid-out) weights: radmmm_flow_prepared_bytes(), zero-initialised ONCE by the
     * caller, filled by radmmm_flow_prepare() whenever the raw parameters changed */
    void* prepared;
    /* conditioning in row layout, produced once per step by radmmm_context_rows() */
    const void* ctx_rows;
    /* per-call activation workspace (radmmm_flow_workspace_bytes()); must outlive backward when training */
    void* workspace;
    /* optional second cudaStream_t; non-NULL enables intra-call concurrency (NULL = everything on `stream`):
     *   forward / inverse: the dependent GEMM chain runs on a library-owned high-priority stream, the res-skip GEMMs
     *     (which only feed the final `end` GEMM) on `side_stream`;
     *   backward: the input-gradient chain runs on the high-priority stream, the weight-gradient work (bias column sums,
     *     weight-grad GEMMs, weight-norm backward) on four more library-owned streams.
     * The auxiliary streams and fork/join events are created once per host thread; every fork is joined back into
     * `stream` before the call returns, so callers (and CUDA-graph capture of `stream`) see ordinary stream semantics. */
    void* side_stream;
} radmmm_flow_desc;

size_t radmmm_flow_prepared_bytes(int mode, int C, int D, int H, int L);
size_t radmmm_flow_workspace_bytes(int mode, int training, int B, int Tp, int C, int D, int H, int L);
size_t radmmm_flow_backward_scratch_bytes(int mode, int B, int Tp, int C, int D, int H, int L);
size_t radmmm_context_rows_bytes(int mode, int B, int Tp, int D);

/* weight norm + layout.  Replaces the per-forward aten::_weight_norm_interface of nn.utils.weight_norm
 * (common.py:174,791,813). */
int radmmm_flow_prepare(const radmmm_flow_desc* d, void* stream);

/* context (B, Tp, D) fp32 [the bi-LSTM output of models/radmmm.py:137-146 before its transpose] -> row layout */
int radmmm_context_rows(int mode, const float* ctx_btd, const int32_t* lens, int B, int Tp, int D,
                        void* rows, void* stream);
/* gradient rows (fp32 [R][Dp]) -> (B, Tp, D) fp32, accumulate != 0 adds */
int radmmm_context_rows_backward(const float* drows, const int32_t* lens, int B, int Tp, int D, float* dctx_btd,
                                 int accumulate, void* stream);

/* forward: z_in (B,C,Tp) -> z_mid (after the 1x1 conv), params (B,C,Tp), z_out (B,C,Tp), log_s (B,C/2,Tp) */
int radmmm_flow_forward(const radmmm_flow_desc* d, const float* z_in, float* z_mid, float* params, float* z_out,
                        float* log_s, void* stream);
/* inverse (decoders.py:73-76): z (B,C,Tp) -> coupling^-1 -> W_inv (+ mean) -> z_out; params is scratch (B,C,Tp) */
int radmmm_flow_inverse(const radmmm_flow_desc* d, const float* z_in, float* params, float* z_tmp, float* z_out,
                        void* stream);
/* backward.  Incoming gradients are taken as zero beyond each length (RADMMMLoss masks them, loss.py:91,102).
 * Outputs: dz_in (B,C,Tp); dctx_rows fp32 [R][Dp] (overwritten); dW (C,C); parameter gradients in the reference
 * layouts (same shapes as the raw parameters).  dz_mid/dparams are (B,C,Tp) scratch. */
typedef struct radmmm_flow_grads {
    float* start_g; float* start_v; float* start_b;
    float* in_g[RADMMM_MAX_LAYERS]; float* in_v[RADMMM_MAX_LAYERS]; float* in_b[RADMMM_MAX_LAYERS];
    float* rs_g[RADMMM_MAX_LAYERS]; float* rs_v[RADMMM_MAX_LAYERS]; float* rs_b[RADMMM_MAX_LAYERS];
    float* end_w; float* end_b;
    float* W;
} radmmm_flow_grads;

int radmmm_flow_backward(const radmmm_flow_desc* d, const float* z_in, const float* z_mid, const float* params,
                         const float* dz_out, const float* dlog_s, float* dz_mid, float* dparams, float* dz_in,
                         float* dctx_rows, const radmmm_flow_grads* grads, void* scratch, void* stream);

/* Context bi-LSTM recurrence (models/radmmm.py:137-146: pack_padded_sequence -> nn.LSTM(bidirectional, batch_first)
 * -> pad_packed_sequence).  xproj [R][8H] fp32 holds W_ih x + b_ih + b_hh for every row (columns: direction, gate
 * i|f|g|o, unit) and is produced by radmmm_conv_rows; the recurrence runs as ONE kernel per pass:
 *   mode RADMMM_MODE_BF16 (H <= 528): two thread-block clusters of 16 CTAs (one per direction), W_hh as bf16 mma.sync
 *     fragments in registers, h exchanged through distributed shared memory (st.async + mbarrier), fp32 state;
 *   other modes: the fp32 cooperative kernel (W_hh in shared memory, exchange through L2 behind a grid barrier).
 * out (B,Tp,2H) must be zero-initialised (frames beyond each length stay zero); gates [R][8H] and cstate [R][2H] are
 * saved for the backward pass; dgates [R][8H] (zero-initialised) receives the pre-activation gate gradients, from which
 * the weight / input gradients follow as contractions (radmmm_wgrad_rows / radmmm_conv_rows).  B <= 64 per call. */
size_t radmmm_lstm_workspace_bytes(int B, int H);
int radmmm_lstm_forward(int mode, const float* xproj, const float* whh_f, const float* whh_r, const int32_t* lens, int B, int Tp,
                        int H, float* out, float* gates, float* cstate, void* workspace, void* stream);
int radmmm_lstm_backward(int mode, const float* dout, const float* gates, const float* cstate, const float* whh_f,
                         const float* whh_r, const int32_t* lens, int B, int Tp, int H, float* dgates, void* workspace,
                         void* stream);

/* Fused multi-tensor RAdam + global-norm gradient clipping.  Replaces radam.RAdam.step (radam.py:63-142, a Python loop over
 * every parameter tensor) and Lightning's `gradient_clip_val` (configs/RADMMM_train_config.yaml:7-8 ->
 * torch.nn.utils.clip_grad_norm_) with three launches for ALL parameters, no host synchronisation (CUDA-graph capturable).
 *   recs:          device array of n_tensors records {float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
 *                  long long numel} (40 bytes each);
 *   chunk_tensor / chunk_off: device tables, chunk c covers elements [chunk_off[c], + radmmm_radam_chunk_elems()) of tensor
 *                  chunk_tensor[c];
 *   state:         6 device doubles, zero-initialised once: [0] scratch (sum of squares), [1] step count (incremented by the
 *                  call), [2] step size, [3] variance-rectified branch taken, [4] clip coefficient, [5] gradient norm;
 *   cfg:           6 device doubles: lr, beta1, beta2, eps, weight_decay, max_grad_norm (<= 0: no clipping). */
int radmmm_radam_chunk_elems(void);
int radmmm_radam_step(const void* recs, const int32_t* chunk_tensor, const long long* chunk_off, int n_chunks, double* state,
                      const double* cfg, void* stream);

/* Stand-alone ops (also used by the tests) ------------------------------------------------------------- */

/* Invertible 1x1 conv: out[b,co,t] = sum_ci W[co,ci]*(in[b,ci,t]-pre[ci]) + post[co]   (common.py:540-548,605-617) */
int radmmm_inv1x1(const float* in, const float* W, const float* pre, const float* post, float* out, int B, int Cin,
                  int Cout, int Tp, void* stream);
int radmmm_inv1x1_wgrad(const float* dz, const float* x, const float* pre, const int32_t* lens, float* dW, int B,
                        int C, int Tp, void* stream);
/* affine coupling tail (common.py:1173-1185) */
int radmmm_coupling_forward(const float* z, const float* params, float* z_out, float* log_s, int B, int C, int Tp,
                            int scaling_fn, int inverse, void* stream);
int radmmm_coupling_backward(const float* dz_out, const float* dlog_s, const float* z, const float* params,
                             const int32_t* lens, float* dz, float* dparams, int B, int C, int Tp, int scaling_fn,
                             void* stream);
/* masked sums for compute_flow_loss (loss.py:85-110): out[0] += sum over t<len of x (square=0) or x^2 (square=1) */
int radmmm_masked_sum(const float* x, const int32_t* lens, int B, int C, int Tp, int square, double* out, void* stream);
int radmmm_masked_sum_backward(const float* x, const int32_t* lens, int B, int C, int Tp, int square,
                               const float* coef, float coef_mul, float* dx, void* stream);

/* Generic masked dilated conv as a row GEMM (testing / FiLM nets): y rows fp32 [R][N] = sum_taps x[r+(j-c)*d] W_j + bias.
 * x_rows / w in the act format of `mode`; see rad-mmm_b200/csrc/gemm.cuh. */
int radmmm_conv_rows(int mode, const void* x_rows, long long x_ld, long long x_plane, const void* w, long long w_ld,
                     long long w_plane, long long w_tap_stride, const float* bias, float* y, long long y_ld, int R,
                     int K, int N, int taps, int dilation, void* stream);
/* Weight-gradient GEMM on rows: out[tap][m][n] = sum_r dy[r][m] * x[r + (tap - taps/2)*dilation + shift_offset][n].
 * dy [R][M] / x [R][N] are act-format row matrices (M, N multiples of 128 in the tensor-core modes).
 * `out` (fp32, [taps][M][out_ld]) is zeroed by the call. */
int radmmm_wgrad_rows(int mode, const void* dy, long long dy_ld, long long dy_plane, const void* x,
                      long long x_ld, long long x_plane, float* out, long long out_ld,
                      long long out_tap_stride, int R, int M, int N, int taps, int dilation, int shift_offset,
                      void* stream);
/* fp32 rows -> act-format rows of `mode` (hi/lo split for BF16X3) */
int radmmm_cast_rows(int mode, const float* src, long long n, void* dst, long long plane_stride, void* stream);

/* Piecewise-quadratic spline coupling (splines.py:241-339; common.py:1040-1090) */
int radmmm_spline_forward(const float* z1, const float* q, const int32_t* lens, float* z1_out, float* log_s, int B,
                          int Ch, int Tp, int n_bins, float lo, float hi, int inverse, void* stream);
int radmmm_spline_backward(const float* z1, const float* q, const int32_t* lens, const float* dz1_out,
                           const float* dlog_s, float* dz1, float* dq, int B, int Ch, int Tp, int n_bins, float lo,
                           float hi, void* stream);

/* Piecewise-linear spline coupling (splines.py:57-142 forward, 145-238 inverse; the use_quadratic=False branch of
 * SplineTransformationLayer, common.py:1019-1020,1069-1075).  q (B, Ch*n_bins, Tp): n_bins un-normalised bin heights per
 * (channel, frame), n_bins in {8, 16, 32}; z1 is normalised with (z1 - lo) / (hi - lo) and de-normalised on the way out;
 * elements outside [lo, hi] pass through.  log_s (B,1,Tp) (may be NULL): sum over channels of log slope (inverse: minus). */
int radmmm_spline_linear_forward(const float* z1, const float* q, const int32_t* lens, float* z1_out, float* log_s, int B,
                                 int Ch, int Tp, int n_bins, float lo, float hi, int inverse, void* stream);
int radmmm_spline_linear_backward(const float* z1, const float* q, const int32_t* lens, const float* dz1_out,
                                  const float* dlog_s, float* dz1, float* dq, int B, int Ch, int Tp, int n_bins, float lo,
                                  float hi, void* stream);

/* STFT + mel filterbank + log (audio_processing.py:137-154, 227-255): audio (B,S) in [-1,1] -> mel (B,n_mel,S/hop+1).
 * mel_basis (n_mel, n_fft/2+1) fp32 as registered by TacotronSTFT.__init__ (audio_processing.py:124-127). */
int radmmm_stft_mel(const float* audio, const float* mel_basis, float* mel, float* magnitude_or_null, int B, int S,
                    int n_fft, int hop, int n_mel, float clip, void* stream);

/* The same with the mel basis' row supports: support (n_mel x 2 int32: first / last non-zero bin of every row, written by
 * radmmm_mel_support once per basis) lets a CTA read ~9 weights per mel row instead of scanning n_fft/2+1. */
int radmmm_mel_support(const float* mel_basis, int n_mel, int n_bins, int32_t* support, void* stream);
int radmmm_stft_mel_sparse(const float* audio, const float* mel_basis, const int32_t* support, float* mel,
                           float* magnitude_or_null, int B, int S, int n_fft, int hop, int n_mel, float clip, void* stream);

/* Soft attention (common.py:1259-1276) and the context matmul (tts_lightning_modules.py:670).
 * q (B,Ca,T1), k (B,Ca,T2) are the projected queries/keys; prior (B,T1,T2) or NULL; in_lens (B).
 * attn, attn_logprob: (B,1,T1,T2).  If txt_enc (B,Dt,T2) != NULL also writes context (B,Dt,T1). */
int radmmm_soft_attention(const float* q, const float* k, const float* prior, const int32_t* in_lens, float* attn,
                          float* attn_logprob, const float* txt_enc, float* context, int B, int Ca, int T1, int T2,
                          int Dt, float temperature, void* stream);

/* Backward of radmmm_soft_attention: gradients w.r.t. the projected queries / keys (dq (B,Ca,T1), dk (B,Ca,T2)) and, when
 * the context matmul was fused (txt_enc/dcontext != NULL), w.r.t. txt_enc (dtxt (B,Dt,T2)).  dattn / dlogprob are the
 * incoming gradients of attn / attn_logprob (either may be NULL).  workspace: device scratch of
 * radmmm_soft_attention_backward_workspace_bytes(B,T1,T2) bytes (the gradient w.r.t. the squared distances). */
long long radmmm_soft_attention_backward_workspace_bytes(int B, int T1, int T2);
int radmmm_soft_attention_backward(const float* q, const float* k, const float* prior, const int32_t* in_lens,
                                   const float* attn, const float* dattn, const float* dlogprob, const float* txt_enc,
                                   const float* dcontext, float* dq, float* dk, float* dtxt, int B, int Ca, int T1, int T2,
                                   int Dt, float temperature, void* workspace, long long workspace_bytes, void* stream);

/* Monotonic alignment search, width 1 (alignment.py:31-59 `mas_width1`), batched the way its caller loops
 * (tts_lightning_modules.py:270-284 `binarize_attention`): attn (B,1,T1,T2) soft attention (probabilities; is_log != 0:
 * already log-probabilities), in_lens / out_lens (B) text / mel lengths, out (B,1,T1,T2) receives the 0/1 map of
 * attn[b,0,:out_len,:in_len] and zeros elsewhere.  T2 <= 4096.  radmmm_mas_workspace_bytes returns 0 when the back
 * pointers (one bit per cell) fit in shared memory, else the size of the device workspace to pass. */
long long radmmm_mas_workspace_bytes(int B, int T1, int T2);
int radmmm_mas_width1(const float* attn, const int32_t* in_lens, const int32_t* out_lens, float* out, int B, int T1, int T2,
                      int is_log, void* workspace, long long workspace_bytes, void* stream);

/* AttentionCTCLoss (loss.py:112-140) forward AND gradient in one launch: attn_logprob (B,1,T1,T2), key lengths in_lens,
 * query lengths out_lens; cost (B) receives each utterance's CTC loss (nn.CTCLoss reduction='mean' over its one target:
 * nll / key_len; zero_infinity: 0 when infeasible) -- the module's value is mean(cost); grad (B,1,T1,T2) receives
 * d mean(cost) / d attn_logprob.  T2 <= 1023. */
int radmmm_attention_ctc(const float* attn_logprob, const int32_t* in_lens, const int32_t* out_lens, float* cost, float* grad,
                         int B, int T1, int T2, float blank_logprob, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADMMM_B200_H */
