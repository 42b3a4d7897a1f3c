// One flow step (invertible 1x1 conv + WN + affine coupling): weight preparation, forward, inverse, backward.
// Host-side orchestration only -- every launch goes to the caller's stream, nothing allocates or synchronises.
#include "../../include/radmmm_b200.h"
#include <stdlib.h>
#include "gemm.cuh"
#include "ops.cuh"

namespace radmmm {

struct Dims {
    int mode, B, C, Ch, Tp, D, H, L, pitch, R, Kz, KzR, Dp, Cp, Nend;
    size_t es;
    int planes;
};

static Dims make_dims(int mode, int B, int C, int Tp, int D, int H, int L) {
    Dims d;
    d.mode = mode; d.B = B; d.C = C; d.Ch = C / 2; d.Tp = Tp; d.D = D; d.H = H; d.L = L;
    d.pitch = Tp + RADMMM_ROW_GAP;
    d.R = (int)round_up((long long)B * d.pitch, 256);      // even number of 128-row tiles: 2x2 cluster tiling
    // every padded channel count is a multiple of 128 so that TMA boxes (128 rows) never exceed a tensor extent
    d.Kz = (int)round_up(d.Ch, 128);
    d.KzR = d.Kz;
    d.Dp = (int)round_up(D, 128);
    d.Cp = (int)round_up(C, 128);
    d.Nend = (int)round_up(C, 128);
    d.es = mode_elem_bytes(mode);
    d.planes = mode_planes(mode);
    return d;
}

struct Bump {
    char* base;
    size_t off;
    void* take(size_t bytes) {
        off = (size_t)round_up((long long)off, 1024);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    }
};

static ActMat take_act(Bump& b, const Dims& d, long long rows, long long ld) {
    ActMat m;
    m.ld = ld;
    m.plane_stride = rows * ld;
    m.ptr = b.take((size_t)d.planes * rows * ld * d.es);
    return m;
}
static ActMat null_act() { ActMat m; m.ptr = nullptr; m.ld = 0; m.plane_stride = 0; return m; }

// ------------------------------------------------------------------------------------------------ prepared weights
struct Prepared {
    float *norm_start, *norm_in, *norm_rs, *rowsum_rs, *padq;          // [H], [L][H] ...
    ActMat Wz, Wc, Win, Wrs, Wend;                                      // Win: [L*5] matrices of [H][H]; Wrs: [L] of [H][H]
    ActMat WzT, WcT, WinT, WrsT, WendT;
    long long HH;                                                       // H*H (elements between matrices)
};

static size_t layout_prepared(const Dims& d, void* base, Prepared* p) {
    Bump b{(char*)base, 0};
    Prepared q;
    const long long H = d.H, L = d.L;
    q.HH = H * H;
    q.norm_start = (float*)b.take(sizeof(float) * H);
    q.norm_in = (float*)b.take(sizeof(float) * L * H);
    q.norm_rs = (float*)b.take(sizeof(float) * L * H);
    q.rowsum_rs = (float*)b.take(sizeof(float) * L * H);
    q.padq = (float*)b.take(sizeof(float) * L * H);
    q.Wz = take_act(b, d, H, d.Kz);
    q.Wc = take_act(b, d, H, d.Dp);
    q.Win = take_act(b, d, L * 5 * H, H);
    q.Wrs = take_act(b, d, L * H, H);
    q.Wend = take_act(b, d, d.Nend, H);
    q.WzT = take_act(b, d, d.KzR, H);
    q.WcT = take_act(b, d, d.Dp, H);
    q.WinT = take_act(b, d, L * 5 * H, H);
    q.WrsT = take_act(b, d, L * H, H);
    q.WendT = take_act(b, d, H, d.Cp);
    if (p) *p = q;
    return (size_t)round_up((long long)b.off, 1024);
}

static ActMat sub_mode(const ActMat& m, long long elem_off, size_t es) {
    ActMat r = m;
    r.ptr = (char*)m.ptr + elem_off * es;
    return r;
}

// ------------------------------------------------------------------------------------------------ activations
struct Workspace {
    ActMat Z0;
    ActMat Hs[RADMMM_MAX_LAYERS + 1];
    ActMat S[RADMMM_MAX_LAYERS];      // s_i = softplus(res_skip_i(h_{i+1})), kept per layer (the skip sum is implicit)
};

static size_t layout_workspace(const Dims& d, int training, void* base, Workspace* w) {
    Bump b{(char*)base, 0};
    Workspace q;
    q.Z0 = take_act(b, d, d.R, d.Kz);
    const int nH = training ? d.L + 1 : 2;
    for (int i = 0; i <= d.L; ++i) {
        if (i < nH) {
            q.Hs[i] = take_act(b, d, d.R, d.H);
        } else {
            q.Hs[i] = q.Hs[i & 1];
        }
    }
    for (int i = 0; i < d.L; ++i) q.S[i] = take_act(b, d, d.R, d.H);
    if (w) *w = q;
    return (size_t)round_up((long long)b.off, 1024);
}

struct Scratch {
    ActMat DP;
    ActMat DQ[RADMMM_MAX_LAYERS];
    ActMat DACC[RADMMM_MAX_LAYERS + 1];    // one per layer + dh0: the chain never waits for the weight-gradient lanes
    float* dW_in[RADMMM_MAX_LAYERS];       // [5][H][H] fp32 per dilated conv
    float* dW_rs[RADMMM_MAX_LAYERS];       // [H][H] fp32 per res-skip conv
    float* dW_start;                       // dWz [H][Kz] + dWc [H][Dp]
};

static size_t layout_scratch(const Dims& d, void* base, Scratch* s) {
    Bump b{(char*)base, 0};
    Scratch q;
    q.DP = take_act(b, d, d.R, d.Cp);
    for (int i = 0; i < d.L; ++i) q.DQ[i] = take_act(b, d, d.R, d.H);
    for (int i = 0; i <= d.L; ++i) q.DACC[i] = take_act(b, d, d.R, d.H);
    for (int i = 0; i < d.L; ++i) {
        q.dW_in[i] = (float*)b.take(sizeof(float) * (size_t)5 * d.H * d.H);
        q.dW_rs[i] = (float*)b.take(sizeof(float) * (size_t)d.H * d.H);
    }
    q.dW_start = (float*)b.take(sizeof(float) * (size_t)d.H * (d.Kz + d.Dp));
    if (s) *s = q;
    return (size_t)round_up((long long)b.off, 1024);
}

static RowGeom geom_of(const Dims& d, const int* lens) {
    RowGeom g;
    g.lens = lens; g.B = d.B; g.Tp = d.Tp; g.pitch = d.pitch; g.R = d.R;
    return g;
}

static int check_desc(const radmmm_flow_desc* f) {
    RADMMM_REQUIRE(f != nullptr, "flow desc is null");
    RADMMM_REQUIRE(f->mode >= 0 && f->mode <= 2, "bad mode %d", f->mode);
    RADMMM_REQUIRE(f->C > 0 && f->C % 2 == 0 && f->C <= 256, "C=%d must be even and <= 256", f->C);
    RADMMM_REQUIRE(f->H > 0 && f->H % 128 == 0, "H=%d must be a multiple of 128", f->H);
    RADMMM_REQUIRE(f->L >= 1 && f->L <= 4, "L=%d: dilation 2^(L-1) must stay within the %d-row gap", f->L, RADMMM_ROW_GAP);
    RADMMM_REQUIRE(f->B > 0 && f->Tp > 0 && f->D > 0, "bad sizes B=%d Tp=%d D=%d", f->B, f->Tp, f->D);
    RADMMM_REQUIRE(f->prepared != nullptr, "prepared weights buffer is null");
    return RADMMM_OK;
}

// ------------------------------------------------------------------------------------------------ prepare
int flow_prepare(const radmmm_flow_desc* f, cudaStream_t st) {
    RADMMM_TRY(check_desc(f));
    const Dims d = make_dims(f->mode, f->B, f->C, f->Tp, f->D, f->H, f->L);
    Prepared p;
    layout_prepared(d, f->prepared, &p);
    const int H = d.H, cin = d.Ch + d.D;
    const ActMat none = null_act();
    // start: (H, Ch + D, 1) -> Wz [H][Kz], Wc [H][Dp] and transposes
    RADMMM_TRY(wn_norm(f->start_v, H, cin, p.norm_start, nullptr, st));
    RADMMM_TRY(wn_scatter(d.mode, f->start_v, f->start_g, p.norm_start, H, cin, 1, 0, d.Ch, p.Wz, 0, p.WzT, 0, st));
    RADMMM_TRY(wn_scatter(d.mode, f->start_v, f->start_g, p.norm_start, H, cin, 1, d.Ch, d.D, p.Wc, 0, p.WcT, 0, st));
    for (int i = 0; i < d.L; ++i) {
        RADMMM_TRY(wn_norm(f->in_v[i], H, H * 5, p.norm_in + (size_t)i * H, nullptr, st));
        RADMMM_TRY(wn_scatter(d.mode, f->in_v[i], f->in_g[i], p.norm_in + (size_t)i * H, H, H, 5, 0, H,
                              sub_mode(p.Win, (long long)i * 5 * p.HH, d.es), p.HH,
                              sub_mode(p.WinT, (long long)i * 5 * p.HH, d.es), p.HH, st));
        RADMMM_TRY(wn_norm(f->rs_v[i], H, H, p.norm_rs + (size_t)i * H, p.rowsum_rs + (size_t)i * H, st));
        RADMMM_TRY(wn_scatter(d.mode, f->rs_v[i], f->rs_g[i], p.norm_rs + (size_t)i * H, H, H, 1, 0, H,
                              sub_mode(p.Wrs, (long long)i * p.HH, d.es), 0, sub_mode(p.WrsT, (long long)i * p.HH, d.es), 0, st));
        RADMMM_TRY(padq_compute(f->rs_g[i], p.norm_rs + (size_t)i * H, p.rowsum_rs + (size_t)i * H, f->rs_b[i],
                                p.padq + (size_t)i * H, H, st));
    }
    // end: plain conv (C, H, 1)
    RADMMM_TRY(wn_scatter(d.mode, f->end_w, nullptr, nullptr, d.C, H, 1, 0, H, p.Wend, 0, p.WendT, 0, st));
    (void)none;
    return RADMMM_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM builders
static void init_args(GemmArgs& a, const Dims& d, const int* lens, int kind, int N) {
    memset(&a, 0, sizeof(a));
    a.R = d.R;
    a.epi.kind = kind;
    a.epi.geom = geom_of(d, lens);
    a.epi.N = N;
    a.epi.n_layers = d.L;
}
static void add_seg(GemmArgs& a, const ActMat& act, const ActMat& w, int K, int shift) {
    GemmSeg& s = a.seg[a.n_seg++];
    s.a = act; s.w = w; s.K = K; s.shift = shift;
}

// Auxiliary streams for the weight-gradient lanes of flow_backward (created once per host thread).  Each lane runs
// complete, mutually independent chains (bias column sum -> weight-grad GEMM -> weight-norm backward) with private
// scratch, so the only serial chain left is the input-gradient GEMMs on the caller's stream.
constexpr int kLanes = 4;
static cudaStream_t lane_stream(int i) {
    static thread_local cudaStream_t st[kLanes] = {};
    if (!st[i]) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    return st[i];
}
static cudaStream_t chain_stream() {
    static thread_local cudaStream_t st = nullptr;
    if (!st) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest);
    }
    return st;
}
static cudaEvent_t lane_event(int i) {
    static thread_local cudaEvent_t ev[64] = {};
    if (!ev[i]) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    return ev[i];
}

// events for the fork/join with the side stream (created once per host thread; timing disabled)
static cudaEvent_t side_event(int i) {
    static thread_local cudaEvent_t ev[32] = {};
    if (!ev[i]) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    return ev[i];
}

// WN forward on rows: fills ws (H*, S*) and writes params (B, C, Tp).
// The dilated convs form the dependent chain h_0 -> h_1 -> ... on `st`; res-skip conv i only feeds the `end` GEMM, so it
// runs on the side stream (when the descriptor carries one) and fills the SMs the chain's 104-tile launches leave idle.
static int wn_forward_rows(const radmmm_flow_desc* f, const Dims& d, const Prepared& p, const Workspace& w,
                           const float* z_mid, float* params, cudaStream_t caller) {
    const RowGeom g = geom_of(d, f->lens);
    ActMat ctx; ctx.ptr = const_cast<void*>(f->ctx_rows); ctx.ld = d.Dp; ctx.plane_stride = (long long)d.R * d.Dp;
    cudaStream_t sd = f->side_stream ? reinterpret_cast<cudaStream_t>(f->side_stream) : caller;
    const bool forked = sd != caller;
    cudaStream_t st = caller;                 // the chain's stream: high priority when forked (see flow_backward)
    if (forked) {
        st = chain_stream();
        RADMMM_CUDA(cudaEventRecord(side_event(22), caller));
        RADMMM_CUDA(cudaStreamWaitEvent(st, side_event(22), 0));
    }
    // z0 = z_mid[:, :Ch] -> rows
    RADMMM_TRY(rows_from_cf(d.mode, z_mid, (long long)d.C * d.Tp, d.Ch, g, w.Z0, d.Kz, 1, st));
    GemmArgs a;
    // start
    init_args(a, d, f->lens, EPI_START, d.H);
    add_seg(a, w.Z0, p.Wz, d.Kz, 0);
    add_seg(a, ctx, p.Wc, d.Dp, 0);
    a.epi.bias = f->start_b;
    a.epi.out0 = w.Hs[0];
    RADMMM_TRY(launch_gemm(a, d.mode, st));
    for (int i = 0; i < d.L; ++i) {
        const int dil = 1 << i;
        // inference re-uses two h buffers: h_{i+1} overwrites h_{i-1}, which res-skip conv i-2 may still be reading
        if (forked && !f->training && i >= 2) RADMMM_CUDA(cudaStreamWaitEvent(st, side_event(20 + ((i - 2) & 1)), 0));
        init_args(a, d, f->lens, EPI_IN, d.H);
        for (int j = 0; j < 5; ++j)
            add_seg(a, w.Hs[i], sub_mode(p.Win, ((long long)i * 5 + j) * p.HH, d.es), d.H, (j - 2) * dil);
        a.epi.bias = f->in_b[i];
        a.epi.dilation = dil;
        a.epi.out0 = w.Hs[i + 1];
        RADMMM_TRY(launch_gemm(a, d.mode, st));
        if (forked) {
            RADMMM_CUDA(cudaEventRecord(side_event(16 + (i & 1)), st));          // h_{i+1} ready
            RADMMM_CUDA(cudaStreamWaitEvent(sd, side_event(16 + (i & 1)), 0));
        }
        init_args(a, d, f->lens, EPI_RS, d.H);
        add_seg(a, w.Hs[i + 1], sub_mode(p.Wrs, (long long)i * p.HH, d.es), d.H, 0);
        a.epi.bias = f->rs_b[i];
        a.epi.padq = p.padq + (size_t)i * d.H;
        a.epi.out0 = w.S[i];
        RADMMM_TRY(launch_gemm(a, d.mode, sd));
        if (forked && !f->training) RADMMM_CUDA(cudaEventRecord(side_event(20 + (i & 1)), sd));   // done reading h_{i+1}
    }
    if (forked) {                                                                 // join: every s_i is ready
        RADMMM_CUDA(cudaEventRecord(side_event(18), sd));
        RADMMM_CUDA(cudaStreamWaitEvent(st, side_event(18), 0));
    }
    // end(sum_i s_i) = sum_i end(s_i): one GEMM whose K runs over the L stored s_i
    init_args(a, d, f->lens, EPI_END, d.C);
    for (int i = 0; i < d.L; ++i) add_seg(a, w.S[i], p.Wend, d.H, 0);
    a.epi.bias = f->end_b;
    a.epi.cf_out = params; a.epi.cf_C = d.C; a.epi.cf_c0 = 0;
    RADMMM_TRY(launch_gemm(a, d.mode, st));
    if (forked) {                                                                 // hand the result back to the caller
        RADMMM_CUDA(cudaEventRecord(side_event(23), st));
        RADMMM_CUDA(cudaStreamWaitEvent(caller, side_event(23), 0));
    }
    return RADMMM_OK;
}

int flow_forward(const radmmm_flow_desc* f, const float* z_in, float* z_mid, float* params, float* z_out,
                 float* log_s, cudaStream_t st) {
    RADMMM_TRY(check_desc(f));
    RADMMM_REQUIRE(f->workspace && f->ctx_rows, "flow_forward: workspace / ctx_rows missing");
    const Dims d = make_dims(f->mode, f->B, f->C, f->Tp, f->D, f->H, f->L);
    Prepared p; layout_prepared(d, f->prepared, &p);
    Workspace w; layout_workspace(d, f->training, f->workspace, &w);
    const long long bs = (long long)d.C * d.Tp;
    if (f->W) {
        RADMMM_TRY(inv1x1(z_in, bs, f->W, f->mean, nullptr, z_mid, bs, d.B, d.C, d.C, d.Tp, st));
    } else {      // coupling only (AffineTransformationLayer on its own): z_mid aliases z_in
        RADMMM_REQUIRE(z_mid == z_in, "flow_forward: without W, z_mid must alias z_in");
    }
    RADMMM_TRY(wn_forward_rows(f, d, p, w, z_mid, params, st));
    RADMMM_TRY(coupling_fwd(z_mid, params, z_out, log_s, d.B, d.C, d.Tp, f->scaling_fn, 0, st));
    return RADMMM_OK;
}

int flow_inverse(const radmmm_flow_desc* f, const float* z_in, float* params, float* z_tmp, float* z_out,
                 cudaStream_t st) {
    RADMMM_TRY(check_desc(f));
    RADMMM_REQUIRE(f->workspace && f->ctx_rows && f->W_inv, "flow_inverse: workspace / ctx_rows / W_inv missing");
    RADMMM_REQUIRE(!f->training, "flow_inverse: training must be 0");
    const Dims d = make_dims(f->mode, f->B, f->C, f->Tp, f->D, f->H, f->L);
    Prepared p; layout_prepared(d, f->prepared, &p);
    Workspace w; layout_workspace(d, 0, f->workspace, &w);
    const long long bs = (long long)d.C * d.Tp;
    RADMMM_TRY(wn_forward_rows(f, d, p, w, z_in, params, st));
    RADMMM_TRY(coupling_fwd(z_in, params, z_tmp, nullptr, d.B, d.C, d.Tp, f->scaling_fn, 1, st));
    RADMMM_TRY(inv1x1(z_tmp, bs, f->W_inv, nullptr, f->mean, z_out, bs, d.B, d.C, d.C, d.Tp, st));
    return RADMMM_OK;
}

// ------------------------------------------------------------------------------------------------ backward
static int wgrad(const Dims& d, const int* lens, const ActMat& dY, const ActMat& X, int M, int N, int taps, int dil, float* out, long long ld, long long tap_stride, cudaStream_t st,
                 bool zero_out = true) {
    GemmArgs a;
    init_args(a, d, lens, EPI_WGRAD, N);
    a.wgrad = 1;
    for (int j = 0; j < taps; ++j) {
        GemmSeg& s = a.seg[a.n_seg++];
        s.a = dY; s.w = X; s.K = d.R;
        s.shift = taps == 1 ? 0 : (j - 2) * dil;
    }
    a.epi.M = M;
    a.epi.f32_out = out; a.epi.f32_ld = ld; a.epi.f32_tap_stride = tap_stride;
    // every launcher picks its own tiling / split-K and zeroes the output itself when it reduces with atomics
    a.split_k = 0;
    a.epi.atomic = 1;
    a.zero_output = zero_out ? 1 : 0;
    return launch_gemm(a, d.mode, st);
}

int flow_backward(const radmmm_flow_desc* f, const float* z_in, const float* z_mid, const float* params,
                  const float* dz_out, const float* dlog_s, float* dz_mid, float* dparams, float* dz_in,
                  float* dctx_rows, const radmmm_flow_grads* gr, void* scratch, cudaStream_t st) {
    RADMMM_TRY(check_desc(f));
    RADMMM_REQUIRE(f->training, "flow_backward: the forward pass must have run with training=1");
    RADMMM_REQUIRE(f->workspace && f->ctx_rows && gr && scratch, "flow_backward: missing buffers");
    const Dims d = make_dims(f->mode, f->B, f->C, f->Tp, f->D, f->H, f->L);
    Prepared p; layout_prepared(d, f->prepared, &p);
    Workspace w; layout_workspace(d, 1, f->workspace, &w);
    Scratch s; layout_scratch(d, scratch, &s);
    const RowGeom g = geom_of(d, f->lens);
    const int H = d.H, L = d.L;
    ActMat ctx; ctx.ptr = const_cast<void*>(f->ctx_rows); ctx.ld = d.Dp; ctx.plane_stride = (long long)d.R * d.Dp;
    GemmArgs a;
    // `st` carries the input-gradient chain (the critical path).  With a side stream in the descriptor the weight-gradient
    // work is spread over kLanes auxiliary streams; every chain waits for the chain product it consumes (event) and all
    // lanes are joined back into `st` before returning.  Without one, everything runs on `st`.
    static const bool lanes_on = []() { const char* e = getenv("RADMMM_B200_LANES"); return !(e && e[0] == '0'); }();
    const bool forked = f->side_stream != nullptr && lanes_on;
    int n_ev = 0;
    // the chain itself runs on a high-priority stream: its CTAs win the SMs whenever a lane kernel and a chain kernel
    // are both waiting for them
    cudaStream_t mc = st;
    if (forked) {
        mc = chain_stream();
        cudaEvent_t e = lane_event(n_ev++);
        RADMMM_CUDA(cudaEventRecord(e, st));
        RADMMM_CUDA(cudaStreamWaitEvent(mc, e, 0));
    }
    auto lane = [&](int i) -> cudaStream_t { return forked ? lane_stream(i % kLanes) : st; };
    auto ready = [&](cudaStream_t to) -> int {       // `to` waits for everything enqueued on the chain so far
        if (!forked) return RADMMM_OK;
        cudaEvent_t e = lane_event(n_ev++);
        RADMMM_CUDA(cudaEventRecord(e, mc));
        RADMMM_CUDA(cudaStreamWaitEvent(to, e, 0));
        return RADMMM_OK;
    };

    // Bias gradients of the WN convs: on the tensor-core path the gradient epilogues accumulate them (column sums with
    // red.global.add), so the 2L + 1 buffers are zeroed here by one small launch; the fp32 path uses colsum() kernels.
    const bool fuse_bias = d.mode != MODE_F32;
    if (fuse_bias) {
        ZeroList z;
        z.n = 0;
        for (int i = 0; i < L; ++i) { z.ptr[z.n] = gr->rs_b[i]; z.count[z.n++] = H; z.ptr[z.n] = gr->in_b[i]; z.count[z.n++] = H; }
        z.ptr[z.n] = gr->start_b; z.count[z.n++] = H;
        RADMMM_TRY(zero_list(z, mc));
    }
    // 1. coupling tail
    RADMMM_TRY(coupling_bwd(dz_out, dlog_s, z_mid, params, f->lens, dz_mid, dparams, d.B, d.C, d.Tp, f->scaling_fn, mc));
    RADMMM_TRY(rows_from_cf(d.mode, dparams, (long long)d.C * d.Tp, d.C, g, s.DP, d.Cp, 1, mc));
    // 2. end conv: input gradient on the chain ...
    init_args(a, d, f->lens, EPI_DOUT, H);
    add_seg(a, s.DP, p.WendT, d.Cp, 0);
    for (int i = 0; i < L; ++i) { a.epi.sig[i] = w.S[i]; a.epi.dq[i] = s.DQ[i]; if (fuse_bias) a.epi.colsum[i] = gr->rs_b[i]; }
    RADMMM_TRY(launch_gemm(a, d.mode, mc));
    //    ... DP and every DQ_i are ready: the end conv and all L res-skip convs can take their weight gradients now
    for (int l = 0; l < (forked ? kLanes : 0); ++l) RADMMM_TRY(ready(lane_stream(l)));
    {
        cudaStream_t sd = lane(3);
        RADMMM_TRY(colsum(d.mode, s.DP, g, d.C, 1, 0, gr->end_b, sd));
        // dW_end = dP^T (sum_i s_i): one weight-grad GEMM accumulating over the L stored s_i
        GemmArgs wa;
        init_args(wa, d, f->lens, EPI_WGRAD, H);
        wa.wgrad = 2;
        for (int i = 0; i < L; ++i) { GemmSeg& sg = wa.seg[wa.n_seg++]; sg.a = s.DP; sg.w = w.S[i]; sg.K = d.R; sg.shift = 0; }
        wa.epi.M = d.C; wa.epi.f32_out = gr->end_w; wa.epi.f32_ld = H; wa.epi.f32_tap_stride = 0;
        wa.split_k = 0; wa.epi.atomic = 1;
        RADMMM_CUDA(cudaMemsetAsync(gr->end_w, 0, sizeof(float) * (size_t)d.C * H, sd));
        RADMMM_TRY(launch_gemm(wa, d.mode, sd));
    }
    {   // res-skip convs (need only DQ_i and the saved h_{i+1}): ONE grouped weight-grad launch for all L layers -- L x 32
        // full-K tiles fill the machine once; one by one each problem is 32 tiles and had to be split-K'ed six ways with fp32
        // atomics into a zeroed buffer (99 TFLOP/s).  Bias column sums run beside it on the other lane.
        cudaStream_t sd = lane(0), sb = lane(1);
        GemmArgs wa;
        init_args(wa, d, f->lens, EPI_WGRAD, H);
        wa.wgrad = 3;
        for (int i = 0; i < L; ++i) {
            GemmSeg& sg = wa.seg[wa.n_seg++];
            sg.a = s.DQ[i]; sg.w = w.Hs[i + 1]; sg.K = d.R; sg.shift = 0;
            wa.epi.group_out[i] = s.dW_rs[i];
        }
        wa.epi.M = H; wa.epi.f32_out = s.dW_rs[0]; wa.epi.f32_ld = H; wa.epi.f32_tap_stride = 0;
        wa.split_k = 1; wa.epi.atomic = 0; wa.zero_output = 0;
        RADMMM_TRY(launch_gemm(wa, d.mode, sd));
        for (int i = L - 1; i >= 0; --i) {
            if (!fuse_bias) RADMMM_TRY(colsum(d.mode, s.DQ[i], g, H, 1, 0, gr->rs_b[i], sb));
            RADMMM_TRY(wn_bwd(s.dW_rs[i], H, 0, H, nullptr, 0, 0, f->rs_v[i], f->rs_g[i], p.norm_rs + (size_t)i * H, H, H, 1,
                              gr->rs_v[i], gr->rs_g[i], sd));        // same lane as the grouped launch that produced dW_rs
        }
    }
    // 3. layers, last to first
    for (int i = L - 1; i >= 0; --i) {
        const int dil = 1 << i;
        // chain: dh_{i+1} = Wrs_i^T dq_i + sum_taps Win_{i+1,j}^T dacc_{i+1}[r - (j-2) d_{i+1}]  -> dacc_i
        init_args(a, d, f->lens, EPI_DH, H);
        add_seg(a, s.DQ[i], sub_mode(p.WrsT, (long long)i * p.HH, d.es), H, 0);
        if (i < L - 1)
            for (int j = 0; j < 5; ++j)
                add_seg(a, s.DACC[i + 1], sub_mode(p.WinT, ((long long)(i + 1) * 5 + j) * p.HH, d.es), H, -(j - 2) * (dil * 2));
        a.epi.h = w.Hs[i + 1];
        a.epi.dilation = dil;
        a.epi.out0 = s.DACC[i];
        if (fuse_bias) a.epi.colsum[0] = gr->in_b[i];
        RADMMM_TRY(launch_gemm(a, d.mode, mc));
        // dilated conv i: bias (un-ratio'd) and weights on lanes 2 and 3
        cudaStream_t sd = lane(2 + (i & 1));
        RADMMM_TRY(ready(sd));                // dacc_i ready
        if (!fuse_bias) RADMMM_TRY(colsum(d.mode, s.DACC[i], g, H, dil, 1, gr->in_b[i], sd));
        RADMMM_TRY(wgrad(d, f->lens, s.DACC[i], w.Hs[i], H, H, 5, dil, s.dW_in[i], H, p.HH, sd));
        RADMMM_TRY(wn_bwd(s.dW_in[i], H, p.HH, H, nullptr, 0, 0, f->in_v[i], f->in_g[i], p.norm_in + (size_t)i * H, H, H, 5,
                          gr->in_v[i], gr->in_g[i], sd));
    }
    // 4. dh0 (masked) from layer 0's dilated conv
    init_args(a, d, f->lens, EPI_DH0, H);
    for (int j = 0; j < 5; ++j) add_seg(a, s.DACC[0], sub_mode(p.WinT, (long long)j * p.HH, d.es), H, -(j - 2));
    a.epi.out0 = s.DACC[L];
    if (fuse_bias) a.epi.colsum[0] = gr->start_b;
    RADMMM_TRY(launch_gemm(a, d.mode, mc));
    const ActMat& DH0 = s.DACC[L];
    // 5. start conv: weight / bias gradients on lane 0, input gradients on the chain
    {
        cudaStream_t sd = lane(0);
        RADMMM_TRY(ready(sd));                // dh0 ready
        if (!fuse_bias) RADMMM_TRY(colsum(d.mode, DH0, g, H, 1, 0, gr->start_b, sd));
        float* dWz = s.dW_start;
        float* dWc = s.dW_start + (size_t)H * d.Kz;
        RADMMM_TRY(wgrad(d, f->lens, DH0, w.Z0, H, d.Kz, 1, 1, dWz, d.Kz, 0, sd));
        RADMMM_TRY(wgrad(d, f->lens, DH0, ctx, H, d.Dp, 1, 1, dWc, d.Dp, 0, sd));
        RADMMM_TRY(wn_bwd(dWz, d.Kz, 0, d.Ch, dWc, d.Dp, 0, f->start_v, f->start_g, p.norm_start, H, d.Ch + d.D, 1,
                          gr->start_v, gr->start_g, sd));
    }
    init_args(a, d, f->lens, EPI_DZ0, d.Ch);
    add_seg(a, DH0, p.WzT, H, 0);
    a.epi.cf_out = dz_mid; a.epi.cf_C = d.C; a.epi.cf_c0 = 0; a.epi.accumulate = 1;
    RADMMM_TRY(launch_gemm(a, d.mode, mc));
    init_args(a, d, f->lens, EPI_DCTX, d.Dp);
    add_seg(a, DH0, p.WcT, H, 0);
    a.epi.f32_out = dctx_rows; a.epi.f32_ld = d.Dp;
    RADMMM_TRY(launch_gemm(a, d.mode, mc));
    // 6. invertible 1x1 conv: dz_in = W^T dz_mid, dW = sum dz_mid (z_in - mean)^T over valid frames
    const long long bs = (long long)d.C * d.Tp;
    if (f->W_T) {
        RADMMM_REQUIRE(gr->W != nullptr, "flow_backward: dW output missing");
        cudaStream_t sd = lane(1);
        RADMMM_TRY(ready(sd));                // dz_mid complete: dW of the 1x1 conv goes to a lane
        RADMMM_TRY(inv1x1_wgrad(dz_mid, z_in, f->mean, f->lens, gr->W, d.B, d.C, d.Tp, sd));
        RADMMM_TRY(inv1x1(dz_mid, bs, f->W_T, nullptr, nullptr, dz_in, bs, d.B, d.C, d.C, d.Tp, mc));
    } else {
        RADMMM_REQUIRE(dz_in == dz_mid, "flow_backward: without W_T, dz_in must alias dz_mid");
    }
    // join: everything the caller sees is ordered on `st` (the scratch buffers are shared by all flow steps)
    if (forked) {
        for (int l = 0; l <= kLanes; ++l) {
            cudaEvent_t e = lane_event(n_ev++);
            RADMMM_CUDA(cudaEventRecord(e, l < kLanes ? lane_stream(l) : mc));
            RADMMM_CUDA(cudaStreamWaitEvent(st, e, 0));
        }
    }
    return RADMMM_OK;
}

size_t flow_prepared_bytes(int mode, int C, int D, int H, int L) {
    return layout_prepared(make_dims(mode, 1, C, 1, D, H, L), nullptr, nullptr);
}
size_t flow_workspace_bytes(int mode, int training, int B, int Tp, int C, int D, int H, int L) {
    return layout_workspace(make_dims(mode, B, C, Tp, D, H, L), training, nullptr, nullptr);
}
size_t flow_scratch_bytes(int mode, int B, int Tp, int C, int D, int H, int L) {
    return layout_scratch(make_dims(mode, B, C, Tp, D, H, L), nullptr, nullptr);
}
size_t context_rows_bytes(int mode, int B, int Tp, int D) {
    Dims d = make_dims(mode, B, 2, Tp, D, 128, 1);
    return (size_t)d.planes * d.R * d.Dp * d.es;
}
int context_rows(int mode, const float* ctx_btd, const int* lens, int B, int Tp, int D, void* rows, cudaStream_t st) {
    Dims d = make_dims(mode, B, 2, Tp, D, 128, 1);
    ActMat r; r.ptr = rows; r.ld = d.Dp; r.plane_stride = (long long)d.R * d.Dp;
    return rows_from_btd(mode, ctx_btd, D, geom_of(d, lens), r, d.Dp, st);
}
int context_rows_backward(const float* drows, const int* lens, int B, int Tp, int D, float* dctx, int accumulate,
                          cudaStream_t st) {
    Dims d = make_dims(MODE_F32, B, 2, Tp, D, 128, 1);
    return btd_from_rows(drows, d.Dp, D, geom_of(d, lens), dctx, accumulate, st);
}

}  // namespace radmmm
