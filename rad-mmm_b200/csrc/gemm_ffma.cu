// fp32 FFMA contraction kernel (MODE_F32): the exact-parity path and the checker for the tcgen05 kernel.
// 128x128 output tile, BK=16, 256 threads, 8x8 register micro-tile, double-buffered shared memory.
// Roofline: fp32 FFMA pipe (148 SMs x 128 FMA/clk); this path exists for bit-faithful fp32 parity with the
// reference, not for throughput -- the tensor-core path is gemm_tc.cu.
#include "gemm.cuh"

namespace radmmm {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS = BM + 4, NT = 256;

template <int KIND, bool WGRAD>
__global__ void __launch_bounds__(NT) gemm_ffma_kernel(const GemmArgs args) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM;      // row GEMM: row tile; weight-grad: output-row (dY channel) tile
    const int n0 = blockIdx.y * BN;
    int tap = 0, split = 0;
    if (WGRAD) { tap = blockIdx.z / args.split_k; split = blockIdx.z % args.split_k; }
    const bool acc_segs = WGRAD && args.wgrad == 2;      // all segments accumulate into one output

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float4 ra[2], rb[2];

    const int seg_begin = (WGRAD && !acc_segs) ? tap : 0;
    const int seg_end = (WGRAD && !acc_segs) ? tap + 1 : args.n_seg;
    for (int s = seg_begin; s < seg_end; ++s) {
        const GemmSeg& sg = args.seg[s];
        const float* A = reinterpret_cast<const float*>(sg.a.ptr);
        const float* W = reinterpret_cast<const float*>(sg.w.ptr);
        const long long lda = sg.a.ld, ldw = sg.w.ld;
        int k_begin = 0, k_end = sg.K;
        if (WGRAD) {   // K runs over rows; split-K chunks of whole BK blocks
            int nkb = args.R / BK;
            int per = (nkb + args.split_k - 1) / args.split_k;
            k_begin = split * per * BK;
            k_end = min(args.R, (split + 1) * per * BK);
        }
        const int nk = (k_end - k_begin + BK - 1) / BK;
        if (nk <= 0) continue;

        auto gload = [&](int kb) {
            const int k0 = k_begin + kb * BK;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * NT;
                if (!WGRAD) {
                    const int row = idx >> 2, kq = idx & 3;
                    const int ra_row = m0 + row + sg.shift;
                    ra[i] = (ra_row >= 0 && ra_row < args.R)
                                ? *reinterpret_cast<const float4*>(A + (long long)ra_row * lda + k0 + 4 * kq)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    rb[i] = *reinterpret_cast<const float4*>(W + (long long)(n0 + row) * ldw + k0 + 4 * kq);
                } else {
                    const int k = idx >> 5, q = idx & 31;
                    const int r = k0 + k;
                    const int mc = m0 + 4 * q, nc = n0 + 4 * q;
                    ra[i] = (r < args.R && mc < lda && mc < ((args.epi.M + 3) & ~3)) ? *reinterpret_cast<const float4*>(A + (long long)r * lda + mc)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                    const int rx = r + sg.shift;
                    rb[i] = (rx >= 0 && rx < args.R && nc < ldw && nc < ((args.epi.N + 3) & ~3))
                                ? *reinterpret_cast<const float4*>(W + (long long)rx * ldw + nc)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto sstore = [&](int buf) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = tid + i * NT;
                if (!WGRAD) {
                    const int row = idx >> 2, kq = idx & 3;
                    As[buf][4 * kq + 0][row] = ra[i].x; As[buf][4 * kq + 1][row] = ra[i].y;
                    As[buf][4 * kq + 2][row] = ra[i].z; As[buf][4 * kq + 3][row] = ra[i].w;
                    Bs[buf][4 * kq + 0][row] = rb[i].x; Bs[buf][4 * kq + 1][row] = rb[i].y;
                    Bs[buf][4 * kq + 2][row] = rb[i].z; Bs[buf][4 * kq + 3][row] = rb[i].w;
                } else {
                    const int k = idx >> 5, q = idx & 31;
                    *reinterpret_cast<float4*>(&As[buf][k][4 * q]) = ra[i];
                    *reinterpret_cast<float4*>(&Bs[buf][k][4 * q]) = rb[i];
                }
            }
        };

        gload(0);
        sstore(0);
        __syncthreads();
        for (int kb = 0; kb < nk; ++kb) {
            const int buf = kb & 1;
            if (kb + 1 < nk) gload(kb + 1);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[8], b[8];
                *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
                *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
                *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
                *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (kb + 1 < nk) sstore(buf ^ 1);
            __syncthreads();
        }
    }

    if (WGRAD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) epi_wgrad<8>(args.epi, tap, m0 + ty * 8 + i, n0 + tx * 8, acc[i]);
    } else {
        const Stager none{nullptr, 0};
#pragma unroll
        for (int i = 0; i < 8; ++i) epi_apply<MODE_F32, KIND, 8>(args.epi, none, m0 + ty * 8 + i, n0 + tx * 8, acc[i]);
    }
}

template <int KIND>
int launch_kind(const GemmArgs& args, cudaStream_t stream) {
    dim3 grid(args.R / BM, cdiv(args.epi.N, BN), 1);
    gemm_ffma_kernel<KIND, false><<<grid, NT, 0, stream>>>(args);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace

int launch_gemm_ffma(const GemmArgs& args, cudaStream_t stream) {
    RADMMM_REQUIRE(args.R % BM == 0, "gemm_ffma: R=%d must be a multiple of %d", args.R, BM);
    RADMMM_REQUIRE(args.n_seg >= 1 && args.n_seg <= kMaxSeg, "gemm_ffma: bad segment count %d", args.n_seg);
    for (int s = 0; s < args.n_seg; ++s)
        RADMMM_REQUIRE(args.wgrad || args.seg[s].K % BK == 0, "gemm_ffma: K=%d must be a multiple of %d", args.seg[s].K, BK);
    if (args.wgrad) {
        GemmArgs a2 = args;
        if (a2.split_k < 1) {      // auto: aim for >= ~300 CTAs
            const long long tiles = (long long)cdiv(args.epi.M, BM) * cdiv(args.epi.N, BN) * (args.wgrad == 2 ? 1 : args.n_seg);
            int split = tiles < 256 ? (int)((296 + tiles - 1) / tiles) : 1;
            if (split > 16) split = 16;
            if (split > args.R / 128) split = args.R / 128;
            a2.split_k = split < 1 ? 1 : split;
        }
        RADMMM_REQUIRE(a2.split_k == 1 || a2.epi.atomic, "gemm_ffma: split-K needs the atomic epilogue");
        if (args.zero_output) RADMMM_TRY(zero_wgrad_output(args, stream));
        dim3 grid(cdiv(args.epi.M, BM), cdiv(args.epi.N, BN), (args.wgrad == 2 ? 1 : args.n_seg) * a2.split_k);
        gemm_ffma_kernel<EPI_WGRAD, true><<<grid, NT, 0, stream>>>(a2);
        RADMMM_LAUNCH_CHECK();
        return RADMMM_OK;
    }
    switch (args.epi.kind) {
        case EPI_START: return launch_kind<EPI_START>(args, stream);
        case EPI_IN: return launch_kind<EPI_IN>(args, stream);
        case EPI_RS: return launch_kind<EPI_RS>(args, stream);
        case EPI_END: return launch_kind<EPI_END>(args, stream);
        case EPI_DOUT: return launch_kind<EPI_DOUT>(args, stream);
        case EPI_DH: return launch_kind<EPI_DH>(args, stream);
        case EPI_DH0: return launch_kind<EPI_DH0>(args, stream);
        case EPI_DZ0: return launch_kind<EPI_DZ0>(args, stream);
        case EPI_DCTX: return launch_kind<EPI_DCTX>(args, stream);
        case EPI_F32: return launch_kind<EPI_F32>(args, stream);
    }
    set_error("gemm_ffma: unknown epilogue kind %d", args.epi.kind);
    return RADMMM_ERR_ARG;
}

}  // namespace radmmm
