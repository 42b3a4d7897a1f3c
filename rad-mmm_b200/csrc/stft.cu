// Fused STFT -> magnitude -> mel filterbank -> log-clamp (audio_processing.py:137-154, 227-255).
// The reference computes the STFT as a dense windowed-DFT convolution (2.1 MFLOP/frame); here each CTA runs
// in-shared-memory radix-2 FFTs for FR consecutive frames, two real frames per complex transform (reflect padding and the
// periodic Hann window applied on load, bit-reversed store), splits the packed spectra into |X| per frame and applies the (n_mel x n_bins) mel basis once for all FR frames so
// each basis element is read once per CTA.  HBM-bound by design: 256 new samples in + 80 values out per frame
// (1344 B); the audio tile of a CTA is read once (overlapping frames are served from L1/L2).
// The mel basis is triangular: row m is non-zero on one short run of bins (727 non-zeros of 80 x 513 for the shipped
// configuration).  `mel_support_kernel` finds every row's [first, last] non-zero bin once per basis; with that table a warp
// reads ~9 weights per mel row instead of scanning 513 (the dense scan made every 4-frame CTA pull 164 KB through L2 for
// 5.4 KB of algorithmic bytes: 225 us per 6400 frames).  8 frames per CTA: each (mel row, CTA) is one 32-byte store.
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int FR = 8;          // frames per CTA: 4 complex FFTs of two real frames each (32 KB) + 16 KB of magnitudes
constexpr int NT = 256;

__device__ __forceinline__ int reflect_index(int j, int S) {
    if (j < 0) j = -j;
    if (j >= S) j = 2 * (S - 1) - j;
    return j;
}

template <int NFFT, int LOG2N>
__global__ void __launch_bounds__(NT) stft_mel_kernel(const float* __restrict__ audio, const float* __restrict__ mel_basis,
                                                      const int2* __restrict__ support, float* __restrict__ mel,
                                                      float* __restrict__ mag_out, int S, int hop, int n_frames, int n_mel,
                                                      float clip) {
    constexpr int NBINS = NFFT / 2 + 1;
    constexpr int NP = FR / 2;                       // complex transforms per CTA: two REAL frames ride one complex FFT
    extern __shared__ float2 stft_smem[];
    float2 (*buf)[NFFT] = reinterpret_cast<float2 (*)[NFFT]>(stft_smem);          // [NP][NFFT]: (frame 2p, frame 2p+1)
    float2* tw = stft_smem + NP * NFFT;                                           // [NFFT / 2]
    float (*mag)[NBINS + 3] = reinterpret_cast<float (*)[NBINS + 3]>(tw + NFFT / 2);   // [FR][NBINS + 3]
    const int b = blockIdx.y, f0 = blockIdx.x * FR, tid = threadIdx.x;
    const float* x = audio + (long long)b * S;
    for (int i = tid; i < NFFT / 2; i += NT) {
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)NFFT, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    // load + window + bit-reverse; frame 2p goes to the real part, frame 2p+1 to the imaginary part
    for (int i = tid; i < NFFT; i += NT) {
        const float win = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)NFFT);   // periodic Hann
        const int dst = __brev((unsigned)i) >> (32 - LOG2N);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int fa = f0 + 2 * p, fb = fa + 1;
            const float va = (fa < n_frames) ? x[reflect_index(fa * hop + i - NFFT / 2, S)] * win : 0.0f;
            const float vb = (fb < n_frames) ? x[reflect_index(fb * hop + i - NFFT / 2, S)] * win : 0.0f;
            buf[p][dst] = make_float2(va, vb);
        }
    }
    __syncthreads();
    // iterative radix-2 decimation-in-time
    for (int s = 1; s <= LOG2N; ++s) {
        const int half = 1 << (s - 1);
        for (int idx = tid; idx < NP * (NFFT / 2); idx += NT) {
            const int fr = idx / (NFFT / 2), j = idx % (NFFT / 2);
            const int k = j & (half - 1);
            const int i0 = ((j >> (s - 1)) << s) + k, i1 = i0 + half;
            const float2 w = tw[k << (LOG2N - s)];
            const float2 a = buf[fr][i0], c = buf[fr][i1];
            const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
            buf[fr][i0] = make_float2(a.x + t.x, a.y + t.y);
            buf[fr][i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    // split the packed spectrum Z = A + iB (A, B spectra of the two real frames):  A[k] = (Z[k] + conj Z[N-k]) / 2,
    // B[k] = (Z[k] - conj Z[N-k]) / (2i); magnitudes of both
    for (int idx = tid; idx < NP * NBINS; idx += NT) {
        const int p = idx / NBINS, k = idx % NBINS;
        const float2 z = buf[p][k], zc = buf[p][(NFFT - k) & (NFFT - 1)];
        const float ar = 0.5f * (z.x + zc.x), ai = 0.5f * (z.y - zc.y);
        const float br = 0.5f * (z.y + zc.y), bi = -0.5f * (z.x - zc.x);
        const float ma = sqrtf(ar * ar + ai * ai), mb = sqrtf(br * br + bi * bi);
        mag[2 * p][k] = ma;
        mag[2 * p + 1][k] = mb;
        if (mag_out) {
            if (f0 + 2 * p < n_frames) mag_out[((long long)b * NBINS + k) * n_frames + f0 + 2 * p] = ma;
            if (f0 + 2 * p + 1 < n_frames) mag_out[((long long)b * NBINS + k) * n_frames + f0 + 2 * p + 1] = mb;
        }
    }
    __syncthreads();
    // mel: warp w handles rows w, w+8, ...
    const int lane = tid & 31, wid = tid >> 5;
    for (int m = wid; m < n_mel; m += NT / 32) {
        const float* row = mel_basis + (long long)m * NBINS;
        float acc[FR];
#pragma unroll
        for (int fr = 0; fr < FR; ++fr) acc[fr] = 0.0f;
        int k_lo = 0, k_hi = NBINS - 1;
        if (support != nullptr) { const int2 sp = support[m]; k_lo = sp.x; k_hi = sp.y; }
        for (int k = k_lo + lane; k <= k_hi; k += 32) {
            const float wgt = __ldg(row + k);
            if (wgt != 0.0f) {
#pragma unroll
                for (int fr = 0; fr < FR; ++fr) acc[fr] = fmaf(wgt, mag[fr][k], acc[fr]);
            }
        }
        // lane fr ends up holding frame fr's sum: FR consecutive frames of one mel row leave as one 32-byte segment
        float mine = 0.0f;
#pragma unroll
        for (int fr = 0; fr < FR; ++fr) {
            const float v = warp_sum(acc[fr]);
            if (lane == fr) mine = v;
        }
        if (lane < FR && f0 + lane < n_frames) mel[((long long)b * n_mel + m) * n_frames + f0 + lane] = logf(fmaxf(mine, clip));
    }
}

// [first, last] non-zero bin of every mel-basis row (an all-zero row gets the empty range [1, 0])
__global__ void mel_support_kernel(const float* __restrict__ mel_basis, int n_mel, int n_bins, int2* __restrict__ support) {
    const int m = blockIdx.x, lane = threadIdx.x;
    int lo = n_bins, hi = -1;
    for (int k = lane; k < n_bins; k += 32)
        if (mel_basis[(long long)m * n_bins + k] != 0.0f) { lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) support[m] = hi < 0 ? make_int2(1, 0) : make_int2(lo, hi);
}

}  // namespace

int mel_support(const float* mel_basis, int n_mel, int n_bins, int* support, cudaStream_t st) {
    RADMMM_REQUIRE(mel_basis && support && n_mel > 0 && n_bins > 0, "mel_support: bad arguments");
    mel_support_kernel<<<n_mel, 32, 0, st>>>(mel_basis, n_mel, n_bins, reinterpret_cast<int2*>(support));
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

int stft_mel(const float* audio, const float* mel_basis, const int* support, float* mel, float* mag, int B, int S, int n_fft,
             int hop, int n_mel, float clip, cudaStream_t st) {
    RADMMM_REQUIRE(n_fft == 1024, "stft_mel: only n_fft=1024 is built (got %d)", n_fft);
    RADMMM_REQUIRE(S > n_fft / 2, "stft_mel: reflect padding needs more than n_fft/2 samples (S=%d)", S);
    RADMMM_REQUIRE(hop > 0 && B > 0 && n_mel > 0, "stft_mel: bad sizes");
    const int n_frames = S / hop + 1;
    dim3 grid(cdiv(n_frames, FR), B);
    constexpr size_t smem = sizeof(float2) * ((FR / 2) * 1024 + 512) + sizeof(float) * FR * (513 + 3);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        RADMMM_CUDA(cudaFuncSetAttribute(stft_mel_kernel<1024, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    stft_mel_kernel<1024, 10><<<grid, NT, smem, st>>>(audio, mel_basis, reinterpret_cast<const int2*>(support), mel, mag, S, hop,
                                                      n_frames, n_mel, clip);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
