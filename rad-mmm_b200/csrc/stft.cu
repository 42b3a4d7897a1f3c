// Fused STFT -> magnitude -> mel filterbank -> log-clamp (audio_processing.py:137-154, 227-255).
// The reference computes the STFT as a dense windowed-DFT convolution (2.1 MFLOP/frame); here each CTA runs
// in-shared-memory radix-2 FFTs for FR consecutive frames (reflect padding and the periodic Hann window applied on
// load, bit-reversed store), forms |X| in place and applies the (n_mel x n_bins) mel basis once for all FR frames so
// each basis element is read once per CTA.  HBM-bound by design: 256 new samples in + 80 values out per frame
// (1344 B); the audio tile of a CTA is read once (overlapping frames are served from L1/L2).
#include "common.cuh"
#include "ops.cuh"

namespace radmmm {

namespace {

constexpr int FR = 4;          // frames per CTA
constexpr int NT = 256;

__device__ __forceinline__ int reflect_index(int j, int S) {
    if (j < 0) j = -j;
    if (j >= S) j = 2 * (S - 1) - j;
    return j;
}

template <int NFFT, int LOG2N>
__global__ void __launch_bounds__(NT) stft_mel_kernel(const float* __restrict__ audio, const float* __restrict__ mel_basis,
                                                      float* __restrict__ mel, float* __restrict__ mag_out, int S,
                                                      int hop, int n_frames, int n_mel, float clip) {
    constexpr int NBINS = NFFT / 2 + 1;
    __shared__ float2 buf[FR][NFFT];
    __shared__ float2 tw[NFFT / 2];
    const int b = blockIdx.y, f0 = blockIdx.x * FR, tid = threadIdx.x;
    const float* x = audio + (long long)b * S;
    for (int i = tid; i < NFFT / 2; i += NT) {
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)NFFT, &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    // load + window + bit-reverse
    for (int fr = 0; fr < FR; ++fr) {
        const int f = f0 + fr;
        for (int i = tid; i < NFFT; i += NT) {
            float v = 0.0f;
            if (f < n_frames) {
                const int j = reflect_index(f * hop + i - NFFT / 2, S);
                const float win = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)NFFT);   // periodic Hann
                v = x[j] * win;
            }
            buf[fr][__brev((unsigned)i) >> (32 - LOG2N)] = make_float2(v, 0.0f);
        }
    }
    __syncthreads();
    // iterative radix-2 decimation-in-time
    for (int s = 1; s <= LOG2N; ++s) {
        const int half = 1 << (s - 1);
        for (int idx = tid; idx < FR * (NFFT / 2); idx += NT) {
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
    // magnitude (into .x of the first NBINS entries)
    for (int idx = tid; idx < FR * NBINS; idx += NT) {
        const int fr = idx / NBINS, k = idx % NBINS;
        const float2 v = buf[fr][k];
        const float m = sqrtf(v.x * v.x + v.y * v.y);
        buf[fr][k].x = m;
        if (mag_out && f0 + fr < n_frames) mag_out[((long long)b * NBINS + k) * n_frames + f0 + fr] = m;
    }
    __syncthreads();
    // mel: warp w handles rows w, w+8, ...
    const int lane = tid & 31, wid = tid >> 5;
    for (int m = wid; m < n_mel; m += NT / 32) {
        const float* row = mel_basis + (long long)m * NBINS;
        float acc[FR];
#pragma unroll
        for (int fr = 0; fr < FR; ++fr) acc[fr] = 0.0f;
        for (int k = lane; k < NBINS; k += 32) {
            const float wgt = __ldg(row + k);
            if (wgt != 0.0f) {
#pragma unroll
                for (int fr = 0; fr < FR; ++fr) acc[fr] = fmaf(wgt, buf[fr][k].x, acc[fr]);
            }
        }
#pragma unroll
        for (int fr = 0; fr < FR; ++fr) {
            const float v = warp_sum(acc[fr]);
            if (lane == 0 && f0 + fr < n_frames) mel[((long long)b * n_mel + m) * n_frames + f0 + fr] = logf(fmaxf(v, clip));
        }
    }
}

}  // namespace

int stft_mel(const float* audio, const float* mel_basis, float* mel, float* mag, int B, int S, int n_fft, int hop,
             int n_mel, float clip, cudaStream_t st) {
    RADMMM_REQUIRE(n_fft == 1024, "stft_mel: only n_fft=1024 is built (got %d)", n_fft);
    RADMMM_REQUIRE(S > n_fft / 2, "stft_mel: reflect padding needs more than n_fft/2 samples (S=%d)", S);
    RADMMM_REQUIRE(hop > 0 && B > 0 && n_mel > 0, "stft_mel: bad sizes");
    const int n_frames = S / hop + 1;
    dim3 grid(cdiv(n_frames, FR), B);
    stft_mel_kernel<1024, 10><<<grid, NT, 0, st>>>(audio, mel_basis, mel, mag, S, hop, n_frames, n_mel, clip);
    RADMMM_LAUNCH_CHECK();
    return RADMMM_OK;
}

}  // namespace radmmm
