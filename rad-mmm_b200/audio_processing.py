"""STFT / mel front end with the reference's class names (audio_processing.py:116-154, 192-255).

``TacotronSTFT.mel_spectrogram(y)`` keeps the reference's contract -- y (B, S) in [-1, 1] -> (B, n_mel, S//hop + 1)
log-mel -- but runs one fused kernel (FFT in shared memory, magnitude, mel filterbank, log-clamp) instead of a dense
DFT convolution + matmul.  The mel basis is the librosa-0.8.0 Slaney filterbank (the reference calls
``librosa.filters.mel``, audio_processing.py:124-125; librosa is not a dependency here, the published algorithm is
restated in :func:`mel_filterbank`).  Importing this module does not touch CUDA, so DataLoader workers can import it.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _native as N


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filters: float32 (n_mels, n_fft//2 + 1)."""
    fmax = sr / 2.0 if fmax is None else fmax
    n_bins = 1 + n_fft // 2
    freqs = np.linspace(0, sr / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - freqs[None, :]
    fb = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        fb[i] = np.maximum(0, np.minimum(-ramps[i] / width[i], ramps[i + 2] / width[i + 1]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    return torch.exp(x) / C


class STFT(torch.nn.Module):
    """audio_processing.py:192-255 (forward transform only; the inverse belongs to the vocoder's denoiser)."""

    def __init__(self, filter_length=800, hop_length=200, win_length=800, window="hann"):
        super().__init__()
        if window != "hann" or win_length != filter_length:
            raise NotImplementedError("radmmm_b200.STFT builds the shipped configuration: periodic Hann window with "
                                      "win_length == filter_length")
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window

    def transform(self, input_data):
        """(B, S) -> (magnitude (B, n_fft/2+1, frames), None).  The reference's unused phase is not computed."""
        lib = N.lib()
        y = input_data.contiguous().float()
        b, s = y.shape
        n_frames = s // self.hop_length + 1
        n_bins = self.filter_length // 2 + 1
        mag = torch.empty(b, n_bins, n_frames, device=y.device)
        dummy_basis = torch.zeros(1, n_bins, device=y.device)
        dummy_mel = torch.empty(b, 1, n_frames, device=y.device)
        N.check(lib.radmmm_stft_mel(N.fptr(y), N.fptr(dummy_basis), N.fptr(dummy_mel), N.fptr(mag), b, s,
                                    self.filter_length, self.hop_length, 1, 1e-5, N.stream()))
        return mag, None


class TacotronSTFT(torch.nn.Module):
    """audio_processing.py:116-154."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=None):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length)
        self.register_buffer("mel_basis", torch.from_numpy(
            mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)).float())

    def spectral_normalize(self, magnitudes):
        return dynamic_range_compression(magnitudes)

    def spectral_de_normalize(self, magnitudes):
        return dynamic_range_decompression(magnitudes)

    def mel_spectrogram(self, y, check_range: bool = True):
        """y (B, S) in [-1, 1] -> (B, n_mel_channels, S // hop + 1).  ``check_range=False`` skips the two asserts of
        audio_processing.py:147-148 (each is a device sync)."""
        if check_range:
            assert torch.min(y.data) >= -1
            assert torch.max(y.data) <= 1
        lib = N.lib()
        y = y.contiguous().float()
        b, s = y.shape
        st = self.stft_fn
        mel = torch.empty(b, self.n_mel_channels, s // st.hop_length + 1, device=y.device)
        N.check(lib.radmmm_stft_mel_sparse(N.fptr(y), N.fptr(self.mel_basis), N.ptr(self._support()), N.fptr(mel), None, b, s,
                                           st.filter_length, st.hop_length, self.n_mel_channels, 1e-5, N.stream()))
        return mel

    def _support(self) -> torch.Tensor:
        """[first, last] non-zero bin of every mel-basis row (int32, on the basis' device); recomputed when the buffer is
        replaced or modified (load_state_dict, .to())."""
        key = (self.mel_basis.data_ptr(), self.mel_basis._version)
        if getattr(self, "_support_key", None) != key:
            sup = torch.empty(self.n_mel_channels, 2, dtype=torch.int32, device=self.mel_basis.device)
            N.check(N.lib().radmmm_mel_support(N.fptr(self.mel_basis), self.n_mel_channels, self.mel_basis.shape[1], N.ptr(sup),
                                               N.stream()))
            self._support_buf, self._support_key = sup, key
        return self._support_buf


class BatchedFrontEnd(torch.nn.Module):
    """The data-path front end as one batched GPU call: what the reference's ``Data.get_mel`` +
    ``Data.get_energy_average`` do per utterance on a CPU worker (data.py:363-376: ``audio / max_wav_value`` ->
    ``TacotronSTFT.mel_spectrogram`` -> optional mel noise; energy = mean over mel channels, ``(x + 20) / 20`` when
    ``use_scaled_energy``), for a padded batch of raw audio in the training process.

    ``forward(audio, audio_lens)``: audio (B, S_max) raw samples (int16 range when ``max_wav_value`` = 32768), audio_lens (B)
    -> mel (B, n_mel, T_max), energy_avg (B, T_max), out_lens (B) with ``out_lens = audio_lens // hop + 1`` -- the frame
    count the per-utterance call produces.  Frames at or beyond ``out_lens`` are zero.  Frames whose analysis window reaches
    past the end of the utterance see the reflect padding of the UTTERANCE in the reference and the batch padding (zeros)
    here; ``exact_tail=True`` (default) re-computes those last ``n_fft / (2 hop) + 1`` frames from a short clip ending at the
    utterance's end, so the result is the per-utterance one everywhere (one tiny extra launch per utterance; ``audio_lens``
    is read on the host)."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0, max_wav_value=32768.0, use_scaled_energy=True, mel_noise_scale=0.0,
                 exact_tail=True):
        super().__init__()
        self.stft = TacotronSTFT(filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax)
        self.max_wav_value = max_wav_value
        self.use_scaled_energy = use_scaled_energy
        self.mel_noise_scale = mel_noise_scale
        self.exact_tail = exact_tail
        self.hop, self.n_fft = hop_length, filter_length

    def energy_avg_normalize(self, x):
        return (x + 20.0) / 20.0 if self.use_scaled_energy else x          # data.py:339-342

    @torch.no_grad()
    def forward(self, audio, audio_lens):
        audio = audio.float() / self.max_wav_value
        b, s_max = audio.shape
        lens = torch.as_tensor(audio_lens, device=audio.device).long()
        out_lens = torch.div(lens, self.hop, rounding_mode="floor") + 1
        mel = self.stft.mel_spectrogram(audio.clamp(-1, 1), check_range=False)
        if self.exact_tail:
            # The last frames of an utterance (window centre within n_fft/2 of its end) see the utterance's own reflection in
            # the reference and the batch padding here.  Re-compute them from a short clip that ends exactly at the utterance
            # end and starts on a hop boundary far enough back that the clip's own leading reflection is not used.
            half, hop = self.n_fft // 2, self.hop
            for i, n in enumerate(lens.tolist()):          # host lengths (the loader knows them); B tiny launches
                if n >= s_max or n <= half:
                    continue                               # the longest utterance is exact already; tiny clips: left as is
                f_first = max(0, (n - half) // hop)
                start = max(0, (f_first - half // hop - 1) * hop)
                clip = audio[i:i + 1, start:n].clamp(-1, 1)
                tm = self.stft.mel_spectrogram(clip, check_range=False)[0]
                f0 = f_first - start // hop
                n_fix = n // hop + 1 - f_first
                mel[i, :, f_first:f_first + n_fix] = tm[:, f0:f0 + n_fix]
        t = torch.arange(mel.shape[2], device=mel.device)[None]
        keep = (t < out_lens[:, None])
        if self.mel_noise_scale > 0:
            mel = mel + torch.randn_like(mel) * self.mel_noise_scale
        mel = mel * keep[:, None]
        energy = self.energy_avg_normalize(mel.mean(1)) * keep              # data.py:358-361
        return mel, energy, out_lens
