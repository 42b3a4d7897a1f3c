"""Oracle (tests only): STFT + mel front end and the soft attention.

STFT / mel: audio_processing.py:116-154 (TacotronSTFT), :192-255 (STFT.__init__/transform).
The mel filterbank comes from ``librosa.filters.mel`` -- librosa==0.8.0 is pinned in the reference's
requirements.txt:92 but is neither vendored nor installed here, so :func:`slaney_mel_basis` restates its
published algorithm (Slaney-scale triangular filters, ``norm='slaney'``, ``htk=False``).  PARITY UNPINNED at
that boundary (no reference test or fixture holds a mel basis); it is cross-checked against
``torchaudio.functional.melscale_fbanks(..., norm='slaney', mel_scale='slaney')`` in the tests.

Soft attention: ConvAttention.forward, common.py:1239-1277; bmm at tts_lightning_modules.py:670.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .flow import weight_norm_weight

Tensor = torch.Tensor


# --------------------------------------------------------------------------- mel basis (librosa 0.8.0 restated)
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr: float, n_fft: int, n_mels: int, fmin: float, fmax: Optional[float]) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` of librosa 0.8.0 -> float32 (n_mels, 1+n_fft//2)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, sr / 2.0, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(n: int) -> np.ndarray:
    """``scipy.signal.get_window('hann', n, fftbins=True)`` (audio_processing.py:216)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def dft_basis(n_fft: int) -> np.ndarray:
    """float32 (2*(n_fft/2+1), n_fft): [Re; Im] of FFT(I)[:cutoff] times the Hann window (audio_processing.py:203-222)."""
    fb = np.fft.fft(np.eye(n_fft))
    cutoff = n_fft // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])]).astype(np.float32)
    return fb * hann_periodic(n_fft).astype(np.float32)[None, :]


def stft_magnitude(y: Tensor, n_fft: int = 1024, hop: int = 256, dense: bool = True) -> Tensor:
    """STFT.transform magnitude (audio_processing.py:227-251): reflect pad n_fft/2, frames at ``hop``.

    ``dense=True`` follows the reference literally (fp32 conv with the windowed DFT basis);
    ``dense=False`` uses an fp64 rFFT -- the higher-precision truth the kernel is also compared to."""
    pad = n_fft // 2
    x = F.pad(y[:, None, :], (pad, pad), mode="reflect")
    if dense:
        basis = torch.from_numpy(dft_basis(n_fft))[:, None, :].to(y.dtype)
        ft = F.conv1d(x, basis, stride=hop)
        cutoff = n_fft // 2 + 1
        return torch.sqrt(ft[:, :cutoff] ** 2 + ft[:, cutoff:] ** 2)
    frames = x[:, 0].double().unfold(1, n_fft, hop)                       # (B, n_frames, n_fft)
    win = torch.from_numpy(hann_periodic(n_fft))
    return torch.fft.rfft(frames * win, dim=-1).abs().transpose(1, 2)


def mel_spectrogram(y: Tensor, sr: float = 22050, n_fft: int = 1024, hop: int = 256, n_mels: int = 80,
                    fmin: float = 0.0, fmax: Optional[float] = 8000.0, dense: bool = True) -> Tensor:
    """TacotronSTFT.mel_spectrogram (audio_processing.py:137-154): log(clamp(mel_basis @ |STFT|, 1e-5))."""
    assert float(y.min()) >= -1 and float(y.max()) <= 1
    mag = stft_magnitude(y, n_fft, hop, dense)
    basis = torch.from_numpy(slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)).to(mag.dtype)
    return torch.log(torch.clamp(torch.matmul(basis, mag), min=1e-5))


# --------------------------------------------------------------------------- soft attention
def _convnorm(sd, pre: str, x: Tensor, ksize: int) -> Tensor:
    """ConvNorm without partial padding (common.py:152-191): weight-normed Conv1d, 'same' padding."""
    w = weight_norm_weight(sd[pre + "conv.weight_g"], sd[pre + "conv.weight_v"])
    return F.conv1d(x, w, sd[pre + "conv.bias"], padding=(ksize - 1) // 2)


def attention_projections(sd: Dict[str, Tensor], pre: str, queries: Tensor, keys: Tensor):
    """key_proj / query_proj stacks, common.py:1196-1210,1254-1257."""
    k = _convnorm(sd, pre + "key_proj.2.", torch.relu(_convnorm(sd, pre + "key_proj.0.", keys, 3)), 1)
    q = torch.relu(_convnorm(sd, pre + "query_proj.0.", queries, 3))
    q = torch.relu(_convnorm(sd, pre + "query_proj.2.", q, 1))
    q = _convnorm(sd, pre + "query_proj.4.", q, 1)
    return q, k


def soft_attention(q: Tensor, k: Tensor, in_lens: Tensor, prior: Optional[Tensor] = None,
                   temp: float = 0.0005):
    """common.py:1259-1276.  q (B,C,T1), k (B,C,T2), prior (B,T1,T2).  Returns (attn (B,1,T1,T2), attn_logprob)."""
    attn = -temp * ((q[:, :, :, None] - k[:, :, None, :]) ** 2).sum(1, keepdim=True)
    if prior is not None:
        attn = F.log_softmax(attn, dim=3) + torch.log(prior[:, None] + 1e-8)
    logprob = attn.clone()
    t2 = k.shape[2]
    pad = torch.arange(t2, device=k.device)[None, :] >= in_lens[:, None]            # (B,T2) True = padded token
    attn = attn.masked_fill(pad[:, None, None, :], -float("inf"))
    return torch.softmax(attn, dim=3), logprob


def attend(txt_enc: Tensor, attn: Tensor) -> Tensor:
    """tts_lightning_modules.py:670: context = bmm(txt_enc (B,D,T2), attn^T) -> (B,D,T1)."""
    return torch.bmm(txt_enc, attn.squeeze(1).transpose(1, 2))
