"""Text encoder and attribute-predictor backbone without the per-utterance Python loops.

Reference: ``common.Encoder`` (common.py:423-500) and ``common.ConvLSTMLinear`` (common.py:240-330).  Both run their conv
stacks one utterance at a time (``for b_ind in range(B): ... x[b_ind:b_ind+1, :, :len]``, "TODO: speed up") so that every
utterance sees its own zero padding / partial-convolution edge ratio / instance-norm statistics, then pack the results for a
bi-LSTM.  Here the same arithmetic runs on the padded batch: length-masked inputs, the partial-conv ratio in closed form from
the lengths, masked instance-norm statistics, and the packed bi-LSTM as this package's ragged LSTM kernels
(``lstm.context_lstm``) -- B times fewer launches, no host sync on the lengths.  The convolutions themselves stay
``torch.nn.functional.conv1d`` ("stays PyTorch").

Same constructor arguments, attribute names and ``state_dict`` keys as the reference classes (weight-normed ``ConvNorm``
under ``.conv``, ``InstanceNorm1d`` affine parameters, ``nn.LSTM`` with ``torch.nn.utils.spectral_norm`` / ``weight_norm`` on
``weight_hh_l0(_reverse)``), so reference checkpoints load unchanged.  Dropout draws one mask for the padded batch instead of
one per utterance (same distribution, different stream); everything else is value-identical to the per-utterance loop.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .common import SequenceLength, _ConvNormHolder
from .lstm import context_lstm


def _lens_tensor(lens, device) -> torch.Tensor:
    if isinstance(lens, SequenceLength):
        lens = lens.lengths
    return torch.as_tensor(lens).to(device=device, dtype=torch.long)


def _weight(conv) -> torch.Tensor:
    """Effective weight of a weight-normed conv holder (nn.utils.weight_norm, dim=0)."""
    g, v = conv.weight_g, conv.weight_v
    return v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1))


def _tap_ratio(lens: torch.Tensor, t_max: int, ksize: int) -> torch.Tensor:
    """PartialConv1d's mask ratio for an all-ones mask over each utterance's own length (partialconv1d.py:74-80):
    ``ksize / (number of taps inside [0, len) + 1e-6)`` at frame t < len, 0 beyond."""
    t = torch.arange(t_max, device=lens.device)[None, :]
    half = (ksize - 1) // 2
    lo = torch.clamp(t - half, min=0)
    hi = torch.minimum(t + half, lens[:, None] - 1)
    u = (hi - lo + 1).clamp(min=0).float()
    valid = (t < lens[:, None]).float()
    return (ksize / (u + 1e-6)) * valid


def _masked_instance_norm(y: torch.Tensor, mask: torch.Tensor, lens: torch.Tensor, inorm: nn.InstanceNorm1d) -> torch.Tensor:
    """InstanceNorm1d(affine=True, no running stats) of every utterance over its own frames (biased variance)."""
    n = lens.clamp(min=1).float()[:, None, None]
    mean = (y * mask).sum(2, keepdim=True) / n
    var = (((y - mean) * mask) ** 2).sum(2, keepdim=True) / n
    out = (y - mean) * torch.rsqrt(var + inorm.eps)
    if inorm.affine:
        out = out * inorm.weight[None, :, None] + inorm.bias[None, :, None]
    return out


def _run_norm_hooks(lstm: nn.LSTM):
    """spectral_norm / weight_norm recompute ``weight_hh_l0(_reverse)`` in a forward pre-hook of the LSTM; the recurrence runs
    in this package's kernels, not in ``lstm.forward``, so fire the hooks by hand (training mode: one power iteration, as the
    reference's call would do)."""
    for hook in lstm._forward_pre_hooks.values():
        hook(lstm, None)


def _packed_bilstm(lstm: nn.LSTM, x_btd: torch.Tensor, lens: torch.Tensor, precision: str) -> torch.Tensor:
    _run_norm_hooks(lstm)
    if lstm.bidirectional:
        return context_lstm(lstm, x_btd.contiguous(), lens, precision)
    raise NotImplementedError("radmmm_b200.encoders: unidirectional LSTM backbones (lstm_type='lstm') are not covered by the "
                              "ragged LSTM kernels; the shipped configs use 'bilstm'")


class Encoder(nn.Module):
    """common.py:423-500: three partial-padding ConvNorm(k=5) + InstanceNorm + ReLU + dropout(0.5) banks, then a packed
    bi-LSTM.  ``forward(x, in_lens)``: x (B, C, L) padded text embeddings -> (B, L, C), zero beyond each length.  (The
    reference returns ``max(in_lens)`` frames; collated batches have ``L == max(in_lens)``.)"""

    def __init__(self, encoder_n_convolutions=3, encoder_embedding_dim=512, encoder_kernel_size=5, lstm_norm_fn=None):
        super().__init__()
        self.kernel_size = encoder_kernel_size
        convolutions = []
        for _ in range(encoder_n_convolutions):
            convolutions.append(nn.Sequential(_ConvNormHolder(encoder_embedding_dim, encoder_embedding_dim, encoder_kernel_size),
                                              nn.InstanceNorm1d(encoder_embedding_dim, affine=True)))
        self.convolutions = nn.ModuleList(convolutions)
        self.lstm = nn.LSTM(encoder_embedding_dim, int(encoder_embedding_dim / 2), 1, batch_first=True, bidirectional=True)
        if lstm_norm_fn is not None:
            fn = torch.nn.utils.spectral_norm if "spectral" in lstm_norm_fn else torch.nn.utils.weight_norm
            self.lstm = fn(self.lstm, "weight_hh_l0")
            self.lstm = fn(self.lstm, "weight_hh_l0_reverse")
        self.precision = "fp32"          # the reference runs the encoder under autocast(False)

    def _convs(self, x, lens):
        t_max = x.shape[2]
        mask = (torch.arange(t_max, device=x.device)[None, :] < lens[:, None]).to(x.dtype)[:, None]
        ratio = _tap_ratio(lens, t_max, self.kernel_size)[:, None]
        for seq in self.convolutions:
            conv, inorm = seq[0].conv, seq[1]
            raw = F.conv1d(x * mask, _weight(conv), None, padding=(self.kernel_size - 1) // 2)
            y = raw * ratio + conv.bias[None, :, None]                   # (conv + b - b) * ratio + b, partialconv1d.py:88-91
            y = _masked_instance_norm(y, mask, lens, inorm)
            x = F.dropout(F.relu(y), 0.5, self.training) * mask
        return x

    def forward(self, x, in_lens):
        lens = _lens_tensor(in_lens, x.device)
        with torch.autocast("cuda", enabled=False):
            x = self._convs(x.float(), lens)
            return _packed_bilstm(self.lstm, x.transpose(1, 2), lens, self.precision)

    def infer(self, x):
        lens = torch.full((x.shape[0],), x.shape[2], device=x.device, dtype=torch.long)
        return self.forward(x, lens)


class ConvLSTMLinear(nn.Module):
    """common.py:240-330 (attribute-predictor backbone): ``n_layers`` x [ConvNorm(k) + ReLU + dropout] on every utterance's own
    zero padding, packed bi-LSTM (spectral norm on the recurrent weights), linear head.  ``forward(context, lens)``: context
    (B, in_dim, T), lens a ``SequenceLength`` or a length tensor -> (B, out_dim, T)."""

    def __init__(self, in_dim: int = None, out_dim: int = None, n_layers=2, n_channels=256, kernel_size=3, p_dropout=0.1,
                 lstm_type: Optional[str] = "bilstm", use_linear=True, use_weight_norm=True):
        super().__init__()
        if not use_weight_norm:
            raise NotImplementedError("radmmm_b200.encoders.ConvLSTMLinear: use_weight_norm=False (no shipped config uses it)")
        self.out_dim = out_dim
        self.lstm_type = lstm_type
        self.use_linear = use_linear
        self.kernel_size = kernel_size
        self.dropout = nn.Dropout(p=p_dropout)
        self.convolutions = nn.ModuleList(
            [_ConvNormHolder(in_dim if i == 0 else n_channels, n_channels, kernel_size) for i in range(n_layers)])
        if not self.use_linear:
            n_channels = out_dim
        if self.lstm_type is not None:
            bi = self.lstm_type == "bilstm"
            self.bilstm = nn.LSTM(n_channels, int(n_channels // 2) if bi else n_channels, 1, batch_first=True, bidirectional=bi)
            self.bilstm = nn.utils.spectral_norm(self.bilstm, "weight_hh_l0")
            if bi:
                self.bilstm = nn.utils.spectral_norm(self.bilstm, "weight_hh_l0_reverse")
        if self.use_linear:
            self.dense = nn.Linear(n_channels, out_dim)
        self.precision = "fp32"

    def forward(self, context, lens):
        ln = _lens_tensor(lens, context.device)
        t_max = context.shape[2]
        mask = (torch.arange(t_max, device=context.device)[None, :] < ln[:, None]).to(context.dtype)[:, None]
        x = context
        for holder in self.convolutions:
            conv = holder.conv
            y = F.conv1d(x * mask, _weight(conv), conv.bias, padding=(self.kernel_size - 1) // 2)
            x = self.dropout(F.relu(y)) * mask
        if self.lstm_type != "" and self.lstm_type is not None:
            x = _packed_bilstm(self.bilstm, x.transpose(1, 2), ln, self.precision).transpose(1, 2)
        if self.use_linear:
            x = self.dense(x.transpose(1, 2)).transpose(1, 2)
        return x


class BottleneckLayer(nn.Module):
    """attribute_predictors.py:27-51: weight-normed ConvNorm(in_dim -> in_dim / reduction_factor, k) on the masked input,
    re-masked, LeakyReLU (or ReLU).  ``norm='instancenorm'`` is not covered (no shipped predictor config uses it)."""

    def __init__(self, in_dim, reduction_factor=16, norm="weightnorm", non_linearity="leakyrelu", kernel_size=3,
                 use_partial_padding=True):
        super().__init__()
        if norm != "weightnorm":
            raise NotImplementedError("radmmm_b200.encoders.BottleneckLayer: only norm='weightnorm' (the shipped configs)")
        self.reduction_factor = reduction_factor
        self.out_dim = int(in_dim / reduction_factor)
        self.kernel_size = kernel_size
        if self.reduction_factor > 1:
            self.projection_fn = _ConvNormHolder(in_dim, self.out_dim, kernel_size)
            self.non_linearity = nn.LeakyReLU() if non_linearity == "leakyrelu" else nn.ReLU()

    def forward(self, x, mask):
        if self.reduction_factor > 1:
            m = mask.unsqueeze(1).to(x.dtype)
            conv = self.projection_fn.conv
            x = F.conv1d(x, _weight(conv), conv.bias, padding=(self.kernel_size - 1) // 2) * m      # ConvNorm.forward(signal, mask)
            x = self.non_linearity(x)
        return x


class ConvLSTMLinearDAP(nn.Module):
    """attribute_predictors.py:142-197 (f0 / energy / voiced / duration predictors of configs/RADMMM_*model_config.yaml):
    bottleneck -> [speaker (accent) embedding concat] -> ConvLSTMLinear.  Same constructor, ``forward`` / ``infer``
    signatures and ``state_dict`` keys; the target transforms are attribute_predictors.py:64-126 without their
    ``assert ...item()`` host syncs."""

    def __init__(self, n_speaker_dim=16, n_accent_dim=0, in_dim=512, out_dim=1, reduction_factor=16, n_backbone_layers=2,
                 n_hidden=256, kernel_size=3, p_dropout=0.25, target_scale=1, target_offset=0, log_target=False,
                 lstm_type: Optional[str] = "bilstm", use_speaker_embedding=True, use_accent_embedding=False,
                 normalize_target=False, normalization_type=None):
        super().__init__()
        self.target_scale, self.target_offset, self.log_target = target_scale, target_offset, log_target
        self.normalize_target, self.normalization_type = normalize_target, normalization_type
        self.use_speaker_embedding = bool(use_speaker_embedding)
        self.use_accent_embedding = bool(use_accent_embedding)
        self.bottleneck_layer = BottleneckLayer(in_dim=in_dim, reduction_factor=reduction_factor)
        backbone_in = self.bottleneck_layer.out_dim + (n_speaker_dim if use_speaker_embedding else 0) + \
            (n_accent_dim if use_accent_embedding else 0)
        self.feat_pred_fn = ConvLSTMLinear(in_dim=backbone_in, out_dim=out_dim, n_layers=n_backbone_layers, n_channels=n_hidden,
                                           kernel_size=kernel_size, p_dropout=p_dropout, lstm_type=lstm_type)

    def tx_data(self, x, x_mean=None, x_std=None):
        if self.normalize_target:
            if self.normalization_type == "norm_lin_space":
                x = torch.log(x - (x_mean / x_std)[:, None] + 10) / 3          # attribute_predictors.py:71-79 (precedence as written)
            elif self.normalization_type == "norm_log_space":
                x = ((x - x_mean[:, None, None]) / x_std[:, None, None] + 5) / 10
            return x
        x = x * self.target_scale + self.target_offset
        return torch.log(x + 1) if self.log_target else x

    def inv_tx_data(self, x, x_mean=None, x_std=None):
        if self.normalize_target:
            if self.normalization_type == "norm_lin_space" and x_mean is not None and x_std is not None:
                x = (torch.exp(x * 3) - 10) * x_std + x_mean
            elif self.normalization_type == "norm_log_space" and x_mean is not None and x_std is not None:
                x = (x * 10 - 5) * x_std[:, None, None] + x_mean[:, None, None]
            return x
        if self.log_target:
            x = torch.exp(x) - 1
        return (x - self.target_offset) / self.target_scale

    def forward(self, x_target, text_enc, spk_emb, lens, x_mean=None, x_std=None, accent_emb=None):
        if x_target is not None:
            x_target = self.tx_data(x_target, x_mean, x_std)
        ln = _lens_tensor(lens, text_enc.device)
        mask = torch.arange(text_enc.shape[2], device=text_enc.device)[None, :] < ln[:, None]
        context = self.bottleneck_layer(text_enc, mask)
        if self.use_speaker_embedding:
            context = torch.cat((context, spk_emb[..., None].expand(-1, -1, text_enc.shape[2])), 1)
        if self.use_accent_embedding:
            context = torch.cat((context, accent_emb[..., None].expand(-1, -1, text_enc.shape[2])), 1)
        return {"x_hat": self.feat_pred_fn(context, ln), "x": x_target}

    def infer(self, text_enc, spk_emb, lens, x_mean=None, x_std=None, accent_emb=None):
        return self.inv_tx_data(self.forward(None, text_enc, spk_emb, lens, accent_emb=accent_emb)["x_hat"], x_mean, x_std)
