"""Decoder base class: conditioning construction (reference: models/radmmm.py:29-168).

``preprocess_context`` keeps the reference's semantics -- squeeze context / f0 / energy by ``n_group_size``,
concatenate speaker (and optionally accent) vectors, run the packed bi-LSTM -- and returns the same
``(B, decoder_cond_dims, T')`` tensor (a transposed view of the batch-first LSTM output, exactly as the reference
returns it).  The LSTM parameters stay in an ``nn.LSTM`` module (state-dict compatible) but the computation runs on the
persistent recurrence kernel of ``radmmm_b200.lstm`` -- cuDNN's fp32 LSTM is ~4000 launch-bound kernels per step.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ..common import LengthRegulator


def squeeze_time(x: torch.Tensor, g: int) -> torch.Tensor:
    """nn.Unfold((g,1), stride=g) on (B,C,T,1) as a reshape: x'[b, c*g+j, t'] = x[b, c, g*t'+j] (decoders.py:119-122)."""
    if g == 1:
        return x
    b, c, t = x.shape
    tp = t // g
    return x[:, :, :tp * g].reshape(b, c, tp, g).permute(0, 1, 3, 2).reshape(b, c * g, tp)


def unsqueeze_time(x: torch.Tensor, g: int) -> torch.Tensor:
    """Inverse of :func:`squeeze_time` (RADMMMFlow.fold, decoders.py:151-161)."""
    if g == 1:
        return x
    b, cg, tp = x.shape
    return x.reshape(b, cg // g, g, tp).permute(0, 1, 3, 2).reshape(b, cg // g, tp * g)


class RADMMM(torch.nn.Module):
    def __init__(self, n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=512, n_group_size=1,
                 n_mel_channels=80, use_spk_emb_for_alignment=False, n_f0_dims=0, n_energy_avg_dims=0,
                 context_w_f0_and_energy=True, use_context_lstm=True, context_lstm_norm: Optional[str] = None,
                 use_accent_emb_for_decoder=False):
        super().__init__()
        self.n_speaker_dim = n_speaker_dim
        self.n_accent_dim = n_accent_dim
        self.n_mel_channels = n_mel_channels
        self.n_f0_dims = n_f0_dims
        self.n_energy_avg_dims = n_energy_avg_dims
        self.length_regulator = LengthRegulator()
        self.context_w_f0_and_energy = context_w_f0_and_energy
        self.n_group_size = n_group_size
        self.use_accent_emb_for_decoder = bool(use_accent_emb_for_decoder)
        decoder_cond_dims = None
        if self.use_accent_emb_for_decoder:
            decoder_cond_dims = (n_speaker_dim + n_accent_dim +
                                 (n_text_dim + n_f0_dims + n_energy_avg_dims) * n_group_size)
        self.use_context_lstm = use_context_lstm
        if use_context_lstm:
            hidden = n_speaker_dim + n_text_dim * n_group_size
            n_in = (n_f0_dims + n_energy_avg_dims + n_text_dim) * n_group_size + n_speaker_dim
            if self.use_accent_emb_for_decoder:
                hidden += n_accent_dim
                n_in += n_accent_dim
            hidden = int(hidden / 2)
            decoder_cond_dims = hidden * 2
            self.context_lstm = nn.LSTM(input_size=n_in, hidden_size=hidden, num_layers=1, batch_first=True,
                                        bidirectional=True)
            self._plain_lstm = context_lstm_norm is None
            if context_lstm_norm is not None:
                fn = nn.utils.spectral_norm if "spectral" in context_lstm_norm else nn.utils.weight_norm
                self.context_lstm = fn(self.context_lstm, "weight_hh_l0")
                self.context_lstm = fn(self.context_lstm, "weight_hh_l0_reverse")
        if decoder_cond_dims is None:
            raise ValueError("decoder_cond_dims is undefined without a context LSTM or decoder accent embedding "
                             "(the reference fails the same way)")
        self.decoder_cond_dims = decoder_cond_dims
        self.decoder_out_dims = n_mel_channels

    def preprocess_context(self, context, spk_vecs, out_lens=None, f0=None, energy_avg=None, accent_vecs=None):
        g = self.n_group_size
        ctx = squeeze_time(context, g)
        tp = ctx.shape[2]
        parts = [ctx, spk_vecs[:, :, None].expand(-1, -1, tp)]
        if self.use_accent_emb_for_decoder:
            assert accent_vecs is not None
            parts.append(accent_vecs[:, :, None].expand(-1, -1, tp))
        if self.context_w_f0_and_energy:
            if f0 is not None:
                parts.append(squeeze_time(f0[:, None], g))
            if energy_avg is not None:
                parts.append(squeeze_time(energy_avg[:, None], g))
        x = torch.cat(parts, 1)
        if not self.use_context_lstm:
            return x
        lens_g = torch.div(out_lens, g, rounding_mode="floor").long()
        if self._plain_lstm:
            from ..lstm import context_lstm
            out = context_lstm(self.context_lstm, x.transpose(1, 2).contiguous(), lens_g,
                               getattr(self, "lstm_precision", "bf16x3"))
            return out.transpose(1, 2)
        # spectral / weight-normed recurrent weights (context_lstm_norm): the norm lives in nn.LSTM forward hooks, so
        # this configuration (not used by the shipped configs) keeps the library LSTM
        packed = nn.utils.rnn.pack_padded_sequence(x.transpose(1, 2), lens_g.cpu(), batch_first=True, enforce_sorted=False)
        self.context_lstm.flatten_parameters()
        out, _ = self.context_lstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=tp)
        return out.transpose(1, 2)

    def remove_norms(self):
        """models/radmmm.py:150-168.  Spectral/weight norm on the LSTM is removed like the reference does; the WN
        stacks keep their (g, v) parameters because the weight norm is folded into the cached weight-preparation
        kernel (it costs nothing at inference once the weights stop changing)."""
        for name, module in self.named_modules():
            for attr in ("weight_hh_l0", "weight_hh_l0_reverse"):
                try:
                    nn.utils.remove_spectral_norm(module, name=attr)
                    print(f"Removed spectral norm from {name}")
                except Exception:
                    pass
            if isinstance(module, nn.LSTM):
                for attr in ("weight_hh_l0", "weight_hh_l0_reverse"):
                    try:
                        nn.utils.remove_weight_norm(module, name=attr)
                        print(f"Removed wnorm from {name}")
                    except Exception:
                        pass
