"""Drop-in for ``decoders.RADMMMFlow`` / ``decoders.FlowStep`` of NVIDIA/RAD-MMM (decoders.py:36-248).

Same constructor signature, attribute names, ``forward`` / ``infer`` signatures, output dictionary and
``state_dict`` layout, so a Lightning config only has to change ``class_path: decoders.RADMMMFlow`` to
``class_path: radmmm_b200.decoders.RADMMMFlow``.  Each flow step runs as hand-written sm_100a kernels through the
C ABI (see common.FlowStepFunction); ``precision`` selects the contraction path:
  "bf16x3"  tcgen05, bf16 hi/lo split, fp32-grade parity with the reference (default)
  "bf16"    tcgen05, bf16 operands, fp32 accumulate (throughput mode, looser tolerance)
  "fp32"    FFMA fp32, bit-faithful checker path
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from . import common
from .common import (AffineTransformationLayer, DataInitializedInvertible1x1Conv, Invertible1x1ConvLUS,
                     SequenceLength, _flow_apply)
from .models.radmmm import RADMMM, squeeze_time, unsqueeze_time


def freeze(model):
    for p in model.parameters():
        p.requires_grad = False


class FlowStep(nn.Module):
    """decoders.py:36-80."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, affine_model="simple_conv", scaling_fn="exp",
                 mode="LUS", affine_activation="softplus", use_partial_padding=False, cache_inverse=False,
                 use_spline=False, use_bn=True):
        super().__init__()
        assert mode in {"LUS", "whiten"}
        if mode == "LUS":
            self.invtbl_conv = Invertible1x1ConvLUS(n_mel_channels, cache_inverse=cache_inverse)
        else:
            self.invtbl_conv = DataInitializedInvertible1x1Conv(n_mel_channels, cache_inverse=cache_inverse)
        self.use_spline = use_spline
        if use_spline:
            from .splines import SplineTransformationLayer
            self.coupling_tfn = SplineTransformationLayer(
                n_mel_channels, n_context_dim, n_layers, scaling_fn=scaling_fn, top=3, bottom=-3, left=-3, right=3,
                n_bins=32, use_quadratic=True, use_bn=use_bn)
        else:
            self.coupling_tfn = AffineTransformationLayer(
                n_mel_channels, n_context_dim, n_layers, affine_model=affine_model, scaling_fn=scaling_fn,
                affine_activation=affine_activation, use_partial_padding=use_partial_padding)

    def enable_inverse_cache(self):
        self.invtbl_conv.cache_inverse = True

    def forward(self, z, context, inverse=False, seq_lens=None):
        conv = self.invtbl_conv
        if self.use_spline:        # un-fused: 1x1 conv kernel, then the spline coupling kernels
            if inverse:
                z = self.coupling_tfn(z, context, inverse, seq_lens=seq_lens)
                return conv(z, inverse, lens=seq_lens)
            z, log_det_W = conv(z, lens=seq_lens)
            z, log_s = self.coupling_tfn(z, context, seq_lens=seq_lens)
            return z, log_det_W, log_s
        tfn = self.coupling_tfn
        mean = conv.input_mean.reshape(-1).contiguous() if hasattr(conv, "input_mean") else None
        if inverse:
            return _flow_apply(tfn.affine_param_predictor, None, conv._inverse_weight(), mean, z, context, seq_lens,
                               tfn.scaling_fn, tfn.precision, inverse=True)
        if hasattr(conv, "maybe_initialize"):
            conv.maybe_initialize(z, seq_lens if seq_lens is not None else
                                  torch.full((z.shape[0],), z.shape[2], device=z.device))
        pre, self._pre_W = getattr(self, "_pre_W", None), None     # assembled on the side stream by RADMMMFlow.forward
        W, log_det_W = pre if pre is not None else (conv._weight(), conv.log_det())
        z_out, log_s, _ = _flow_apply(tfn.affine_param_predictor, W, None, mean, z, context, seq_lens,
                                      tfn.scaling_fn, tfn.precision)
        return z_out, log_det_W, log_s


class RADMMMFlow(RADMMM):
    """decoders.py:82-248."""

    def __init__(self, n_speaker_dim=16, use_accent=True, n_accent_dim=1, n_text_dim=512, n_group_size=1,
                 n_mel_channels=80, use_spk_emb_for_alignment=False, n_f0_dims=1, n_energy_avg_dims=1,
                 context_w_f0_and_energy=True, use_context_lstm=True, context_lstm_norm: Optional[str] = None,
                 n_flows=8, n_conv_layers_per_step=4, n_early_size=2, n_early_every=2, affine_model: str = "wavenet",
                 scaling_fn: str = "tanh", affine_activation: str = "softplus", use_partial_padding=True,
                 n_splines=0, use_bn=True, freeze_whitening_layer=False, use_accent_emb_for_decoder=False):
        super().__init__(n_speaker_dim, use_accent, n_accent_dim, n_text_dim, n_group_size, n_mel_channels,
                         use_spk_emb_for_alignment, n_f0_dims, n_energy_avg_dims, context_w_f0_and_energy,
                         use_context_lstm, context_lstm_norm, use_accent_emb_for_decoder=use_accent_emb_for_decoder)
        assert n_speaker_dim % 2 == 0
        assert n_early_size % 2 == 0
        self.use_accent = bool(use_accent)
        if self.use_accent:
            assert n_accent_dim % 2 == 0
        self.matrix_decomposition = "LUS"
        self.use_partial_padding = use_partial_padding
        self.flows = nn.ModuleList()
        self.affine_activation = affine_activation
        self.freeze_whitening_layer = freeze_whitening_layer
        self.n_flows = n_flows
        self.n_group_size = n_group_size
        self.exit_steps = []
        self.n_early_size = n_early_size
        chans = n_mel_channels * n_group_size
        for i in range(n_flows):
            if i > 0 and i % n_early_every == 0:
                chans -= n_early_size
                self.exit_steps.append(i)
            self.flows.append(FlowStep(chans, self.decoder_cond_dims, n_conv_layers_per_step, affine_model,
                                       scaling_fn, "whiten" if i == 0 else "LUS", affine_activation=affine_activation,
                                       use_partial_padding=use_partial_padding, use_spline=i < n_splines,
                                       use_bn=use_bn))
        if freeze_whitening_layer:
            freeze(self.flows[0].invtbl_conv)

    # -- precision of the WN contractions ("bf16x3" | "bf16" | "fp32")
    def set_precision(self, precision: str):
        from . import _native
        if precision not in _native.MODES:
            raise ValueError(f"precision must be one of {sorted(_native.MODES)}")
        for fs in self.flows:
            if hasattr(fs.coupling_tfn, "precision"):
                fs.coupling_tfn.precision = precision
        # the context LSTM follows the decoder's mode: "bf16" runs the cluster-resident tensor-core recurrence (bf16 W_hh / h,
        # fp32 state), "bf16x3" / "fp32" keep the fp32 recurrence (RADMMM_B200_LSTM_FP32=1 pins the fp32-grade path everywhere)
        pin = os.environ.get("RADMMM_B200_LSTM_FP32", "0") == "1"
        self.lstm_precision = "bf16x3" if (pin and precision == "bf16") else precision
        return self

    def is_attribute_unconditional(self):
        return self.n_f0_dims == 0 and self.n_energy_avg_dims == 0

    def unfold(self, x4d):
        """Stand-in for the ``nn.Unfold`` attribute of the reference (decoders.py:119-122): (B,C,T,1) -> (B,C*g,T')."""
        return squeeze_time(x4d.squeeze(-1), self.n_group_size)

    def fold(self, mel):
        return unsqueeze_time(mel, self.n_group_size)

    def enable_inverse_cache(self):
        for fs in self.flows:
            fs.enable_inverse_cache()

    def invalidate_weight_cache(self):
        """Treat every parameter as changed: the next forward re-runs weight norm + re-layout for all flows and the
        LSTM input weights.  Training forwards do that on every call anyway (they never trust version counters: the
        reference RAdam updates ``p.data`` without bumping them); this is for INFERENCE after weights were swapped
        through ``p.data`` outside a training step (EMA copies and the like)."""
        from . import lstm as _lstm
        for fs in self.flows:
            tfn = fs.coupling_tfn
            if hasattr(tfn, "affine_param_predictor"):
                tfn.affine_param_predictor.invalidate_prepared()
        _lstm._wcache.clear()

    def _prepare_weights_async(self, device):
        """Weight preparation of all flows depends only on the parameters, so it is forked onto its own stream and
        overlaps the (latency-bound) context LSTM; returns one event per flow step (flow i waits for its own event only,
        so the flow chain starts while later flows are still being prepared)."""
        prep = common._prep_stream(device)
        if prep is None:
            return None
        main_stream = torch.cuda.current_stream(device)
        prep.wait_stream(main_stream)
        events = []
        grad_on = torch.is_grad_enabled()
        with torch.cuda.stream(prep):
            for fs in self.flows:
                tfn = fs.coupling_tfn
                if hasattr(tfn, "affine_param_predictor"):
                    wn = tfn.affine_param_predictor
                    # a forward that will be differentiated re-derives the weights every time (see WN.prepared)
                    wn.prepare(tfn.precision, training=grad_on and any(p.requires_grad for p in wn.parameters()))
                # the 1x1-conv matrix W = P (L + I) (U + diag) and log|det W| depend on the parameters only: assemble
                # them here too (a dozen tiny kernels per flow, forward and backward) instead of on the flow chain
                conv = fs.invtbl_conv
                ready = (not hasattr(conv, "maybe_initialize")) or getattr(conv, "_init_seen", False) or not self.training
                if not fs.use_spline and ready and os.environ.get("RADMMM_B200_PRE_W", "1") != "0":
                    W, log_det = conv._weight(), conv.log_det()
                    # allocated on the preparation stream, consumed by kernels on the main stream: tell the caching
                    # allocator, or the blocks can be handed out again (to this stream) while those kernels are pending
                    W.record_stream(main_stream)
                    log_det.record_stream(main_stream)
                    fs._pre_W = (W, log_det)
                ev = torch.cuda.Event()
                ev.record(prep)
                events.append(ev)
        return events

    def forward(self, mel, spk_vecs, context, out_lens, f0=None, energy_avg=None, accent_vecs=None):
        if mel.is_cuda and mel.device.index != torch.cuda.current_device():
            with torch.cuda.device(mel.device):      # the native launches go to the CURRENT device's current stream
                return self.forward(mel, spk_vecs, context, out_lens, f0, energy_avg, accent_vecs)
        lengths = out_lens.lengths if hasattr(out_lens, "lengths") else out_lens
        prep_done = self._prepare_weights_async(mel.device) if mel.is_cuda else None
        context_w_spkvec = self.preprocess_context(context, spk_vecs, lengths, f0, energy_avg, accent_vecs=accent_vecs)
        if self.n_group_size > 1:
            mel = squeeze_time(mel, self.n_group_size)
        lens_g = torch.div(lengths, self.n_group_size, rounding_mode="floor")
        seq = _Lens(lens_g)
        z_out, log_s_list, log_det_W_list = [], [], []
        for i, flow_step in enumerate(self.flows):
            if prep_done is not None:
                torch.cuda.current_stream(mel.device).wait_event(prep_done[i])
            if i in self.exit_steps:
                z_out.append(mel[:, :self.n_early_size])
                mel = mel[:, self.n_early_size:]
            mel, log_det_W, log_s = flow_step(mel, context_w_spkvec, seq_lens=seq)
            log_s_list.append(log_s)
            log_det_W_list.append(log_det_W)
        z_out.append(mel)
        return {"z_mel": torch.cat(z_out, 1), "log_det_W_list": log_det_W_list, "log_s_list": log_s_list,
                "context_w_spkvec": context_w_spkvec}

    def infer(self, spk_vec, txt_enc, sigma, dur=None, f0=None, energy_avg=None, out_lens=None, accent_vecs=None,
              residual=None, max_frames: Optional[int] = None):
        """decoders.py:207-248.  ``residual`` (B, n_mel*g, T') optionally injects the latent sample instead of
        drawing it (the reference draws from the CUDA RNG, decoders.py:221-225); it is multiplied by nothing."""
        if txt_enc.is_cuda and txt_enc.device.index != torch.cuda.current_device():
            with torch.cuda.device(txt_enc.device):
                return self.infer(spk_vec, txt_enc, sigma, dur, f0, energy_avg, out_lens, accent_vecs, residual, max_frames)
        if out_lens is None:
            out_lens = dur.sum(1).long().to(txt_enc.device)
        # ``max_frames`` (the padded output length, e.g. f0.shape[1]) avoids two device syncs; radmmm_b200.graphs.GraphedInfer
        # needs it because a captured call cannot read device values on the host
        max_n_frames = int(out_lens.max()) if max_frames is None else int(max_frames)
        txt_enc_time_expanded = self.length_regulator(txt_enc.transpose(1, 2), dur, total=max_frames).transpose(1, 2)
        context_w_spkvec = self.preprocess_context(txt_enc_time_expanded, spk_vec, out_lens, f0, energy_avg,
                                                   accent_vecs=accent_vecs)
        g = self.n_group_size
        if residual is None:
            residual = torch.randn(txt_enc.shape[0], self.n_mel_channels * g, max_n_frames // g,
                                   device=txt_enc.device, dtype=torch.float32) * sigma
        stack = self.exit_steps.copy()
        ne = self.n_early_size
        mel = residual[:, len(stack) * ne:]
        rest = residual[:, :len(stack) * ne]
        seq = _Lens(torch.div(out_lens, g, rounding_mode="floor"))
        with torch.no_grad():
            for i, flow_step in enumerate(reversed(self.flows)):
                cur = len(self.flows) - i - 1
                mel = flow_step(mel, context_w_spkvec, inverse=True, seq_lens=seq)
                if stack and cur == stack[-1]:
                    stack.pop()
                    mel = torch.cat((rest[:, len(stack) * ne:], mel), 1)
                    rest = rest[:, :len(stack) * ne]
        if g > 1:
            mel = self.fold(mel)
        return {"mel": mel}


class _Lens:
    """Length carrier passed between flow steps: like SequenceLength but without building the mask (no sync)."""

    def __init__(self, lengths):
        self.lengths = lengths.long()
        self.lens_i32 = lengths.to(torch.int32).contiguous()
