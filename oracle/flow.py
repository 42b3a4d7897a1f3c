"""Functional torch-CPU restatement of the RAD-MMM flow decoder (oracle; tests only).

State is a flat dict ``sd`` keyed exactly like ``decoders.RADMMMFlow.state_dict()``
in the reference (``flows.<i>.invtbl_conv.upper`` ...).  All functions are dtype
agnostic (float32 or float64) so the same code gives the fp64 "truth" used for the
noise-floor numbers in DESIGN.md.

Reference files followed: decoders.py, models/radmmm.py, common.py,
partialconv1d.py, loss.py (line numbers in each docstring, relative to
/root/reference at commit adc9ad9).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
@dataclass
class DecoderConfig:
    """Mirror of ``RADMMMFlow.__init__`` arguments (decoders.py:83-96)."""
    n_speaker_dim: int = 16
    use_accent: bool = True
    n_accent_dim: int = 1
    n_text_dim: int = 512
    n_group_size: int = 1
    n_mel_channels: int = 80
    n_f0_dims: int = 1
    n_energy_avg_dims: int = 1
    context_w_f0_and_energy: bool = True
    use_context_lstm: bool = True
    n_flows: int = 8
    n_conv_layers_per_step: int = 4
    n_early_size: int = 2
    n_early_every: int = 2
    scaling_fn: str = "tanh"
    n_splines: int = 0
    use_bn: bool = True
    use_accent_emb_for_decoder: bool = False
    wn_channels: int = 1024          # AffineTransformationLayer(n_channels=1024), common.py:1097
    film_hidden: int = 512           # SplineTransformationLayer -> FiLMStack(..., 512, ...), common.py:1036
    spline_bins: int = 32            # decoders.py:56

    @staticmethod
    def radmmm() -> "DecoderConfig":
        """init_args of configs/RADMMM_model_config.yaml:16-39."""
        return DecoderConfig(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520,
                             n_group_size=2, n_flows=8)

    def channels_per_flow(self) -> List[int]:
        """decoders.py:126-131 -- channel count drops by n_early_size at exit steps."""
        c = self.n_mel_channels * self.n_group_size
        out = []
        for i in range(self.n_flows):
            if i > 0 and i % self.n_early_every == 0:
                c -= self.n_early_size
            out.append(c)
        return out

    def exit_steps(self) -> List[int]:
        return [i for i in range(self.n_flows) if i > 0 and i % self.n_early_every == 0]

    def lstm_in_dim(self) -> int:
        """models/radmmm.py:74-79."""
        n = (self.n_f0_dims + self.n_energy_avg_dims + self.n_text_dim) * self.n_group_size
        n += self.n_speaker_dim
        if self.use_accent_emb_for_decoder:
            n += self.n_accent_dim
        return n

    def lstm_hidden(self) -> int:
        """models/radmmm.py:62-72."""
        n = self.n_speaker_dim + self.n_text_dim * self.n_group_size
        if self.use_accent_emb_for_decoder:
            n += self.n_accent_dim
        return int(n / 2)

    def cond_dims(self) -> int:
        """models/radmmm.py:81 (with LSTM) / :52-57 (without)."""
        if self.use_context_lstm:
            return 2 * self.lstm_hidden()
        assert self.use_accent_emb_for_decoder, "reference leaves decoder_cond_dims undefined otherwise"
        return (self.n_speaker_dim + self.n_accent_dim +
                (self.n_text_dim + self.n_f0_dims + self.n_energy_avg_dims) * self.n_group_size)


# --------------------------------------------------------------------------- small helpers
def squeeze_time(x: Tensor, g: int) -> Tensor:
    """``nn.Unfold((g,1), stride=g)`` on (B,C,T,1): x'[b, c*g+j, t'] = x[b, c, g*t'+j].

    decoders.py:119-122,178 and models/radmmm.py:114-120.  A trailing partial group is dropped.
    """
    if g == 1:
        return x
    b, c, t = x.shape
    tp = t // g
    return x[:, :, :tp * g].reshape(b, c, tp, g).permute(0, 1, 3, 2).reshape(b, c * g, tp)


def unsqueeze_time(x: Tensor, g: int) -> Tensor:
    """Exact inverse of :func:`squeeze_time` (``RADMMMFlow.fold``, decoders.py:151-161)."""
    if g == 1:
        return x
    b, cg, tp = x.shape
    c = cg // g
    return x.reshape(b, c, g, tp).permute(0, 1, 3, 2).reshape(b, c, tp * g)


def length_mask(lens: Tensor, max_len: Optional[int] = None) -> Tensor:
    """common.py:105-116: bool (B, max_len), True where t < len."""
    if max_len is None:
        max_len = int(lens.max())
    return torch.arange(max_len, device=lens.device)[None, :] < lens[:, None]


def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """``nn.utils.weight_norm`` default dim=0: w[o] = g[o] * v[o] / ||v[o]||_2 (common.py:174,791,813)."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def valid_tap_count(lens: Tensor, t_len: int, dilation: int, ksize: int = 5) -> Tensor:
    """Closed form of ``F.conv1d(mask, ones(1,1,k), padding=d*(k-1)/2, dilation=d)``.

    partialconv1d.py:74-77.  Returns float-like integer counts (B, T).
    """
    t = torch.arange(t_len, device=lens.device)[None, :]
    u = torch.zeros(lens.shape[0], t_len, dtype=torch.long, device=lens.device)
    half = (ksize - 1) // 2
    for j in range(ksize):
        s = t + (j - half) * dilation
        u += ((s >= 0) & (s < lens[:, None]) & (s < t_len)).long()
    return u


def partial_conv1d(x: Tensor, w: Tensor, bias: Tensor, lens: Tensor, dilation: int) -> Tensor:
    """ConvNorm(use_partial_padding=True) forward with a length mask.

    partialconv1d.py:65-94 followed by the ``conv_signal * mask`` of common.py:186-190.
    x: (B, Cin, T); w: (Cout, Cin, k); mask derived from ``lens``.
    """
    b, _, t = x.shape
    k = w.shape[-1]
    m = length_mask(lens, t)[:, None, :].to(x.dtype)
    u = valid_tap_count(lens, t, dilation, k)[:, None, :].to(x.dtype)
    ratio = k / (u + 1e-6)
    u1 = u.clamp(0, 1)
    ratio = ratio * u1
    raw = F.conv1d(x * m, w, bias, padding=dilation * (k - 1) // 2, dilation=dilation)
    bv = bias.view(1, -1, 1)
    out = ((raw - bv) * ratio + bv) * u1
    return out * m


def softplus(x: Tensor) -> Tensor:
    """``torch.nn.Softplus()`` defaults: beta=1, threshold=20 (common.py:793)."""
    return F.softplus(x, beta=1.0, threshold=20.0)


# --------------------------------------------------------------------------- invertible 1x1
def lus_weight(sd: Dict[str, Tensor], pre: str) -> Tensor:
    """W = P (tril(L,-1)+I) (triu(U,1)+diag(d)); common.py:528-531."""
    upper = torch.triu(sd[pre + "upper"], 1) + torch.diag(sd[pre + "upper_diag"])
    lower = torch.tril(sd[pre + "lower"], -1) + torch.diag(sd[pre + "lower_diag"].to(upper.dtype))
    return sd[pre + "p"].to(upper.dtype) @ (lower @ upper)


def whiten_weight(sd: Dict[str, Tensor], pre: str) -> Tensor:
    """W = triu(U,1)+diag(d); common.py:598."""
    return torch.triu(sd[pre + "upper"], 1) + torch.diag(sd[pre + "upper_diag"])


def inv1x1_forward(sd, pre: str, z: Tensor, mode: str):
    """common.py:544-548 (LUS) / :612-617 (whiten).  Returns (z, log|det W|)."""
    if mode == "whiten":
        w = whiten_weight(sd, pre)
        z = z - sd[pre + "input_mean"].to(z.dtype).unsqueeze(0)
    else:
        w = lus_weight(sd, pre)
    z = torch.einsum("oc,bct->bot", w, z)
    log_det = torch.sum(torch.log(torch.abs(sd[pre + "upper_diag"])))
    return z, log_det


def inv1x1_inverse(sd, pre: str, z: Tensor, mode: str) -> Tensor:
    """common.py:532-543 (LUS) / :599-611 (whiten)."""
    if mode == "whiten":
        w_inv = torch.linalg.inv(whiten_weight(sd, pre))
        return torch.einsum("oc,bct->bot", w_inv, z) + sd[pre + "input_mean"].to(z.dtype).unsqueeze(0)
    w_inv = torch.linalg.inv(lus_weight(sd, pre))
    return torch.einsum("oc,bct->bot", w_inv, z)


def whitening_init(z: Tensor, lens: Tensor):
    """Data-dependent init of flow 0 (common.py:569-591).  Returns (input_mean (C,1), upper, upper_diag)."""
    cols = [z[b, :, :int(lens[b])] for b in range(z.shape[0])]
    data = torch.cat(cols, dim=1)
    n = data.shape[1]
    mean = data.mean(1, keepdim=True)
    cen = data - mean
    cov = (cen @ cen.t()) / n
    wm = torch.linalg.cholesky(torch.linalg.inv(cov), upper=True)
    return mean, torch.triu(wm, 1), torch.diag(wm)


# --------------------------------------------------------------------------- WN + affine coupling
def wn_forward(sd, pre: str, z0: Tensor, ctx: Tensor, lens: Tensor, n_layers: int) -> Tensor:
    """common.py:816-835.  z0 (B,C/2,T'), ctx (B,D,T') -> (B,C,T')."""
    w = weight_norm_weight(sd[pre + "start.weight_g"], sd[pre + "start.weight_v"])
    h = F.conv1d(torch.cat((z0, ctx), 1), w, sd[pre + "start.bias"])
    out = torch.zeros_like(h)
    for i in range(n_layers):
        p = f"{pre}in_layers.{i}.conv."
        w = weight_norm_weight(sd[p + "weight_g"], sd[p + "weight_v"])
        h = softplus(partial_conv1d(h, w, sd[p + "bias"], lens, 2 ** i))
        p = f"{pre}res_skip_layers.{i}."
        w = weight_norm_weight(sd[p + "weight_g"], sd[p + "weight_v"])
        out = out + softplus(F.conv1d(h, w, sd[p + "bias"]))
    return F.conv1d(out, sd[pre + "end.weight"], sd[pre + "end.bias"])


def scale_and_log(a: Tensor, scaling_fn: str):
    """AffineTransformationLayer.get_scaling_and_logs, common.py:1127-1141."""
    if scaling_fn == "tanh":
        s = torch.tanh(a) + 1 + 1e-6
        return s, torch.log(s)
    if scaling_fn == "exp":
        return torch.exp(a), a
    if scaling_fn == "sigmoid":
        s = torch.sigmoid(a + 10) + 1e-6
        return s, torch.log(s)
    if scaling_fn == "translate":
        return torch.exp(a * 0), a * 0
    raise ValueError(scaling_fn)


def affine_coupling(sd, pre: str, z: Tensor, ctx: Tensor, lens: Tensor, n_layers: int,
                    scaling_fn: str = "tanh", inverse: bool = False):
    """common.py:1163-1185."""
    n_half = z.shape[1] // 2
    z0, z1 = z[:, :n_half], z[:, n_half:]
    params = wn_forward(sd, pre + "affine_param_predictor.", z0, ctx, lens, n_layers)
    s, log_s = scale_and_log(params[:, :n_half], scaling_fn)
    b = params[:, n_half:]
    if inverse:
        return torch.cat((z0, (z1 - b) / s), 1)
    return torch.cat((z0, s * z1 + b), 1), log_s


# --------------------------------------------------------------------------- context
def build_context_lstm(sd, cfg: DecoderConfig, dtype=torch.float32) -> torch.nn.LSTM:
    lstm = torch.nn.LSTM(cfg.lstm_in_dim(), cfg.lstm_hidden(), 1, batch_first=True, bidirectional=True)
    lstm.load_state_dict({k[len("context_lstm."):]: v for k, v in sd.items() if k.startswith("context_lstm.")})
    return lstm.to(dtype)


def preprocess_context(sd, cfg: DecoderConfig, context: Tensor, spk: Tensor, lens: Tensor,
                       f0: Optional[Tensor], energy: Optional[Tensor], accent: Optional[Tensor] = None,
                       lstm: Optional[torch.nn.LSTM] = None) -> Tensor:
    """models/radmmm.py:103-148.  Returns (B, cond_dims, T')."""
    g = cfg.n_group_size
    ctx = squeeze_time(context, g)
    tp = ctx.shape[2]
    parts = [ctx, spk[:, :, None].expand(-1, -1, tp)]
    if cfg.use_accent_emb_for_decoder:
        parts.append(accent[:, :, None].expand(-1, -1, tp))
    if cfg.context_w_f0_and_energy:
        if f0 is not None:
            parts.append(squeeze_time(f0[:, None], g))
        if energy is not None:
            parts.append(squeeze_time(energy[:, None], g))
    x = torch.cat(parts, 1)
    if not cfg.use_context_lstm:
        return x
    if lstm is None:
        lstm = build_context_lstm(sd, cfg, x.dtype)
    lens_g = torch.div(lens, g, rounding_mode="floor").long().cpu()
    packed = torch.nn.utils.rnn.pack_padded_sequence(x.transpose(1, 2), lens_g, batch_first=True,
                                                     enforce_sorted=False)
    out, _ = lstm(packed)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    return out.transpose(1, 2)


# --------------------------------------------------------------------------- decoder
def _flow_mode(i: int) -> str:
    return "whiten" if i == 0 else "LUS"      # decoders.py:133-135


def decoder_forward(sd, cfg: DecoderConfig, mel: Tensor, spk: Tensor, context: Tensor, lens: Tensor,
                    f0=None, energy=None, accent=None, lstm=None, coupling=None) -> Dict[str, object]:
    """RADMMMFlow.forward, decoders.py:168-205.  ``lens`` are un-grouped frame counts."""
    from . import spline as _spline
    ctx = preprocess_context(sd, cfg, context, spk, lens, f0, energy, accent, lstm)
    z = squeeze_time(mel, cfg.n_group_size)
    lens_g = torch.div(lens, cfg.n_group_size, rounding_mode="floor")
    exits, log_s_list, log_det_list = [], [], []
    for i in range(cfg.n_flows):
        if i in cfg.exit_steps():
            exits.append(z[:, :cfg.n_early_size])
            z = z[:, cfg.n_early_size:]
        pre = f"flows.{i}."
        z, log_det = inv1x1_forward(sd, pre + "invtbl_conv.", z, _flow_mode(i))
        if i < cfg.n_splines:
            z, log_s = _spline.spline_coupling(sd, pre + "coupling_tfn.", z, ctx, lens_g, cfg)
        else:
            z, log_s = affine_coupling(sd, pre + "coupling_tfn.", z, ctx, lens_g,
                                       cfg.n_conv_layers_per_step, cfg.scaling_fn)
        log_s_list.append(log_s)
        log_det_list.append(log_det)
    exits.append(z)
    return {"z_mel": torch.cat(exits, 1), "log_det_W_list": log_det_list, "log_s_list": log_s_list,
            "context_w_spkvec": ctx}


def decoder_inverse(sd, cfg: DecoderConfig, residual: Tensor, ctx: Tensor, lens_g: Tensor) -> Tensor:
    """The flow loop of RADMMMFlow.infer (decoders.py:227-246) with an injected ``residual``
    (the reference draws it from the CUDA RNG at :221-225, which cannot be matched).
    ``ctx`` is the output of :func:`preprocess_context`; returns the un-grouped mel (B, n_mel, T)."""
    from . import spline as _spline
    stack = list(cfg.exit_steps())
    ne = cfg.n_early_size
    z = residual[:, len(stack) * ne:]
    rest = residual[:, :len(stack) * ne]
    for i in reversed(range(cfg.n_flows)):
        pre = f"flows.{i}."
        if i < cfg.n_splines:
            z = _spline.spline_coupling(sd, pre + "coupling_tfn.", z, ctx, lens_g, cfg, inverse=True)
        else:
            z = affine_coupling(sd, pre + "coupling_tfn.", z, ctx, lens_g,
                                cfg.n_conv_layers_per_step, cfg.scaling_fn, inverse=True)
        z = inv1x1_inverse(sd, pre + "invtbl_conv.", z, _flow_mode(i))
        if stack and i == stack[-1]:
            stack.pop()
            z = torch.cat((rest[:, len(stack) * ne:], z), 1)
            rest = rest[:, :len(stack) * ne]
    return unsqueeze_time(z, cfg.n_group_size)


def length_regulate(x: Tensor, dur: Tensor) -> Tensor:
    """common.LengthRegulator (common.py:208-237): x (B,T2,C), integer dur (B,T2) -> (B, max sum dur, C)."""
    outs = [torch.repeat_interleave(x[b], dur[b].long(), dim=0) for b in range(x.shape[0])]
    t = max(o.shape[0] for o in outs)
    return torch.stack([F.pad(o, (0, 0, 0, t - o.shape[0])) for o in outs])


# --------------------------------------------------------------------------- loss
def flow_loss(z: Tensor, log_det_list: Sequence[Tensor], log_s_list: Sequence[Tensor], lens_g: Tensor,
              sigma: float = 1.0, n_elements=None):
    """compute_flow_loss (loss.py:85-110).  ``n_elements`` defaults to ``sum(lens_g)``; RADMMMLoss.forward
    (loss.py:520) passes ``floor(sum(out_lens) / n_group_size)`` instead, which differs for odd lengths."""
    n = lens_g.sum() if n_elements is None else n_elements
    mask = length_mask(lens_g, z.shape[2])[:, None].to(z.dtype)
    log_s_total = sum(torch.sum(ls * mask) for ls in log_s_list)
    log_det_total = sum(log_det_list) * n
    prior = torch.sum((z * mask) ** 2) / (2 * sigma * sigma)
    denom = n * z.shape[1]
    return (prior - log_s_total - log_det_total) / denom, prior / denom


# --------------------------------------------------------------------------- optimizer (next-row 8f-1)
def clip_grad_norm(grads: Sequence[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (what Lightning's ``gradient_clip_val`` / ``gradient_clip_algorithm: norm`` calls,
    configs/RADMMM_train_config.yaml:7-8): scales ``grads`` IN PLACE, returns the total norm."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def radam_step(params: Sequence[Tensor], grads: Sequence[Tensor], exp_avg: Sequence[Tensor], exp_avg_sq: Sequence[Tensor],
               step: int, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """One ``RAdam.step`` (radam.py:63-142) on explicit state, in place; ``step`` is the count AFTER this update
    (``state['step'] += 1`` happens before it is used, radam.py:103)."""
    beta1, beta2 = betas
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma >= 5:                                                              # radam.py:116-124
        step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma *
                                   n_sma_max / (n_sma_max - 2)) / (1 - beta1 ** step)
    else:
        step_size = lr / (1 - beta1 ** step)                                    # radam.py:125-126
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)                           # radam.py:98
        m.mul_(beta1).add_(g, alpha=1 - beta1)                                  # radam.py:100
        if weight_decay != 0:
            p.add_(p, alpha=-weight_decay * lr)                                 # radam.py:129-132
        if n_sma >= 5:
            p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)                 # radam.py:135-137
        else:
            p.add_(m, alpha=-step_size)                                         # radam.py:138-139
