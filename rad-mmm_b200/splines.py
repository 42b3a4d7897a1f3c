"""Spline coupling step (reference: common.py:706-773 FiLMResBlock / FiLMStack, common.py:1006-1090
SplineTransformationLayer, splines.py:241-339 piecewise-quadratic transform, maskedbatchnorm1d.py).

The piecewise-quadratic transform (forward / inverse / backward) runs in the native spline kernels; the FiLM parameter
network's convolutions (k=1 and dilated k=5 partial convs, hidden width 512) run on the contraction kernels through
:class:`ConvRowsFunction` (row GEMM forward and dgrad, weight-grad GEMM), with the thin element-wise glue between them
(FiLM scale/shift, LeakyReLU, masked batch-norm statistics) as torch ops on the row layout.
"""
from __future__ import annotations


import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from .common import _ConvNormHolder, _lens_of, _DEFAULT_PRECISION


# --------------------------------------------------------------------------------------------- conv on rows
def _cast(mode: int, x: torch.Tensor) -> torch.Tensor:
    if mode == N.MODE_F32:
        return x.contiguous()
    lib = N.lib()
    x = x.contiguous()
    planes = 2 if mode == N.MODE_BF16X3 else 1
    buf = torch.empty(planes * x.numel(), dtype=torch.bfloat16, device=x.device)
    N.check(lib.radmmm_cast_rows(mode, N.fptr(x), x.numel(), N.ptr(buf), x.numel(), N.stream()))
    return buf


def _pad_weight(w: torch.Tensor, n_pad: int, k_pad: int) -> torch.Tensor:
    """(Cout, Cin, k) -> [k][n_pad][k_pad] fp32, zero padded."""
    cout, cin, ks = w.shape
    out = torch.zeros(ks, n_pad, k_pad, device=w.device, dtype=torch.float32)
    out[:, :cout, :cin] = w.permute(2, 0, 1)
    return out


class ConvRowsFunction(torch.autograd.Function):
    """y[r] = sum_j W_j x[r + (j - c) d]   on row matrices ([R][K_pad] fp32 in, [R][N_pad] fp32 out, zero rows between
    utterances).  Forward and input gradient are row GEMMs, the weight gradient is the weight-grad GEMM."""

    @staticmethod
    def forward(ctx, x_rows, weight, dilation: int, mode: int):
        lib = N.lib()
        r, k_pad = x_rows.shape
        cout, cin, ks = weight.shape
        n_pad = N.round_up(cout, 128)
        xa = _cast(mode, x_rows)
        wa = _cast(mode, _pad_weight(weight.detach(), n_pad, k_pad).reshape(ks * n_pad, k_pad))
        y = torch.zeros(r, n_pad, device=x_rows.device)
        N.check(lib.radmmm_conv_rows(mode, N.ptr(xa), k_pad, r * k_pad, N.ptr(wa), k_pad, ks * n_pad * k_pad, n_pad * k_pad,
                                     None, N.fptr(y), n_pad, r, k_pad, cout, ks, dilation, N.stream()))
        ctx.save_for_backward(x_rows, weight)
        ctx.dilation, ctx.mode = dilation, mode
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = N.lib()
        x_rows, weight = ctx.saved_tensors
        mode, dil = ctx.mode, ctx.dilation
        r, k_pad = x_rows.shape
        cout, cin, ks = weight.shape
        n_pad = N.round_up(cout, 128)
        kp128 = N.round_up(k_pad, 128)
        dya = _cast(mode, dy.contiguous())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dx[r] = sum_j W_j^T dy[r - (j - c) d]: taps reversed, weights transposed
            wt = torch.zeros(ks, kp128, n_pad, device=dy.device)
            wt[:, :cin, :cout] = weight.detach().permute(2, 1, 0).flip(0)
            wta = _cast(mode, wt.reshape(ks * kp128, n_pad))
            dxf = torch.zeros(r, kp128, device=dy.device)
            N.check(lib.radmmm_conv_rows(mode, N.ptr(dya), n_pad, r * n_pad, N.ptr(wta), n_pad, ks * kp128 * n_pad,
                                         kp128 * n_pad, None, N.fptr(dxf), kp128, r, n_pad, k_pad, ks, dil, N.stream()))
            dx = dxf[:, :k_pad].contiguous() if kp128 != k_pad else dxf
        if ctx.needs_input_grad[1]:
            xa = _cast(mode, x_rows)
            out = torch.empty(ks, n_pad, k_pad, device=dy.device)
            N.check(lib.radmmm_wgrad_rows(mode, N.ptr(dya), n_pad, r * n_pad, N.ptr(xa), k_pad, r * k_pad, N.fptr(out), k_pad,
                                          n_pad * k_pad, r, n_pad, k_pad, ks, dil, 0, N.stream()))
            dw = out[:, :cout, :cin].permute(1, 2, 0).contiguous()
        return dx, dw, None, None


def _effective_weight(holder) -> torch.Tensor:
    g, v = holder.weight_g, holder.weight_v
    return v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1))


class _RowGeometry:
    """Row layout bookkeeping for one batch: validity mask and partial-conv ratios as [R,1] tensors."""

    def __init__(self, lens: torch.Tensor, batch: int, tp: int):
        self.B, self.Tp = batch, tp
        self.pitch = tp + N.ROW_GAP
        self.R = N.rows(batch, tp)
        dev = lens.device
        r = torch.arange(self.R, device=dev)
        b = torch.div(r, self.pitch, rounding_mode="floor")
        self.t = r - b * self.pitch
        ln = torch.where(b < batch, lens.long().clamp(max=tp)[b.clamp(max=batch - 1)], torch.zeros_like(b))
        self.len = ln
        self.mask = (self.t < ln).float()[:, None]
        self._ratio = {}

    def ratio(self, ksize: int, dilation: int) -> torch.Tensor:
        key = (ksize, dilation)
        if key not in self._ratio:
            u = torch.zeros_like(self.t)
            for j in range(ksize):
                s = self.t + (j - ksize // 2) * dilation
                u = u + ((s >= 0) & (s < self.len)).long()
            self._ratio[key] = (ksize / (u.float() + 1e-6))[:, None] * self.mask
        return self._ratio[key]

    def rows_from_cf(self, x: torch.Tensor, k_pad: int) -> torch.Tensor:
        """(B, C, Tp) -> [R][k_pad] fp32 (all frames of the padded batch, gap rows zero)."""
        b, c, tp = x.shape
        y = F.pad(x.permute(0, 2, 1), (0, k_pad - c, 0, self.pitch - tp)).reshape(b * self.pitch, k_pad)
        return F.pad(y, (0, 0, 0, self.R - b * self.pitch))

    def cf_from_rows(self, y: torch.Tensor, c: int) -> torch.Tensor:
        return y[:self.B * self.pitch].reshape(self.B, self.pitch, -1)[:, :self.Tp, :c].permute(0, 2, 1).contiguous()


def _pconv(holder: _ConvNormHolder, x_rows: torch.Tensor, geo: _RowGeometry, dilation: int, mode: int) -> torch.Tensor:
    """ConvNorm(use_partial_padding=True) on rows: ((W (x m)) ratio + b) m   (partialconv1d.py:65-94, common.py:186-190)."""
    conv = holder.conv
    w = _effective_weight(conv)
    ks = w.shape[-1]
    y = ConvRowsFunction.apply(x_rows * geo.mask, w, dilation, mode)[:, :w.shape[0]]
    return (y * geo.ratio(ks, dilation) + conv.bias[None, :]) * geo.mask


class MaskedBatchNorm1d(nn.BatchNorm1d):
    """maskedbatchnorm1d.py:30-118 on the row layout: statistics over valid frames only, optional cross-rank sync."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.distributed_sync = False

    def forward_rows(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        n = mask.sum()
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.momentum is None and self.num_batches_tracked is not None:
            factor = 1.0 / float(self.num_batches_tracked)
        if self.training:
            sum_x = (mask * x).sum(0)
            sum_xsq = (mask * x * x).sum(0)
            if self.distributed_sync and dist.is_available() and dist.is_initialized():
                import torch.distributed.nn as distnn
                packed = torch.stack([sum_x, sum_xsq, n + torch.zeros_like(sum_x)])
                packed = distnn.all_reduce(packed, op=dist.ReduceOp.SUM)
                sum_x, sum_xsq, n = packed[0], packed[1], packed[2, 0]
            mean = sum_x / n
            var = sum_xsq / n - mean ** 2
            with torch.no_grad():
                self.running_mean = factor * mean + (1 - factor) * self.running_mean
                self.running_var = factor * var * n / (n - 1) + (1 - factor) * self.running_var
        else:
            mean, var = self.running_mean, self.running_var
        x = (x - mean[None, :]) / torch.sqrt(var[None, :] + self.eps)
        if self.affine:
            x = x * self.weight[None, :] + self.bias[None, :]
        return x


class FiLMResBlock(nn.Module):
    """common.py:706-735."""

    def __init__(self, in_channels, cond_channels, out_channels, kernel_size=1, stride=1, dilation=1, use_bn=True,
                 use_partial_padding=True):
        super().__init__()
        self.out_channels = out_channels
        self.dilation = dilation
        self.input_conv = _ConvNormHolder(in_channels, out_channels, 1)
        self.cond_conv = _ConvNormHolder(cond_channels, 2 * out_channels, 1)
        self.hidden_conv = _ConvNormHolder(out_channels, out_channels, kernel_size)
        self.use_bn = use_bn
        self.bn = MaskedBatchNorm1d(out_channels) if use_bn else None

    def forward_rows(self, x, cond, geo: _RowGeometry, mode: int):
        x1 = _pconv(self.input_conv, x, geo, 1, mode)
        c1 = _pconv(self.cond_conv, cond, geo, 1, mode)
        scale, bias = c1[:, :self.out_channels] + 1, c1[:, self.out_channels:]
        r = F.leaky_relu(x1, 0.01)
        x2 = _pconv(self.hidden_conv, F.pad(r, (0, N.round_up(r.shape[1], 64) - r.shape[1])), geo, self.dilation, mode)
        if self.use_bn:
            x2 = self.bn.forward_rows(x2, geo.mask)
        x2 = F.leaky_relu(x2 * scale + bias, 0.01)
        return 0.5 * (x2 + r)


class FiLMStack(nn.Module):
    """common.py:737-773."""

    def __init__(self, n_in_channels, n_context_dim, n_hidden_channels, n_out_channels, n_layers, kernel_size=5,
                 use_partial_padding=True, use_dilation=True, use_bn=True):
        super().__init__()
        assert kernel_size % 2 == 1
        self.n_layers = n_layers
        end = nn.Conv1d(n_hidden_channels, n_out_channels, 1)
        end.weight.data.zero_()
        end.bias.data.zero_()
        self.end = end
        self.in_layers = nn.ModuleList()
        for i in range(n_layers):
            self.in_layers.append(FiLMResBlock(n_in_channels if i == 0 else n_hidden_channels, n_context_dim,
                                               n_hidden_channels, kernel_size=kernel_size,
                                               dilation=2 ** i if use_dilation else 1, use_bn=use_bn))

    def forward_rows(self, x, cond, geo: _RowGeometry, mode: int):
        for layer in self.in_layers:
            x = layer.forward_rows(F.pad(x, (0, N.round_up(x.shape[1], 64) - x.shape[1])), cond, geo, mode)
        x = F.pad(x, (0, N.round_up(x.shape[1], 64) - x.shape[1]))
        y = ConvRowsFunction.apply(x, self.end.weight, 1, mode)[:, :self.end.weight.shape[0]]
        return y + self.end.bias[None, :]


# --------------------------------------------------------------------------------------------- spline transform
class _QuadraticSpline(torch.autograd.Function):
    """z1 (B, Ch, T), q (B, Ch*65, T) -> (z1', log_s (B,1,T));  splines.py:241-339 with bounds [lo, hi]."""

    @staticmethod
    def forward(ctx, z1, q, lens, lo: float, hi: float, n_bins: int):
        lib = N.lib()
        z1, q = z1.contiguous().float(), q.contiguous().float()
        b, ch, t = z1.shape
        out = torch.empty_like(z1)
        log_s = torch.empty(b, 1, t, device=z1.device)
        N.check(lib.radmmm_spline_forward(N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(out), N.fptr(log_s), b, ch, t, n_bins,
                                          lo, hi, 0, N.stream()))
        ctx.save_for_backward(z1, q, lens)
        ctx.cfg = (lo, hi, n_bins)
        return out, log_s

    @staticmethod
    def backward(ctx, dz_out, dlog_s):
        lib = N.lib()
        z1, q, lens = ctx.saved_tensors
        lo, hi, n_bins = ctx.cfg
        b, ch, t = z1.shape
        dz_out = dz_out.contiguous() if dz_out is not None else torch.zeros_like(z1)
        dls = dlog_s.contiguous() if dlog_s is not None else None
        dz, dq = torch.empty_like(z1), torch.empty_like(q)
        N.check(lib.radmmm_spline_backward(N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(dz_out), N.fptr(dls), N.fptr(dz),
                                           N.fptr(dq), b, ch, t, n_bins, lo, hi, N.stream()))
        return dz, dq, None, None, None, None


class _LinearSpline(torch.autograd.Function):
    """z1 (B, Ch, T), q (B, Ch*n_bins, T) -> (z1', log_s (B,1,T));  splines.py:57-142 with bounds [lo, hi]."""

    @staticmethod
    def forward(ctx, z1, q, lens, lo: float, hi: float, n_bins: int):
        lib = N.lib()
        z1, q = z1.contiguous().float(), q.contiguous().float()
        b, ch, t = z1.shape
        out = torch.empty_like(z1)
        log_s = torch.empty(b, 1, t, device=z1.device)
        N.check(lib.radmmm_spline_linear_forward(N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(out), N.fptr(log_s), b, ch, t, n_bins,
                                                 lo, hi, 0, N.stream()))
        ctx.save_for_backward(z1, q, lens)
        ctx.cfg = (lo, hi, n_bins)
        return out, log_s

    @staticmethod
    def backward(ctx, dz_out, dlog_s):
        lib = N.lib()
        z1, q, lens = ctx.saved_tensors
        lo, hi, n_bins = ctx.cfg
        b, ch, t = z1.shape
        dz_out = dz_out.contiguous() if dz_out is not None else torch.zeros_like(z1)
        dls = dlog_s.contiguous() if dlog_s is not None else None
        dz, dq = torch.empty_like(z1), torch.empty_like(q)
        N.check(lib.radmmm_spline_linear_backward(N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(dz_out), N.fptr(dls), N.fptr(dz),
                                                  N.fptr(dq), b, ch, t, n_bins, lo, hi, N.stream()))
        return dz, dq, None, None, None, None


class SplineTransformationLayer(nn.Module):
    """common.py:1006-1090: piecewise-quadratic (``use_quadratic=True``, what FlowStep builds, decoders.py:51-61) or
    piecewise-linear (``use_quadratic=False``, the class default) spline coupling on a FiLM parameter network."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, with_dilation=True, kernel_size=5, scaling_fn="exp",
                 affine_activation="softplus", n_bins=8, left=-4, right=4, bottom=-4, top=4, use_quadratic=False,
                 use_bn=True):
        super().__init__()
        if not use_quadratic and n_bins not in (8, 16, 32):
            raise NotImplementedError("piecewise-linear spline: n_bins must be 8, 16 or 32")
        if use_quadratic and n_bins != 32:
            raise NotImplementedError("piecewise-quadratic spline: n_bins must be 32 (decoders.py:56)")
        if (left, right) != (bottom, top):
            raise NotImplementedError("equal input / output bounds only (as in decoders.py:52-55)")
        self.n_mel_channels = n_mel_channels
        self.half_mel_channels = int(n_mel_channels / 2)
        self.left, self.right, self.bottom, self.top = left, right, bottom, top
        self.use_quadratic = use_quadratic
        self.n_bins = 2 * n_bins + 1 if use_quadratic else n_bins
        self.param_predictor = FiLMStack(self.half_mel_channels, n_context_dim, 512, self.half_mel_channels * self.n_bins,
                                         n_layers, use_dilation=with_dilation, kernel_size=kernel_size, use_bn=use_bn)
        self.precision = _DEFAULT_PRECISION

    def forward(self, z, context, inverse=False, seq_lens=None):
        if not z.is_cuda:
            raise RuntimeError("radmmm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        mode = N.MODES[self.precision]
        b, c, t = z.shape
        n_half = self.half_mel_channels
        lens = _lens_of(seq_lens, b, t, z.device)
        geo = _RowGeometry(lens, b, t)
        z0, z1 = z[:, :n_half], z[:, n_half:]
        x_rows = geo.rows_from_cf(z0.float(), N.round_up(n_half, 64))
        ctx_rows = geo.rows_from_cf(context.float(), N.round_up(context.shape[1], 64))
        q_rows = self.param_predictor.forward_rows(x_rows, ctx_rows, geo, mode)
        q = geo.cf_from_rows(q_rows, n_half * self.n_bins)
        lo, hi = float(self.left), float(self.right)
        if inverse:
            lib = N.lib()
            with torch.no_grad():
                z1c, qc = z1.contiguous().float(), q.contiguous()
                out = torch.empty_like(z1c)
                if self.use_quadratic:
                    N.check(lib.radmmm_spline_forward(N.fptr(z1c), N.fptr(qc), N.ptr(lens), N.fptr(out), None, b, n_half, t,
                                                      (self.n_bins - 1) // 2, lo, hi, 1, N.stream()))
                else:
                    N.check(lib.radmmm_spline_linear_forward(N.fptr(z1c), N.fptr(qc), N.ptr(lens), N.fptr(out), None, b, n_half, t,
                                                             self.n_bins, lo, hi, 1, N.stream()))
            return torch.cat((z0, out), dim=1)
        if self.use_quadratic:
            z1o, log_s = _QuadraticSpline.apply(z1, q, lens, lo, hi, (self.n_bins - 1) // 2)
        else:
            z1o, log_s = _LinearSpline.apply(z1, q, lens, lo, hi, self.n_bins)
        log_s = log_s + n_half * (np.log(self.top - self.bottom) - np.log(self.right - self.left))
        return torch.cat((z0, z1o), dim=1), log_s
