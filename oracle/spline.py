"""Oracle (tests only): spline coupling, FiLM parameter net, masked batch-norm.

Restates splines.py, common.py:706-773 (FiLMResBlock/FiLMStack), common.py:1006-1090
(SplineTransformationLayer) and maskedbatchnorm1d.py of the reference.  Unlike the reference
(boolean-index scatter, splines.py:251-259) every element is evaluated and the pass-through
for values outside [0,1) is applied with ``torch.where`` -- same results, no host syncs.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .flow import length_mask, partial_conv1d, weight_norm_weight

Tensor = torch.Tensor


# --------------------------------------------------------------------------- quadratic spline
def quadratic_spline(x: Tensor, w_tilde: Tensor, v_tilde: Tensor, inverse: bool = False):
    """unbounded_piecewise_quadratic_transform with lower=0, upper=1 (splines.py:241-339).

    x: (...,), w_tilde: (..., K), v_tilde: (..., K+1).  Returns (y, log_j); log_j is zero outside
    [0,1) and is ``None`` for the inverse, as in the reference.
    """
    eps = torch.finfo(x.dtype).eps
    inside = (x >= 0) & (x < 1)
    k = w_tilde.shape[-1]
    w = torch.softmax(w_tilde, dim=-1)
    v = torch.exp(v_tilde - v_tilde.max(dim=-1, keepdim=True)[0]) + 1e-8           # weighted_softmax :267-272
    v = v / torch.sum((v[..., :-1] + v[..., 1:]) / 2 * w, dim=-1, keepdim=True)
    w_cum = torch.cumsum(w, dim=-1)
    w_cum[..., -1] = 1.0
    w_cum_shift = F.pad(w_cum, (1, 0))
    cdf = torch.cumsum((v[..., 1:] + v[..., :-1]) / 2 * w, dim=-1)
    cdf[..., -1] = 1.0
    cdf_shift = F.pad(cdf, (1, 0))
    xs = torch.where(inside, x, torch.zeros_like(x)).unsqueeze(-1)
    idx = torch.searchsorted(cdf if inverse else w_cum, xs.contiguous()).clamp(max=k - 1)
    w_b = torch.gather(w, -1, idx).squeeze(-1)
    w_bn1 = torch.gather(w_cum_shift, -1, idx).squeeze(-1)
    v_b = torch.gather(v, -1, idx).squeeze(-1)
    v_bp1 = torch.gather(v, -1, idx + 1).squeeze(-1)
    cdf_bn1 = torch.gather(cdf_shift, -1, idx).squeeze(-1)
    xs = xs.squeeze(-1)
    if not inverse:
        alpha = (xs - w_bn1) / w_b.clamp(min=eps)
        c = (alpha ** 2) / 2 * (v_bp1 - v_b) * w_b + alpha * v_b * w_b + cdf_bn1
        log_j = torch.lerp(v_b, v_bp1, alpha).clamp(min=eps).log()
        c = c.clamp(min=eps, max=1.0 - eps)
        return torch.where(inside, c, x), torch.where(inside, log_j, torch.zeros_like(x))
    a = (v_bp1 - v_b) * w_b / 2
    b = v_b * w_b
    c = cdf_bn1 - xs
    alpha = (-b + torch.sqrt(b ** 2 - 4 * a * c)) / (2 * a)
    inv = (alpha * w_b + w_bn1).clamp(min=eps, max=1.0 - eps)
    return torch.where(inside, inv, x), None


# --------------------------------------------------------------------------- linear spline
def linear_spline(x: Tensor, q_tilde: Tensor):
    """piecewise_linear_transform, splines.py:57-142 (outlier pass-through on)."""
    b = q_tilde.shape[-1]
    w = 1.0 / b
    q = torch.softmax(q_tilde, dim=-1) / w
    mx = torch.clamp(torch.floor(b * x), 0, b - 1).long()
    slopes = torch.gather(q, -1, mx.unsqueeze(-1)).squeeze(-1)
    left = torch.roll(torch.cumsum(q, -1) * w, 1, -1)
    left[..., 0] = 0
    out = (x - mx * w) * slopes + torch.gather(left, -1, mx.unsqueeze(-1)).squeeze(-1)
    eps = torch.finfo(out.dtype).eps
    out = out.clamp(min=eps, max=1.0 - eps)
    oob = ((x < 0.0) | (x > 1.0)).to(x.dtype)
    out = out * (1 - oob) + x * oob
    slopes = slopes * (1 - oob) + oob
    return out, torch.sum(torch.log(slopes), -1)


def linear_spline_inverse(y: Tensor, q_tilde: Tensor):
    """piecewise_linear_inverse_transform, splines.py:145-238."""
    b = q_tilde.shape[-1]
    w = 1.0 / b
    q = torch.softmax(q_tilde, dim=-1) / w
    left = torch.roll(torch.cumsum(q, -1) * w, 1, -1)
    left[..., 0] = 0
    edges = y.unsqueeze(-1) - left
    edges = torch.where(edges < 0, torch.full_like(edges, 2.0), edges)
    idx = torch.clamp(torch.argmin(edges, dim=-1), 0, b - 1)
    left_g = torch.gather(left, -1, idx.unsqueeze(-1)).squeeze(-1)
    q_g = torch.gather(q, -1, idx.unsqueeze(-1)).squeeze(-1)
    x = (y - left_g) / q_g + idx * w
    eps = torch.finfo(x.dtype).eps
    x = x.clamp(min=eps, max=1.0 - eps)
    oob = ((y < 0.0) | (y > 1.0)).to(y.dtype)
    x = x * (1 - oob) + y * oob
    q_g = q_g * (1 - oob) + oob
    return x, -torch.sum(torch.log(q_g), -1)


# --------------------------------------------------------------------------- masked BN / FiLM
def masked_batchnorm(sd: Dict[str, Tensor], pre: str, x: Tensor, mask: Tensor, training: bool,
                     eps: float = 1e-5, momentum: float = 0.1, update: bool = False) -> Tensor:
    """maskedbatchnorm1d.py:53-118 (single process; the optional all-reduce at :88-95 sums the same three
    statistics over ranks).  ``mask`` is (B,1,T) float.  With ``update`` the running stats in ``sd`` are
    updated in place like the reference does."""
    n = mask.sum()
    if training and n > 1:
        mean = (mask * x).sum([0, 2]) / n
        var = (mask * x ** 2).sum([0, 2]) / n - mean ** 2
        if update:
            sd[pre + "running_mean"] = momentum * mean.detach() + (1 - momentum) * sd[pre + "running_mean"]
            sd[pre + "running_var"] = momentum * var.detach() * n / (n - 1) + (1 - momentum) * sd[pre + "running_var"]
            sd[pre + "num_batches_tracked"] = sd[pre + "num_batches_tracked"] + 1
    else:
        mean, var = sd[pre + "running_mean"], sd[pre + "running_var"]
    x = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + eps)
    return x * sd[pre + "weight"][None, :, None] + sd[pre + "bias"][None, :, None]


def _pconv(sd, pre: str, x: Tensor, lens: Tensor, dilation: int = 1) -> Tensor:
    w = weight_norm_weight(sd[pre + "conv.weight_g"], sd[pre + "conv.weight_v"])
    return partial_conv1d(x, w, sd[pre + "conv.bias"], lens, dilation)


def film_stack(sd, pre: str, x: Tensor, ctx: Tensor, lens: Tensor, n_layers: int, use_bn: bool,
               training: bool, update_bn: bool = False) -> Tensor:
    """FiLMStack.forward (common.py:764-773) over FiLMResBlock.forward (common.py:723-735)."""
    mask = length_mask(lens, x.shape[2])[:, None].to(x.dtype)
    for i in range(n_layers):
        p = f"{pre}in_layers.{i}."
        x1 = _pconv(sd, p + "input_conv.", x, lens)
        c1 = _pconv(sd, p + "cond_conv.", ctx, lens)
        n_out = x1.shape[1]
        scale, bias = c1[:, :n_out] + 1, c1[:, n_out:]
        r = F.leaky_relu(x1, 0.01)
        x2 = _pconv(sd, p + "hidden_conv.", r, lens, 2 ** i)
        if use_bn:
            x2 = masked_batchnorm(sd, p + "bn.", x2, mask, training, update=update_bn)
        x2 = F.leaky_relu(x2 * scale + bias, 0.01)
        x = 0.5 * (x2 + r)
    return F.conv1d(x, sd[pre + "end.weight"], sd[pre + "end.bias"])


def spline_coupling(sd, pre: str, z: Tensor, ctx: Tensor, lens: Tensor, cfg, inverse: bool = False,
                    training: bool = True, update_bn: bool = False):
    """SplineTransformationLayer.forward with use_quadratic=True (common.py:1040-1090) as built by
    FlowStep (decoders.py:51-61): bounds +-3, 32 bins -> 65 parameters per channel."""
    left = bottom = -3.0
    right = top = 3.0
    nb = 2 * cfg.spline_bins + 1
    b_s, c_s, t_s = z.shape
    n_half = c_s // 2
    z0, z1 = z[:, :n_half], z[:, n_half:]
    x = (z1 - bottom) / (top - bottom) if inverse else (z1 - left) / (right - left)
    q = film_stack(sd, pre + "param_predictor.", z0, ctx, lens, cfg.n_conv_layers_per_step, cfg.use_bn,
                   training, update_bn)
    q = q.permute(0, 2, 1).reshape(b_s, t_s, n_half, nb)
    y, log_j = quadratic_spline(x.permute(0, 2, 1), q[..., :nb // 2], q[..., nb // 2:], inverse=inverse)
    y = y.permute(0, 2, 1)
    if inverse:
        return torch.cat((z0, y * (right - left) + left), 1)
    z1 = y * (top - bottom) + bottom
    log_s = log_j.sum(-1).unsqueeze(1) + n_half * (math.log(top - bottom) - math.log(right - left))
    return torch.cat((z0, z1), 1), log_s
