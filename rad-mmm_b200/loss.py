"""Flow negative log-likelihood (reference: loss.py:85-110 ``compute_flow_loss`` and the flow part of
``RADMMMLoss.forward``, loss.py:518-538) on masked-sum reduction kernels (fp64 accumulation, warp shuffles), and the
alignment CTC loss (loss.py:112-140 ``AttentionCTCLoss``) as one batched kernel."""
from __future__ import annotations

import torch

from . import _native as N


class _MaskedSum(torch.autograd.Function):
    """sum over t < len_b of x (square=False) or x^2 (square=True), x: (B, C, T)."""

    @staticmethod
    def forward(ctx, x, lens, square: bool):
        lib = N.lib()
        x = x.contiguous().float()
        b, c, t = x.shape
        out = torch.zeros(1, dtype=torch.float64, device=x.device)
        N.check(lib.radmmm_masked_sum(N.fptr(x), N.ptr(lens), b, c, t, int(square), N.ptr(out), N.stream()))
        ctx.save_for_backward(x, lens)
        ctx.square = square
        return out.float().reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = N.lib()
        x, lens = ctx.saved_tensors
        b, c, t = x.shape
        dx = torch.empty_like(x)
        coef = g.reshape(1).float().contiguous()
        N.check(lib.radmmm_masked_sum_backward(N.fptr(x), N.ptr(lens), b, c, t, int(ctx.square), N.fptr(coef), 1.0,
                                               N.fptr(dx), N.stream()))
        return dx, None, None


def flow_nll(z, log_det_W_list, log_s_list, lens_g, sigma: float = 1.0, n_elements=None, n_dims=None):
    """(loss, loss_prior) of loss.py:85-110 from grouped lengths instead of a dense mask.

    ``n_elements`` is the log-det multiplier and, times ``n_dims``, the normaliser.  The reference's caller passes
    ``floor(sum(out_lens) / n_group_size)`` (RADMMMLoss.forward, loss.py:520), which differs from ``sum(lens_g)`` when
    lengths are odd; use :func:`n_elements_like_reference` for that convention.  Defaults: ``sum(lens_g)`` and ``z.size(1)``.
    """
    lens = lens_g.to(device=z.device, dtype=torch.int32).contiguous()
    if n_elements is None:
        n_elements = lens.sum()
    n = torch.as_tensor(n_elements, device=z.device).to(torch.float32)
    log_s_total = sum(_MaskedSum.apply(ls, lens, False) for ls in log_s_list)
    log_det_total = sum(log_det_W_list) * n if len(log_det_W_list) else 0.0
    prior = _MaskedSum.apply(z, lens, True) / (2 * sigma * sigma)
    denom = n * (z.size(1) if n_dims is None else n_dims)
    return (prior - log_s_total - log_det_total) / denom, prior / denom


def n_elements_like_reference(out_lens, n_group_size: int):
    """``torch.div(out_lens.sum(), n_group_size, rounding_mode='floor')`` -- RADMMMLoss.forward, loss.py:520."""
    return torch.div(out_lens.sum(), n_group_size, rounding_mode="floor")


def compute_flow_loss(z, log_det_W_list, log_s_list, n_elements, n_dims, mask, sigma=1.0):
    """Same signature and arithmetic as the reference (loss.py:85-110): ``n_elements`` and ``n_dims`` are honoured as
    passed.  ``mask`` is the (B,1,T') prefix mask; the per-utterance lengths are recovered from it.  Unlike the
    reference, ``log_det_W_list[0]`` is NOT modified in place."""
    lens = mask.reshape(mask.shape[0], -1).sum(1).to(torch.int32)
    return flow_nll(z, list(log_det_W_list), log_s_list, lens, sigma, n_elements=n_elements, n_dims=n_dims)


class RADMMMFlowLoss(torch.nn.Module):
    """Flow part of RADMMMLoss (loss.py:500-538): {'loss_mel': (loss, 1.0), 'loss_prior_mel': (prior, 0.0)}."""

    def __init__(self, sigma=1.0, n_group_size=1):
        super().__init__()
        self.sigma = sigma
        self.n_group_size = n_group_size

    def forward(self, model_output, out_lens):
        lengths = out_lens.lengths if hasattr(out_lens, "lengths") else out_lens
        lens_g = torch.div(lengths, self.n_group_size, rounding_mode="floor")
        loss, prior = flow_nll(model_output["z_mel"], model_output["log_det_W_list"], model_output["log_s_list"],
                               lens_g, self.sigma, n_elements=n_elements_like_reference(lengths, self.n_group_size))
        return {"loss_mel": (loss, 1.0), "loss_prior_mel": (prior, 0.0)}


class _AttentionCTC(torch.autograd.Function):
    """Batched forward + gradient of loss.py:112-140 in one launch (csrc/alignment.cu)."""

    @staticmethod
    def forward(ctx, attn_logprob, in_lens, out_lens, blank_logprob: float):
        lib = N.lib()
        x = attn_logprob.contiguous().float()
        b, _, t1, t2 = x.shape
        cost = torch.zeros(b, device=x.device)
        grad = torch.empty_like(x)
        with N.on_device_of(x):
            N.check(lib.radmmm_attention_ctc(N.fptr(x), N.ptr(in_lens), N.ptr(out_lens), N.fptr(cost), N.fptr(grad), b, t1, t2,
                                             float(blank_logprob), N.stream()))
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(cost)
        return cost.mean(), cost

    @staticmethod
    def backward(ctx, g, _g_each):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class AttentionCTCLoss(torch.nn.Module):
    """Drop-in for loss.py:112-140: same constructor, same ``forward(attn_logprob, in_lens, out_lens, return_all=False)``.
    The reference loops over the batch (log_softmax + nn.CTCLoss(zero_infinity=True) per utterance); here the whole batch
    is one kernel that also produces the gradient w.r.t. ``attn_logprob``."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.blank_logprob = blank_logprob

    def forward(self, attn_logprob, in_lens, out_lens, return_all=False):
        if not attn_logprob.is_cuda:
            raise RuntimeError("radmmm_b200.loss.AttentionCTCLoss: expected a CUDA tensor (the kernels have no CPU path)")
        dev = attn_logprob.device
        il = torch.as_tensor(in_lens).to(device=dev, dtype=torch.int32).contiguous()
        ol = torch.as_tensor(out_lens).to(device=dev, dtype=torch.int32).contiguous()
        cost, each = _AttentionCTC.apply(attn_logprob, il, ol, float(self.blank_logprob))
        if return_all:
            return cost, list(each.unbind(0))
        return cost
