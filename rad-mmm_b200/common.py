"""Flow-op modules with the reference's names, constructor arguments, parameter names and call signatures
(reference: common.py of NVIDIA/RAD-MMM), backed by the sm_100a kernels in csrc/ through the C ABI.

Mirrored classes (reference file:line):
  SequenceLength                      common.py:123-128
  Invertible1x1ConvLUS                common.py:507-548
  DataInitializedInvertible1x1Conv    common.py:551-617
  Invertible1x1Conv                   common.py:621-662
  WN                                  common.py:776-835
  AffineTransformationLayer           common.py:1093-1185
  LengthRegulator                     common.py:208-237
The fused flow step (1x1 conv + WN + coupling in one autograd node) lives in :class:`FlowStepFunction`; the
standalone modules route through the same native entry points, so there is exactly one implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native as N

_DEFAULT_PRECISION = os.environ.get("RADMMM_B200_PRECISION", "bf16x3")


# --------------------------------------------------------------------------------------------- lengths / masks
def get_mask_from_lengths(lengths: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    """common.py:105-116.  ``max_len`` (from a tensor shape) avoids the reference's ``.item()`` device sync."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device)
    return ids < lengths.unsqueeze(1)


class SequenceLength:
    """common.py:123-128."""

    def __init__(self, lengths: torch.Tensor, max_len: Optional[int] = None):
        self.lengths = lengths.long()
        self.mask = get_mask_from_lengths(lengths, max_len)


def _lens_of(seq_lens, batch: int, tp: int, device) -> torch.Tensor:
    """int32 device lengths from a SequenceLength (ours or the reference's), a tensor, or None (= full length)."""
    if seq_lens is None:
        return torch.full((batch,), tp, dtype=torch.int32, device=device)
    cached = getattr(seq_lens, "lens_i32", None)          # decoders._Lens: converted once per step, so every flow step
    if cached is not None and cached.device == device:    # sees the SAME tensor (context-rows cache key, no re-casts)
        return cached
    lengths = seq_lens.lengths if hasattr(seq_lens, "lengths") else seq_lens
    return lengths.to(device=device, dtype=torch.int32).contiguous()


class LengthRegulator(nn.Module):
    """common.py:208-237 without the per-token Python loops: x (B,T2,C), dur (B,T2) -> (B, max sum dur, C)."""

    def forward(self, x: torch.Tensor, dur: torch.Tensor, total: Optional[int] = None) -> torch.Tensor:
        """``total`` (the padded output length) avoids the device sync when the caller already knows it."""
        dur = dur.long().clamp(min=0)
        ends = torch.cumsum(dur, dim=1)                               # (B,T2)
        if total is None:
            total = int(ends[:, -1].max().item())
        frames = torch.arange(total, device=x.device)[None, :]        # (1,T)
        idx = torch.searchsorted(ends, frames.expand(x.shape[0], -1).contiguous(), right=True)
        valid = frames < ends[:, -1:]
        idx = idx.clamp(max=x.shape[1] - 1)
        out = torch.gather(x, 1, idx[..., None].expand(-1, -1, x.shape[2]))
        return out * valid[..., None].to(x.dtype)


# --------------------------------------------------------------------------------------------- context rows cache
class _ContextRows:
    """Conditioning (B, Tp, D) fp32 re-laid-out once per step into the row format the WN kernels read."""

    def __init__(self, ctx_btd: torch.Tensor, lens: torch.Tensor, mode: int):
        lib = N.lib()
        b, tp, d = ctx_btd.shape
        self.key = (ctx_btd.data_ptr(), ctx_btd._version, tuple(ctx_btd.shape), mode, lens.data_ptr(), lens._version)
        # the key is only meaningful while the source tensors are alive (a freed block can come back at the same address
        # with version 0): the cache entry keeps them
        self.src = (ctx_btd.detach(), lens)        # detached: keeps the storage, not the autograd graph
        self.rows = torch.empty(lib.radmmm_context_rows_bytes(mode, b, tp, d), dtype=torch.uint8, device=ctx_btd.device)
        N.check(lib.radmmm_context_rows(mode, N.fptr(ctx_btd), N.ptr(lens), b, tp, d, N.ptr(self.rows), N.stream()))


_ctx_cache: List[_ContextRows] = []


def _context_rows(ctx_btd: torch.Tensor, lens: torch.Tensor, mode: int) -> _ContextRows:
    key = (ctx_btd.data_ptr(), ctx_btd._version, tuple(ctx_btd.shape), mode, lens.data_ptr(), lens._version)
    for c in _ctx_cache:
        if c.key == key:
            return c
    c = _ContextRows(ctx_btd, lens, mode)
    _ctx_cache.append(c)
    del _ctx_cache[:-2]
    return c


def _as_btd(context: torch.Tensor) -> torch.Tensor:
    """(B, D, Tp) conditioning -> contiguous (B, Tp, D) fp32.  The decoder's context is the transpose of the
    batch-first LSTM output (models/radmmm.py:146), so this is a zero-copy view in the hot path."""
    x = context.transpose(1, 2)
    return x if x.is_contiguous() else x.contiguous()


_scratch_cache = {}
_side_streams = {}


def _side_stream(device) -> int:
    """One extra stream per device for the weight-gradient lane of radmmm_flow_backward (RADMMM_B200_SIDE_STREAM=0
    disables the fork)."""
    if os.environ.get("RADMMM_B200_SIDE_STREAM", "1") == "0":
        return 0
    st = _side_streams.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _side_streams[device.index] = st
    return st.cuda_stream

_prep_streams = {}


def _prep_stream(device):
    """Stream for the per-step weight preparation (weight norm + re-layout + 1x1-conv matrix assembly).  Separate from the
    side stream so that flow i's res-skip GEMMs never queue behind the preparation of flows i+1 ..; None when the
    side-stream fork is disabled."""
    if os.environ.get("RADMMM_B200_SIDE_STREAM", "1") == "0":
        return None
    st = _prep_streams.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _prep_streams[device.index] = st
    return st


# Optional gradient sink (radmmm_b200.ddp.BucketedGradReducer): maps a parameter's storage pointer to a fresh view of
# its all-reduce bucket so FlowStepFunction.backward writes parameter gradients straight into bucket storage.
_grad_sinks: list = []


def set_grad_sink(fn):
    """Register ``fn(param_tensor) -> Tensor | None`` (tried in registration order; the first non-None answer wins), or
    clear every sink with ``None``."""
    if fn is None:
        _grad_sinks.clear()
    elif fn not in _grad_sinks:
        _grad_sinks.append(fn)


def _grad_sink(p):
    for fn in _grad_sinks:
        buf = fn(p)
        if buf is not None:
            return buf
    return None


_scratch_retired: list = []


def _backward_scratch(nbytes: int, device) -> torch.Tensor:
    """Backward scratch shared by all flow steps of a device (they run one after the other).  A buffer that has to grow
    is RETIRED, not freed: CUDA graphs captured at a smaller shape keep replaying into the old one
    (radmmm_b200.graphs.GraphedTrainStepPool holds several shapes at once)."""
    key = (device.index,)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _scratch_retired.append(buf)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _scratch_cache[key] = buf
    return buf


# --------------------------------------------------------------------------------------------- weight-normed convs
class _NormedConv1d(nn.Module):
    """Parameter holder with the key layout ``nn.utils.weight_norm(nn.Conv1d(...))`` produces:
    ``bias``, ``weight_g`` (out,1,1), ``weight_v`` (out,in,k).  ``init`` follows the reference call site."""

    def __init__(self, cin: int, cout: int, ksize: int, xavier: bool):
        super().__init__()
        conv = nn.Conv1d(cin, cout, ksize)
        if xavier:                                      # ConvNorm, common.py:172-173
            nn.init.xavier_uniform_(conv.weight, gain=nn.init.calculate_gain("linear"))
        v = conv.weight.detach()
        self.bias = nn.Parameter(conv.bias.detach().clone())
        self.weight_g = nn.Parameter(v.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1).clone())
        self.weight_v = nn.Parameter(v.clone())

    def gv(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.weight_g, self.weight_v


class _ConvNormHolder(nn.Module):
    """``ConvNorm`` keeps its conv under ``.conv`` (common.py:168-174) -> keys ``in_layers.i.conv.*``."""

    def __init__(self, cin, cout, ksize):
        super().__init__()
        self.conv = _NormedConv1d(cin, cout, ksize, xavier=True)


class WN(nn.Module):
    """common.py:776-835 (softplus activation, partial padding, dilation 2^i)."""

    def __init__(self, n_in_channels, n_context_dim, n_layers, n_channels, kernel_size=5,
                 affine_activation="softplus", use_partial_padding=True, use_dilation=True):
        super().__init__()
        assert kernel_size % 2 == 1
        assert n_channels % 2 == 0
        if kernel_size != 5 or affine_activation != "softplus" or not use_partial_padding or not use_dilation:
            raise NotImplementedError("radmmm_b200.WN builds the configuration RADMMMFlow uses: kernel_size=5, softplus, "
                                      "partial padding, dilation 2^i")
        self.n_layers = n_layers
        self.n_channels = n_channels
        self.n_in_channels = n_in_channels
        self.n_context_dim = n_context_dim
        self.affine_activation = affine_activation
        self.use_partial_padding = use_partial_padding
        self.use_dilation = use_dilation
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        self.start = _NormedConv1d(n_in_channels + n_context_dim, n_channels, 1, xavier=False)
        end = nn.Conv1d(n_channels, 2 * n_in_channels, 1)
        end.weight.data.zero_()                       # common.py:799-802: couplings start as the identity
        end.bias.data.zero_()
        self.end = end
        for _ in range(n_layers):
            self.in_layers.append(_ConvNormHolder(n_channels, n_channels, kernel_size))
            self.res_skip_layers.append(_NormedConv1d(n_channels, n_channels, 1, xavier=False))
        self._prepared = {}        # mode -> (version key | None = stale, uint8 tensor)
        self._fresh_mode = None    # mode whose weights were prepared for the forward in flight (see ``prepare``)

    # -- raw parameter list in the fixed order FlowStepFunction uses
    def raw_params(self) -> List[Optional[torch.Tensor]]:
        out: List[Optional[torch.Tensor]] = []
        g, v = self.start.gv()
        out += [g, v, self.start.bias]
        for i in range(self.n_layers):
            g, v = self.in_layers[i].conv.gv()
            out += [g, v, self.in_layers[i].conv.bias]
        for i in range(self.n_layers):
            g, v = self.res_skip_layers[i].gv()
            out += [g, v, self.res_skip_layers[i].bias]
        out += [self.end.weight, self.end.bias]
        return out

    def fill_desc(self, d: N.FlowDesc, params: List[Optional[torch.Tensor]]):
        L = self.n_layers
        d.start_g, d.start_v, d.start_b = (N.fptr(p) for p in params[0:3])
        for i in range(L):
            d.in_g[i], d.in_v[i], d.in_b[i] = (N.fptr(p) for p in params[3 + 3 * i: 6 + 3 * i])
            o = 3 + 3 * L + 3 * i
            d.rs_g[i], d.rs_v[i], d.rs_b[i] = (N.fptr(p) for p in params[o: o + 3])
        d.end_w, d.end_b = N.fptr(params[3 + 6 * L]), N.fptr(params[4 + 6 * L])

    def prepared(self, d: N.FlowDesc, params: List[Optional[torch.Tensor]], training: bool = False,
                 consume_fresh: bool = True) -> torch.Tensor:
        """Weight-normed, re-laid-out weights for ``d.mode``.

        ``training`` (a forward that will be differentiated): the preparation runs on EVERY call -- optimizers that
        update ``p.data`` in place (the reference's RAdam, radam.py:63-142) do not bump the tensors' version counters, so
        a version-keyed cache would keep serving the step-0 weights.  The one exception is the preparation RADMMMFlow
        enqueued for this very forward on its preparation stream (``prepare`` leaves a one-shot token).
        Otherwise (inference, no grad): cached on (data_ptr, version) of every raw parameter; a training-mode
        preparation leaves the cache marked stale, so the first inference call after training steps re-derives them.
        """
        lib = N.lib()
        hit = self._prepared.get(d.mode)
        if hit is None or hit[1].device != params[1].device:
            nbytes = lib.radmmm_flow_prepared_bytes(d.mode, d.C, d.D, d.H, d.L)
            hit = (None, torch.zeros(nbytes, dtype=torch.uint8, device=params[1].device))
        d.prepared = N.ptr(hit[1])
        if training:
            if consume_fresh and self._fresh_mode == d.mode:
                self._fresh_mode = None
            else:
                N.check(lib.radmmm_flow_prepare(C.byref(d), N.stream()))
            self._prepared[d.mode] = (None, hit[1])
            return hit[1]
        key = tuple((p.data_ptr(), p._version) if p is not None else None for p in params)
        if hit[0] != key:
            N.check(lib.radmmm_flow_prepare(C.byref(d), N.stream()))
        self._prepared[d.mode] = (key, hit[1])
        return hit[1]

    def invalidate_prepared(self) -> None:
        """Forget the prepared weights (e.g. after parameters were swapped through ``p.data`` outside a training step)."""
        self._prepared = {m: (None, buf) for m, (_, buf) in self._prepared.items()}
        self._fresh_mode = None

    def prepare(self, precision: str, training: bool = False) -> None:
        """Weight-norm + re-layout for ``precision`` (enqueued on the current stream).  RADMMMFlow.forward runs this for
        every flow on its preparation stream while the context LSTM runs; in training it always recomputes and leaves
        a one-shot token that the flow step of the same forward consumes."""
        d = N.FlowDesc()
        d.mode, d.B, d.C, d.Tp = N.MODES[precision], 1, 2 * self.n_in_channels, 1
        d.D, d.H, d.L = self.n_context_dim, self.n_channels, self.n_layers
        params = self.raw_params()
        with torch.no_grad():
            self.fill_desc(d, params)
            self.prepared(d, params, training=training, consume_fresh=False)
        if training:
            self._fresh_mode = d.mode

    def forward(self, forward_input, seq_lens=None):
        """(z0 (B,Cin,T), context (B,D,T)) -> (B, 2*Cin, T).  Inference-only when called on its own; training goes
        through AffineTransformationLayer / FlowStep (one fused autograd node)."""
        z0, context = forward_input
        if torch.is_grad_enabled() and (z0.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("radmmm_b200.WN.forward is not differentiable on its own; call it under torch.no_grad() "
                               "or use AffineTransformationLayer / FlowStep")
        zeros = torch.zeros_like(z0)
        z = torch.cat((z0, zeros), 1)
        _, _, params = _flow_apply(self, None, None, None, z, context, seq_lens, "translate", _DEFAULT_PRECISION,
                                   want_params=True)
        return params


# --------------------------------------------------------------------------------------------- fused flow step
def _make_desc(wn: WN, mode: int, batch: int, chans: int, tp: int, scaling_fn: str, training: bool,
               lens: torch.Tensor) -> N.FlowDesc:
    d = N.FlowDesc()
    d.mode, d.B, d.C, d.Tp = mode, batch, chans, tp
    d.D, d.H, d.L = wn.n_context_dim, wn.n_channels, wn.n_layers
    d.scaling_fn = N.SCALING[scaling_fn]
    d.training = int(training)
    d.lens = N.ptr(lens)
    return d


class FlowStepFunction(torch.autograd.Function):
    """One flow step as a single autograd node: z -> W(z - mean) -> affine coupling parameterised by WN.

    forward  -> radmmm_flow_forward   (decoders.py:72-80 forward branch)
    backward -> radmmm_flow_backward  (hand-written dgrad / wgrad chain; gradients w.r.t. z, the conditioning,
                the 1x1 matrix and every WN parameter in the reference's (g, v, bias) parametrisation)
    """

    @staticmethod
    def forward(ctx, wn: WN, scaling_fn: str, mode: int, z, ctx_btd, lens, W, mean, *params):
        lib = N.lib()
        z = z.contiguous()
        batch, chans, tp = z.shape
        training = any(ctx.needs_input_grad)
        d = _make_desc(wn, mode, batch, chans, tp, scaling_fn, training, lens)
        plist = list(params)
        wn.fill_desc(d, plist)
        prepared = wn.prepared(d, plist, training=training)
        rows = _context_rows(ctx_btd, lens, mode)
        d.ctx_rows = N.ptr(rows.rows)
        ws = torch.empty(lib.radmmm_flow_workspace_bytes(mode, int(training), batch, tp, chans, d.D, d.H, d.L),
                         dtype=torch.uint8, device=z.device)
        d.workspace = N.ptr(ws)
        d.side_stream = _side_stream(z.device) or None
        Wc = W.contiguous() if W is not None else None
        d.W, d.mean = N.fptr(Wc), N.fptr(mean)
        z_mid = torch.empty_like(z) if W is not None else z
        p_out = torch.empty_like(z)
        z_out = torch.empty_like(z)
        log_s = torch.empty(batch, chans // 2, tp, dtype=z.dtype, device=z.device)
        N.check(lib.radmmm_flow_forward(C.byref(d), N.fptr(z), N.fptr(z_mid), N.fptr(p_out), N.fptr(z_out),
                                        N.fptr(log_s), N.stream()))
        ctx.wn, ctx.scaling_fn, ctx.mode, ctx.rows, ctx.ws, ctx.prepared_buf = wn, scaling_fn, mode, rows, ws, prepared
        ctx.has_W, ctx.has_mean = W is not None, mean is not None
        ctx.ctx_shape = tuple(ctx_btd.shape)
        ctx.save_for_backward(z, z_mid, p_out, lens, Wc if Wc is not None else z.new_empty(0),
                              mean if mean is not None else z.new_empty(0), *[p for p in plist])
        ctx.mark_non_differentiable(p_out)
        return z_out, log_s, p_out

    @staticmethod
    def backward(ctx, dz_out, dlog_s, _dparams_unused):
        lib = N.lib()
        z, z_mid, p_out, lens, W, mean, *plist = ctx.saved_tensors
        wn, mode = ctx.wn, ctx.mode
        batch, chans, tp = z.shape
        d = _make_desc(wn, mode, batch, chans, tp, ctx.scaling_fn, True, lens)
        wn.fill_desc(d, plist)
        d.prepared = N.ptr(ctx.prepared_buf)
        d.ctx_rows = N.ptr(ctx.rows.rows)
        d.workspace = N.ptr(ctx.ws)
        d.side_stream = _side_stream(z.device) or None
        W_T = W.t().contiguous() if ctx.has_W else None
        d.W, d.W_T, d.mean = (N.fptr(W) if ctx.has_W else None), N.fptr(W_T), (N.fptr(mean) if ctx.has_mean else None)
        dz_out = dz_out.contiguous() if dz_out is not None else torch.zeros_like(z)
        dlog_s = dlog_s.contiguous() if dlog_s is not None else None
        grads = []
        for p in plist:
            buf = _grad_sink(p)
            grads.append(buf if buf is not None else torch.empty_like(p))
        g = N.FlowGrads()
        L = wn.n_layers
        g.start_g, g.start_v, g.start_b = (N.fptr(t) for t in grads[0:3])
        for i in range(L):
            g.in_g[i], g.in_v[i], g.in_b[i] = (N.fptr(t) for t in grads[3 + 3 * i: 6 + 3 * i])
            o = 3 + 3 * L + 3 * i
            g.rs_g[i], g.rs_v[i], g.rs_b[i] = (N.fptr(t) for t in grads[o: o + 3])
        g.end_w, g.end_b = N.fptr(grads[3 + 6 * L]), N.fptr(grads[4 + 6 * L])
        dW = torch.empty_like(W) if ctx.has_W else None
        g.W = N.fptr(dW)
        dz_mid = torch.empty_like(z)
        dparams = torch.empty_like(z)
        dz_in = torch.empty_like(z) if ctx.has_W else dz_mid
        R, Dp = N.rows(batch, tp), N.round_up(d.D, 128)
        dctx_rows = torch.empty(R, Dp, dtype=torch.float32, device=z.device)
        scratch = _backward_scratch(lib.radmmm_flow_backward_scratch_bytes(mode, batch, tp, chans, d.D, d.H, d.L), z.device)
        N.check(lib.radmmm_flow_backward(C.byref(d), N.fptr(z), N.fptr(z_mid), N.fptr(p_out), N.fptr(dz_out),
                                         N.fptr(dlog_s), N.fptr(dz_mid), N.fptr(dparams), N.fptr(dz_in),
                                         N.fptr(dctx_rows), C.byref(g), N.ptr(scratch), N.stream()))
        dctx = None
        if ctx.needs_input_grad[4]:
            dctx = torch.empty(ctx.ctx_shape, dtype=torch.float32, device=z.device)
            N.check(lib.radmmm_context_rows_backward(N.fptr(dctx_rows), N.ptr(lens), batch, tp, d.D, N.fptr(dctx), 0,
                                                     N.stream()))
        return (None, None, None, dz_in, dctx, None, dW, None, *grads)


def _flow_apply(wn: WN, W, W_inv, mean, z, context, seq_lens, scaling_fn: str, precision: str, inverse: bool = False,
                want_params: bool = False):
    """Shared driver.  ``context`` is (B, D, Tp) like the reference's ``context_w_spkvec``."""
    if not z.is_cuda:
        raise RuntimeError("radmmm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    if z.device.index != torch.cuda.current_device():
        with torch.cuda.device(z.device):
            return _flow_apply(wn, W, W_inv, mean, z, context, seq_lens, scaling_fn, precision, inverse, want_params)
    mode = N.MODES[precision]
    batch, chans, tp = z.shape
    lens = _lens_of(seq_lens, batch, tp, z.device)
    ctx_btd = _as_btd(context.float())
    if ctx_btd.shape[1] != tp:
        raise RuntimeError(f"context has {ctx_btd.shape[1]} frames, z has {tp}")
    params = wn.raw_params()
    if inverse:
        lib = N.lib()
        z = z.contiguous().float()
        d = _make_desc(wn, mode, batch, chans, tp, scaling_fn, False, lens)
        with torch.no_grad():
            wn.fill_desc(d, params)
            wn.prepared(d, params)
            rows = _context_rows(ctx_btd, lens, mode)
            d.ctx_rows = N.ptr(rows.rows)
            ws = torch.empty(lib.radmmm_flow_workspace_bytes(mode, 0, batch, tp, chans, d.D, d.H, d.L),
                             dtype=torch.uint8, device=z.device)
            d.workspace = N.ptr(ws)
            d.side_stream = _side_stream(z.device) or None
            if W_inv is None:
                W_inv = torch.eye(chans, device=z.device)
            W_inv = W_inv.contiguous()
            d.W_inv, d.mean = N.fptr(W_inv), N.fptr(mean)
            p_out, z_tmp, z_out = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
            N.check(lib.radmmm_flow_inverse(C.byref(d), N.fptr(z), N.fptr(p_out), N.fptr(z_tmp), N.fptr(z_out),
                                            N.stream()))
        return z_out
    return FlowStepFunction.apply(wn, scaling_fn, mode, z.float(), ctx_btd, lens, W, mean, *params)


class AffineTransformationLayer(nn.Module):
    """common.py:1093-1185 for ``affine_model='wavenet'`` (the only model RADMMMFlow builds, decoders.py:63-66)."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, affine_model="simple_conv", with_dilation=True,
                 kernel_size=5, scaling_fn="exp", affine_activation="softplus", n_channels=1024,
                 use_partial_padding=False):
        super().__init__()
        if affine_model not in ("wavenet", "simple_conv", "film_stack"):
            raise Exception("{} affine model not supported".format(affine_model))
        if isinstance(scaling_fn, list) or scaling_fn not in ("translate", "exp", "tanh", "sigmoid"):
            if isinstance(scaling_fn, list):
                raise NotImplementedError("per-channel scaling_fn lists are not built in radmmm_b200")
            raise Exception("{} scaling fn not supported".format(scaling_fn))
        if affine_model != "wavenet":
            raise NotImplementedError("radmmm_b200 builds affine_model='wavenet' (what RADMMMFlow instantiates)")
        self.affine_model = affine_model
        self.scaling_fn = scaling_fn
        self.affine_param_predictor = WN(int(n_mel_channels / 2), n_context_dim, n_layers=n_layers,
                                         n_channels=n_channels, affine_activation=affine_activation,
                                         use_partial_padding=use_partial_padding)
        self.n_mel_channels = n_mel_channels
        self.precision = _DEFAULT_PRECISION

    def forward(self, z, context, inverse=False, seq_lens=None):
        if inverse:
            return _flow_apply(self.affine_param_predictor, None, None, None, z, context, seq_lens, self.scaling_fn,
                               self.precision, inverse=True)
        z_out, log_s, _ = _flow_apply(self.affine_param_predictor, None, None, None, z, context, seq_lens,
                                      self.scaling_fn, self.precision)
        return z_out, log_s


# --------------------------------------------------------------------------------------------- invertible 1x1 convs
class _Inv1x1Function(torch.autograd.Function):
    """out = W (z - pre) + post on (B, C, T);  common.py:540-548 / 605-617."""

    @staticmethod
    def forward(ctx, z, W, pre, post, lens):
        lib = N.lib()
        z = z.contiguous().float()
        W = W.contiguous().float()
        b, c, t = z.shape
        out = torch.empty(b, W.shape[0], t, dtype=torch.float32, device=z.device)
        N.check(lib.radmmm_inv1x1(N.fptr(z), N.fptr(W), N.fptr(pre), N.fptr(post), N.fptr(out), b, c, W.shape[0], t,
                                  N.stream()))
        ctx.save_for_backward(z, W, pre if pre is not None else z.new_empty(0), lens)
        ctx.has_pre = pre is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = N.lib()
        z, W, pre, lens = ctx.saved_tensors
        b, c, t = z.shape
        dout = dout.contiguous()
        dz = torch.empty_like(z)
        Wt = W.t().contiguous()
        N.check(lib.radmmm_inv1x1(N.fptr(dout), N.fptr(Wt), None, None, N.fptr(dz), b, W.shape[0], c, t, N.stream()))
        dW = torch.empty_like(W)
        N.check(lib.radmmm_inv1x1_wgrad(N.fptr(dout), N.fptr(z), N.fptr(pre) if ctx.has_pre else None, N.ptr(lens),
                                        N.fptr(dW), b, c, t, N.stream()))
        return dz, dW, None, None, None


class _InvertibleBase(nn.Module):
    cache_inverse: bool

    def _weight(self) -> torch.Tensor:
        raise NotImplementedError

    def _compute_inverse(self) -> torch.Tensor:
        raise NotImplementedError

    def _inverse_weight(self) -> torch.Tensor:
        """W^-1 for the inference pass (common.py:532-537, 599-604), cached when ``cache_inverse`` is set.  Built from
        triangular solves of the LU factors (cuBLAS trsm) rather than a dense inverse: no cuSOLVER on the path."""
        if self.cache_inverse and getattr(self, "W_inverse", None) is not None:
            return self.W_inverse
        with torch.no_grad():
            w_inv = self._compute_inverse().float().contiguous()
        if self.cache_inverse:
            self.W_inverse = w_inv
        return w_inv

    def log_det(self) -> torch.Tensor:
        return torch.sum(torch.log(torch.abs(self.upper_diag)))


class Invertible1x1ConvLUS(_InvertibleBase):
    """common.py:507-548: W = P (L + I) (U + diag(d)), log|det W| = sum log|d|."""

    def __init__(self, c, cache_inverse=False):
        super().__init__()
        W = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(W) < 0:
            W[:, 0] = -1 * W[:, 0]
        p, lower, upper = torch.linalg.lu(W)
        self.register_buffer("p", p)
        self.register_buffer("lower_diag", torch.ones(c))
        self.lower = nn.Parameter(torch.tril(lower, -1))
        self.upper_diag = nn.Parameter(torch.diag(upper).clone())
        self.upper = nn.Parameter(torch.triu(upper, 1))
        self.cache_inverse = cache_inverse

    def _weight(self):
        U = torch.triu(self.upper, 1) + torch.diag(self.upper_diag)
        Lm = torch.tril(self.lower, -1) + torch.diag(self.lower_diag)
        return torch.mm(self.p, torch.mm(Lm, U))

    def _compute_inverse(self):
        # W = P L U  ->  W^-1 = U^-1 L^-1 P^T
        c = self.upper.shape[0]
        eye = torch.eye(c, device=self.upper.device, dtype=torch.float64)
        U = (torch.triu(self.upper, 1) + torch.diag(self.upper_diag)).double()
        Lm = (torch.tril(self.lower, -1) + torch.diag(self.lower_diag)).double()
        u_inv = torch.linalg.solve_triangular(U, eye, upper=True)
        l_inv = torch.linalg.solve_triangular(Lm, eye, upper=False)
        return u_inv @ l_inv @ self.p.double().t()

    def forward(self, z, inverse=False, lens=None):
        b, c, t = z.shape
        ln = _lens_of(lens, b, t, z.device)
        if inverse:
            return _Inv1x1Function.apply(z, self._inverse_weight(), None, None, ln)
        return _Inv1x1Function.apply(z, self._weight(), None, None, ln), self.log_det()


class DataInitializedInvertible1x1Conv(_InvertibleBase):
    """common.py:551-617: z <- U (z - mean) with a one-off data-dependent whitening initialisation."""

    def __init__(self, c, cache_inverse=False):
        super().__init__()
        self.register_buffer("input_mean", torch.zeros(c, 1))
        self.register_buffer("initialized", torch.tensor(False))
        W = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(W) < 0:
            W[:, 0] = -1 * W[:, 0]
        p, _, upper = torch.linalg.lu(W)
        self.register_buffer("p", p)
        self.upper_diag = nn.Parameter(torch.diag(upper).clone())
        self.upper = nn.Parameter(torch.triu(upper, 1))
        self.cache_inverse = cache_inverse

    def initialize(self, data, lens):
        """common.py:569-591: mean / covariance over valid frames, W = chol(cov^-1) (upper); rank 0 broadcasts."""
        with torch.no_grad():
            lengths = lens.lengths if hasattr(lens, "lengths") else lens
            mask = get_mask_from_lengths(lengths.to(data.device), data.shape[2])[:, None].to(data.dtype)
            n = mask.sum()
            mean = (data * mask).sum((0, 2)) / n
            cen = (data - mean[None, :, None]) * mask
            covar = torch.einsum("bct,bdt->cd", cen, cen) / n
            self.covar = covar
            # one-off 160x160 factorisation: done on the host in fp64 (keeps cuSOLVER off the GPU path)
            wm = torch.linalg.cholesky(torch.linalg.inv(covar.double().cpu()), upper=True).float().to(data.device).contiguous()
            mean = mean[:, None].contiguous()
            if dist.is_available() and dist.is_initialized():
                dist.broadcast(wm, 0)
                dist.broadcast(mean, 0)
            self.input_mean.copy_(mean)
            self.upper_diag.copy_(torch.diag(wm))
            self.upper.copy_(torch.triu(wm, 1))
            self.initialized.fill_(True)

    def maybe_initialize(self, z, lens):
        """Runs :meth:`initialize` on the first training forward (common.py:594-597).  The device-side flag is read
        once; afterwards a host flag short-circuits the check so the hot path has no device sync."""
        if not self.training or getattr(self, "_init_seen", False):
            return
        if not bool(self.initialized):
            self.initialize(z, lens)
            print("initialized invertible conv")
        self._init_seen = True

    def _weight(self):
        return torch.triu(self.upper, 1) + torch.diag(self.upper_diag)

    def _compute_inverse(self):
        c = self.upper.shape[0]
        eye = torch.eye(c, device=self.upper.device, dtype=torch.float64)
        return torch.linalg.solve_triangular(self._weight().double(), eye, upper=True)

    def forward(self, z, inverse=False, lens=None):
        b, c, t = z.shape
        ln = _lens_of(lens, b, t, z.device)
        mean = self.input_mean.reshape(-1).contiguous()
        if inverse:
            return _Inv1x1Function.apply(z, self._inverse_weight(), None, mean, ln)
        self.maybe_initialize(z, lens if lens is not None else ln)
        return _Inv1x1Function.apply(z, self._weight(), mean, None, ln), self.log_det()


class Invertible1x1Conv(nn.Module):
    """common.py:621-662: plain-W variant (never instantiated by the shipped configs), log-det via ``logdet``."""

    def __init__(self, c, cache_inverse=False):
        super().__init__()
        self.conv = nn.Conv1d(c, c, kernel_size=1, stride=1, padding=0, bias=False)
        W = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(W) < 0:
            W[:, 0] = -1 * W[:, 0]
        self.conv.weight.data = W.view(c, c, 1).contiguous()
        self.cache_inverse = cache_inverse

    def forward(self, z, inverse=False):
        W = self.conv.weight.squeeze(-1)
        ln = _lens_of(None, z.shape[0], z.shape[2], z.device)
        if inverse:
            if not (self.cache_inverse and getattr(self, "W_inverse", None) is not None):
                # dense W: invert once on the host in fp64 (a 160x160 matrix), no cuSOLVER on the GPU path
                self.W_inverse = torch.linalg.inv(W.detach().double().cpu()).float().to(W.device)
            w_inv = self.W_inverse
            if not self.cache_inverse:
                self.W_inverse = None
            return _Inv1x1Function.apply(z, w_inv, None, None, ln)
        return _Inv1x1Function.apply(z, W, None, None, ln), torch.logdet(W).clone()


# --------------------------------------------------------------------------------------------- soft attention
class _SoftAttentionFunction(torch.autograd.Function):
    """(q (B,Ca,T1), k (B,Ca,T2), prior (B,T1,T2) | None, in_lens, txt_enc (B,Dt,T2) | None) ->
    (attn (B,1,T1,T2), attn_logprob (B,1,T1,T2), context (B,Dt,T1) | empty).  common.py:1259-1276 +
    tts_lightning_modules.py:670 in one kernel each way."""

    @staticmethod
    def forward(ctx, q, k, prior, in_lens, txt_enc, temperature: float):
        lib = N.lib()
        q, k = q.contiguous().float(), k.contiguous().float()
        b, ca, t1 = q.shape
        t2 = k.shape[2]
        prior_c = prior.contiguous().float() if prior is not None else None
        txt_c = txt_enc.contiguous().float() if txt_enc is not None else None
        dt = txt_c.shape[1] if txt_c is not None else 0
        attn = torch.empty(b, 1, t1, t2, device=q.device)
        logp = torch.empty(b, 1, t1, t2, device=q.device)
        context = torch.empty(b, dt, t1, device=q.device) if txt_c is not None else q.new_empty(0)
        N.check(lib.radmmm_soft_attention(N.fptr(q), N.fptr(k), N.fptr(prior_c), N.ptr(in_lens), N.fptr(attn), N.fptr(logp),
                                          N.fptr(txt_c), N.fptr(context) if txt_c is not None else None, b, ca, t1, t2, dt,
                                          temperature, N.stream()))
        ctx.save_for_backward(q, k, prior_c if prior_c is not None else q.new_empty(0), in_lens, attn,
                              txt_c if txt_c is not None else q.new_empty(0))
        ctx.has_prior, ctx.has_txt, ctx.temperature = prior_c is not None, txt_c is not None, temperature
        return attn, logp, context

    @staticmethod
    def backward(ctx, dattn, dlogp, dcontext):
        lib = N.lib()
        q, k, prior, in_lens, attn, txt = ctx.saved_tensors
        b, ca, t1 = q.shape
        t2 = k.shape[2]
        dt = txt.shape[1] if ctx.has_txt else 0
        dattn = dattn.contiguous() if dattn is not None else None
        dlogp = dlogp.contiguous() if dlogp is not None else None
        dctx = dcontext.contiguous() if (ctx.has_txt and dcontext is not None) else None
        use_txt = ctx.has_txt and dctx is not None
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        dtxt = torch.empty_like(txt) if use_txt else None
        nws = lib.radmmm_soft_attention_backward_workspace_bytes(b, t1, t2)
        ws = torch.empty(nws, dtype=torch.uint8, device=q.device)
        N.check(lib.radmmm_soft_attention_backward(
            N.fptr(q), N.fptr(k), N.fptr(prior) if ctx.has_prior else None, N.ptr(in_lens), N.fptr(attn), N.fptr(dattn),
            N.fptr(dlogp), N.fptr(txt) if use_txt else None, N.fptr(dctx), N.fptr(dq), N.fptr(dk), N.fptr(dtxt), b, ca, t1, t2,
            dt, ctx.temperature, N.ptr(ws), nws, N.stream()))
        return dq, dk, None, None, dtxt, None


class _PlainConvNorm(nn.Module):
    """ConvNorm without partial padding (common.py:152-191): weight-normed Conv1d with 'same' padding.  The attention's
    projections are three tiny convolutions per side; they stay library (cuDNN) calls."""

    def __init__(self, cin, cout, kernel_size, w_init_gain="linear"):
        super().__init__()
        conv = nn.Conv1d(cin, cout, kernel_size, padding=(kernel_size - 1) // 2)
        nn.init.xavier_uniform_(conv.weight, gain=nn.init.calculate_gain(w_init_gain))
        self.conv = nn.utils.weight_norm(conv)

    def forward(self, x):
        return self.conv(x)


class ConvAttention(nn.Module):
    """common.py:1188-1277: Gaussian-distance soft attention between mel frames (queries) and text tokens (keys)."""

    def __init__(self, n_mel_channels=80, n_text_channels=512, n_att_channels=80, temperature=1.0):
        super().__init__()
        self.temperature = temperature
        self.key_proj = nn.Sequential(_PlainConvNorm(n_text_channels, n_text_channels * 2, 3, "relu"), nn.ReLU(),
                                      _PlainConvNorm(n_text_channels * 2, n_att_channels, 1))
        self.query_proj = nn.Sequential(_PlainConvNorm(n_mel_channels, n_mel_channels * 2, 3, "relu"), nn.ReLU(),
                                        _PlainConvNorm(n_mel_channels * 2, n_mel_channels, 1), nn.ReLU(),
                                        _PlainConvNorm(n_mel_channels, n_att_channels, 1))

    def _lens(self, keys, mask, key_lens):
        if key_lens is not None:
            return key_lens.to(device=keys.device, dtype=torch.int32).contiguous()
        if mask is not None:       # mask: (B, T2, 1), True on padded tokens (tts_lightning_modules.py:451)
            return (~mask.reshape(mask.shape[0], -1).bool()).sum(1).to(torch.int32)
        return torch.full((keys.shape[0],), keys.shape[2], dtype=torch.int32, device=keys.device)

    def forward(self, queries, keys, query_lens, mask=None, key_lens=None, attn_prior=None):
        """Same arguments and outputs as the reference: (attn (B,1,T1,T2), attn_logprob (B,1,T1,T2))."""
        keys_enc = self.key_proj(keys)
        queries_enc = self.query_proj(queries)
        attn, logp, _ = _SoftAttentionFunction.apply(queries_enc, keys_enc, attn_prior, self._lens(keys, mask, key_lens),
                                                     None, 0.0005)
        return attn, logp

    def forward_with_context(self, queries, keys, txt_enc, mask=None, key_lens=None, attn_prior=None):
        """Fused variant: also returns context = bmm(txt_enc, attn^T) (tts_lightning_modules.py:670) from the same kernel."""
        keys_enc = self.key_proj(keys)
        queries_enc = self.query_proj(queries)
        return _SoftAttentionFunction.apply(queries_enc, keys_enc, attn_prior, self._lens(keys, mask, key_lens), txt_enc,
                                            0.0005)
