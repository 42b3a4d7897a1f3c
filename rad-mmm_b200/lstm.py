"""Context bi-LSTM on the native kernels (reference: models/radmmm.py:137-146 -- pack_padded_sequence ->
nn.LSTM(bidirectional=True, batch_first=True) -> pad_packed_sequence).

The parameters stay in an ordinary ``nn.LSTM`` module (same ``state_dict`` keys as the reference); only the
computation is replaced: input projections for all frames are one contraction, the recurrence is one persistent
cooperative kernel per pass (csrc/lstm.cu), and the weight / input gradients are contractions over the saved gate
gradients.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import _native as N

_wcache = {}


def _prepared_weights(lstm: torch.nn.LSTM, mode: int, params, training: bool):
    """[W_ih_fwd; W_ih_rev] zero-padded to (8H_pad, In_pad) and its transpose, in the act format of ``mode``.
    Training forwards redo it every call (optimizers that write ``p.data`` do not bump version counters, so a
    version-keyed cache would go stale -- see common.WN.prepared); inference caches on (data_ptr, version), and a
    training call leaves that cache marked stale."""
    lib = N.lib()
    wih_f, wih_r = params[0], params[4]
    key = None if training else (id(lstm), mode, wih_f.data_ptr(), wih_f._version, wih_r.data_ptr(), wih_r._version)
    hit = _wcache.get(id(lstm))
    if key is not None and hit is not None and hit[0] == key:
        return hit[1]
    h4, n_in = wih_f.shape
    inp = N.round_up(n_in, 128)
    n8 = N.round_up(2 * h4, 128)
    with torch.no_grad():
        w = torch.zeros(n8, inp, device=wih_f.device)
        w[:h4, :n_in] = wih_f
        w[h4:2 * h4, :n_in] = wih_r
        wt = w.t().contiguous()
        planes = 2 if mode == N.MODE_BF16X3 else 1
        if mode == N.MODE_F32:
            wa, wta = w, wt
        else:
            wa = torch.empty(planes * w.numel(), dtype=torch.bfloat16, device=w.device)
            wta = torch.empty(planes * w.numel(), dtype=torch.bfloat16, device=w.device)
            N.check(lib.radmmm_cast_rows(mode, N.fptr(w), w.numel(), N.ptr(wa), w.numel(), N.stream()))
            N.check(lib.radmmm_cast_rows(mode, N.fptr(wt), wt.numel(), N.ptr(wta), wt.numel(), N.stream()))
    out = (wa, wta, inp, n8)
    _wcache[id(lstm)] = (key, out)
    return out


def _rows(mode: int, x_btd: torch.Tensor, lens: torch.Tensor):
    """(B, T, D) fp32 -> act rows [R][round_up(D,128)] (zero beyond each length / in the gap rows)."""
    lib = N.lib()
    b, t, d = x_btd.shape
    buf = torch.empty(lib.radmmm_context_rows_bytes(mode, b, t, d), dtype=torch.uint8, device=x_btd.device)
    N.check(lib.radmmm_context_rows(mode, N.fptr(x_btd), N.ptr(lens), b, t, d, N.ptr(buf), N.stream()))
    return buf


def _cast(mode: int, x: torch.Tensor):
    if mode == N.MODE_F32:
        return x
    lib = N.lib()
    planes = 2 if mode == N.MODE_BF16X3 else 1
    buf = torch.empty(planes * x.numel(), dtype=torch.bfloat16, device=x.device)
    N.check(lib.radmmm_cast_rows(mode, N.fptr(x), x.numel(), N.ptr(buf), x.numel(), N.stream()))
    return buf


_tail_streams = {}


def _tail_stream(device):
    """Side stream for the independent tail of the LSTM backward (RADMMM_B200_LSTM_TAIL_STREAM=0: everything on one stream)."""
    if os.environ.get("RADMMM_B200_LSTM_TAIL_STREAM", "1") == "0":
        return None
    # one side stream per CALLING stream: independent LSTMs running on different streams (the attribute predictors of the joint
    # step) must not meet on a shared one
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    if key not in _tail_streams:
        _tail_streams[key] = torch.cuda.Stream(device=device)
    return _tail_streams[key]


class ContextLSTMFunction(torch.autograd.Function):
    """x (B, T, In) fp32, grouped lengths (B) int32 -> (B, T, 2H); zero beyond each length."""

    @staticmethod
    def forward(ctx, lstm, mode, x, lens, wih_f, whh_f, bih_f, bhh_f, wih_r, whh_r, bih_r, bhh_r):
        lib = N.lib()
        x = x.contiguous().float()
        b, t, n_in = x.shape
        hid = whh_f.shape[1]
        r = N.rows(b, t)
        params = (wih_f, whh_f, bih_f, bhh_f, wih_r, whh_r, bih_r, bhh_r)
        wa, wta, inp, n8 = _prepared_weights(lstm, mode, params, training=any(ctx.needs_input_grad))
        ctx.wta = wta
        x_rows = _rows(mode, x, lens)
        bias = torch.zeros(n8, device=x.device)
        bias[:4 * hid] = bih_f + bhh_f
        bias[4 * hid:8 * hid] = bih_r + bhh_r
        xproj = torch.empty(r, 8 * hid, device=x.device)
        N.check(lib.radmmm_conv_rows(mode, N.ptr(x_rows), inp, r * inp, N.ptr(wa), inp, n8 * inp, 0, N.fptr(bias),
                                     N.fptr(xproj), 8 * hid, r, inp, 8 * hid, 1, 1, N.stream()))
        out = torch.zeros(b, t, 2 * hid, device=x.device)
        gates = torch.empty(r, 8 * hid, device=x.device)
        cstate = torch.empty(r, 2 * hid, device=x.device)
        ws = torch.empty(lib.radmmm_lstm_workspace_bytes(b, hid), dtype=torch.uint8, device=x.device)
        whf, whr = whh_f.contiguous(), whh_r.contiguous()
        N.check(lib.radmmm_lstm_forward(mode, N.fptr(xproj), N.fptr(whf), N.fptr(whr), N.ptr(lens), b, t, hid, N.fptr(out),
                                        N.fptr(gates), N.fptr(cstate), N.ptr(ws), N.stream()))
        ctx.lstm, ctx.mode = lstm, mode
        ctx.dims = (b, t, n_in, hid, r, inp, n8)
        ctx.save_for_backward(x_rows, lens, gates, cstate, out, whf, whr, wih_f, wih_r)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = N.lib()
        x_rows, lens, gates, cstate, out, whf, whr, wih_f, wih_r = ctx.saved_tensors
        mode = ctx.mode
        b, t, n_in, hid, r, inp, n8 = ctx.dims
        dev = dout.device
        dout = dout.contiguous().float()
        dg = torch.zeros(r, 8 * hid, device=dev)
        ws = torch.empty(lib.radmmm_lstm_workspace_bytes(b, hid), dtype=torch.uint8, device=dev)
        N.check(lib.radmmm_lstm_backward(mode, N.fptr(dout), N.fptr(gates), N.fptr(cstate), N.fptr(whf), N.fptr(whr),
                                         N.ptr(lens), b, t, hid, N.fptr(dg), N.ptr(ws), N.stream()))
        wta = ctx.wta                       # the transposed input weights the forward pass prepared
        dg_act = _cast(mode, dg)
        k8 = 8 * hid
        # dX = dG . [W_ih_f; W_ih_r]  (row GEMM, K = 8H) -- only when the input wants a gradient: in decoder-only training the
        # LSTM input (text encoding, speaker vector, f0, energy) is data, and this GEMM sits on the step's exposed tail
        dx = None
        if ctx.needs_input_grad[2]:
            dx_rows = torch.empty(r, inp, device=dev)
            N.check(lib.radmmm_conv_rows(mode, N.ptr(dg_act), k8, r * k8, N.ptr(wta), n8, inp * n8, 0, None, N.fptr(dx_rows),
                                         inp, r, k8, inp, 1, 1, N.stream()))
            dx = torch.empty(b, t, n_in, device=dev)
            N.check(lib.radmmm_context_rows_backward(N.fptr(dx_rows), N.ptr(lens), b, t, n_in, N.fptr(dx), 0, N.stream()))
        # The three weight-gradient contractions and the bias reduction are independent and sit on the exposed tail of the
        # training step (nothing else is left to run): dW_hh and the bias sums go to a side stream next to dW_ih.
        main = torch.cuda.current_stream(dev)
        side = _tail_stream(dev)
        hp = N.round_up(2 * hid, 128)
        es = 4 if mode == N.MODE_F32 else 2
        npad = N.round_up(hid, 128)
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            # dW_hh[dir] = dG_dir^T H_prev_dir: the previous state of the forward direction is row r-1, of the reverse
            # direction row r+1 (rows beyond a sequence are zero, which is exactly h_{-1} = 0)
            h_rows = _rows(mode, out, lens)
            dwhh_f = torch.empty(4 * hid, npad, device=dev)
            dwhh_r = torch.empty(4 * hid, npad, device=dev)
            N.check(lib.radmmm_wgrad_rows(mode, N.ptr(dg_act), k8, r * k8, N.ptr(h_rows), hp, r * hp, N.fptr(dwhh_f), npad,
                                          4 * hid * npad, r, 4 * hid, hid, 1, 1, -1, N.stream()))
            N.check(lib.radmmm_wgrad_rows(mode, dg_act.data_ptr() + 4 * hid * es, k8, r * k8, h_rows.data_ptr() + hid * es,
                                          hp, r * hp, N.fptr(dwhh_r), npad, 4 * hid * npad, r, 4 * hid, hid, 1, 1, 1,
                                          N.stream()))
            db = dg.sum(0)
            if side is not None:          # allocated on the side stream, consumed on the main one
                for t_ in (dwhh_f, dwhh_r, db):
                    t_.record_stream(main)
                for t_ in (dg_act, dg, out, h_rows):      # and the other way round
                    t_.record_stream(side)
        # dW_ih = dG^T X
        dwih = torch.empty(k8, inp, device=dev)
        N.check(lib.radmmm_wgrad_rows(mode, N.ptr(dg_act), k8, r * k8, N.ptr(x_rows), inp, r * inp, N.fptr(dwih), inp,
                                      k8 * inp, r, k8, inp, 1, 1, 0, N.stream()))
        if side is not None:
            main.wait_stream(side)
        dbf, dbr = db[:4 * hid].contiguous(), db[4 * hid:].contiguous()
        return (None, None, dx, None, dwih[:4 * hid, :n_in].contiguous(), dwhh_f[:, :hid].contiguous(), dbf, dbf.clone(),
                dwih[4 * hid:, :n_in].contiguous(), dwhh_r[:, :hid].contiguous(), dbr, dbr.clone())


def context_lstm(lstm: torch.nn.LSTM, x_btd: torch.Tensor, lens_g: torch.Tensor, precision: str) -> torch.Tensor:
    """Drop-in for the packed bi-LSTM call.  Large batches are processed in chunks (sequences are independent).
    ``precision`` "bf16": contractions in bf16 and the cluster-resident tensor-core recurrence (bf16 W_hh / h, fp32 state);
    "bf16x3" / "fp32": fp32-grade contractions and the fp32 recurrence."""
    if not x_btd.is_cuda:
        raise RuntimeError("radmmm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    if x_btd.device.index != torch.cuda.current_device():
        with torch.cuda.device(x_btd.device):
            return context_lstm(lstm, x_btd, lens_g, precision)
    mode = N.MODES[precision]
    lens = lens_g.to(device=x_btd.device, dtype=torch.int32).contiguous()
    p = (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
         lstm.weight_ih_l0_reverse, lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0_reverse)
    hid = lstm.hidden_size
    hp = N.round_up(hid, 16)
    if hp != hid:
        # the kernels want a hidden size that is a multiple of 16 (8H a multiple of 128): run a zero-padded LSTM -- a padded
        # unit has zero weights and biases, so its gates are (1/2, 1/2, 0, 1/2), its cell and output stay exactly 0 and
        # nothing depends on it -- and slice its outputs away (RADTTS variant: hidden 524, configs/RADTTS_model_config.yaml)
        p = _pad_hidden(p, hid, hp)
        y = _run_chunks(lstm, mode, x_btd, lens, p)
        return torch.cat((y[..., :hid], y[..., hp:hp + hid]), dim=-1)
    return _run_chunks(lstm, mode, x_btd, lens, p)


def _pad_hidden(p, hid: int, hp: int):
    """Zero-pad every gate block of (W_ih, W_hh, b_ih, b_hh) x 2 directions from ``hid`` to ``hp`` units (differentiable)."""
    F = torch.nn.functional
    out = []
    for i, t in enumerate(p):
        kind = i % 4
        if kind == 0:                                   # W_ih (4H, In)
            t = F.pad(t.reshape(4, hid, -1), (0, 0, 0, hp - hid)).reshape(4 * hp, -1)
        elif kind == 1:                                 # W_hh (4H, H)
            t = F.pad(t.reshape(4, hid, hid), (0, hp - hid, 0, hp - hid)).reshape(4 * hp, hp)
        else:                                           # biases (4H)
            t = F.pad(t.reshape(4, hid), (0, hp - hid)).reshape(4 * hp)
        out.append(t.contiguous())
    return tuple(out)


def _run_chunks(lstm, mode, x_btd, lens, p):
    outs = []
    chunk = 64      # sequences per launch of the recurrence kernels (csrc/lstm_cluster.cu / lstm.cu)
    for s in range(0, x_btd.shape[0], chunk):
        outs.append(ContextLSTMFunction.apply(lstm, mode, x_btd[s:s + chunk], lens[s:s + chunk].contiguous(), *p))
    return outs[0] if len(outs) == 1 else torch.cat(outs, 0)
