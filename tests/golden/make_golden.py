"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Usage:  python tests/golden/make_golden.py
Inputs and weights are NOT stored: they are pure functions of their names (radmmm_b200.synthetic), so the
fixtures hold reference OUTPUTS only and stay small.  The reference needs three shims to import here
(SURVEY.md 8c): empty ``matplotlib`` modules (alignment.py:23), ``vocoders/`` on sys.path (decoders.py:30-31),
and -- for audio_processing.py only -- a stand-in ``librosa`` (see ``_stub_librosa``).
"""
import json
import math
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path[:0] = [REF, os.path.join(REF, "vocoders")]
for name in ("matplotlib", "matplotlib.pylab"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
warnings.filterwarnings("ignore")

from radmmm_b200 import synthetic as syn          # noqa: E402
from oracle.frontend import slaney_mel_basis      # noqa: E402  (only to stand in for the absent librosa)


def _stub_librosa():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filt = types.ModuleType("librosa.filters")

    def pad_center(data, size, **kw):
        assert len(data) == size            # win_length == filter_length in every shipped config
        return data
    util.pad_center = pad_center
    util.tiny = lambda x: np.finfo(np.float32).tiny
    util.normalize = lambda x, **kw: x
    filt.mel = lambda sr, n_fft, n_mels, fmin, fmax: slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    lib.util, lib.filters = util, filt
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})


def npz(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(f"{name}: {os.path.getsize(os.path.join(HERE, name)) / 1024:.0f} KiB")


def grad_checksums(named_params):
    rows, names = [], []
    for n, p in named_params:
        if p.grad is None:
            continue
        g = p.grad.double().flatten()
        probe = syn.hash_uniform("probe." + n, (g.numel(),)).double()
        rows.append([g.sum().item(), g.abs().sum().item(), (g * probe).sum().item()])
        names.append(n)
    return names, np.asarray(rows)


def decoder_case(fname, n_flows, batch, frames, n_splines=0):
    from decoders import RADMMMFlow
    from common import SequenceLength
    from loss import compute_flow_loss
    torch.manual_seed(0)
    dec = RADMMMFlow(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2,
                     n_mel_channels=80, n_flows=n_flows, n_conv_layers_per_step=4, n_early_size=2,
                     n_early_every=2, affine_model="wavenet", scaling_fn="tanh", affine_activation="softplus",
                     use_partial_padding=True, n_splines=n_splines)
    sd = syn.synthetic_state_dict(n_flows=n_flows, n_splines=n_splines)
    missing = dec.load_state_dict(sd, strict=True)
    dec.train()
    bt = syn.synthetic_batch(batch, frames, tag=fname)
    lens = bt["out_lens"]
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(lens), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    lens_g = lens // 2
    mask = (torch.arange(frames // 2)[None] < lens_g[:, None])[:, None].float()
    # NB compute_flow_loss accumulates INTO log_det_W_list[0] in place (loss.py:92-100): snapshot it first.
    log_det_snapshot = torch.stack([t.detach().clone() for t in out["log_det_W_list"]])
    # n_elements exactly as RADMMMLoss.forward computes it (loss.py:520): floor(sum(out_lens) / n_group_size)
    n_elements = torch.div(lens.sum(), 2, rounding_mode="floor")
    loss, loss_prior = compute_flow_loss(out["z_mel"], list(out["log_det_W_list"]), out["log_s_list"],
                                         n_elements, out["z_mel"].size(1), mask, 1.0)
    loss.backward()
    gnames, gsums = grad_checksums(dec.named_parameters())
    # inverse: replay decoders.py:227-246 with an injected residual (infer() itself hard-codes CUDA at :221)
    dec.eval()
    with torch.no_grad():
        residual = syn.hash_uniform("residual" + fname, (batch, 160, frames // 2), -1.5, 1.5)
        from common import SequenceLength as SL
        stack = dec.exit_steps.copy()
        mel = residual[:, len(stack) * 2:]
        rest = residual[:, :len(stack) * 2]
        ctx = out["context_w_spkvec"].detach()
        sl = SL(lens_g)
        for i, fs in enumerate(reversed(dec.flows)):
            cur = len(dec.flows) - i - 1
            mel = fs(mel, ctx, inverse=True, seq_lens=sl)
            if stack and cur == stack[-1]:
                stack.pop()
                mel = torch.cat((rest[:, len(stack) * 2:], mel), 1)
                rest = rest[:, :len(stack) * 2]
        mel_inv = dec.fold(mel)
    npz(fname, z_mel=out["z_mel"], log_s=torch.stack([ls.sum(1) for ls in out["log_s_list"]]) if n_splines == 0 else
        np.zeros(1), **{f"log_s_{i}": ls for i, ls in enumerate(out["log_s_list"])},
        log_det=log_det_snapshot, context=out["context_w_spkvec"][:, ::33],
        loss=loss, loss_prior=loss_prior, grad_names=np.asarray(gnames), grad_sums=gsums,
        mel_inv=mel_inv, meta=np.asarray([n_flows, batch, frames, n_splines]))
    return dec


def state_dict_keys():
    from decoders import RADMMMFlow
    specs = {}
    for tag, kw in {"radmmm": dict(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=8),
                    "radmmm_spline2": dict(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=4, n_splines=2),
                    "radtts_accent": dict(n_accent_dim=8, n_text_dim=512, n_group_size=2, n_flows=2,
                                          use_accent_emb_for_decoder=True)}.items():
        dec = RADMMMFlow(**kw)
        specs[tag] = {"init_args": kw,
                      "state": [[k, list(v.shape), str(v.dtype)] for k, v in dec.state_dict().items()],
                      "params": [n for n, _ in dec.named_parameters()],
                      "decoder_cond_dims": dec.decoder_cond_dims, "exit_steps": dec.exit_steps}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(specs, f)
    print("state_dict_keys.json written")


def op_cases():
    import common
    from common import WN, SequenceLength, AffineTransformationLayer, ConvAttention
    from common import Invertible1x1ConvLUS, DataInitializedInvertible1x1Conv, Invertible1x1Conv
    import splines
    from maskedbatchnorm1d import MaskedBatchNorm1d
    from loss import compute_flow_loss
    # --- WN at a small width, ragged lengths incl. a very short one (halo > length)
    torch.manual_seed(1)
    lens = torch.tensor([37, 20, 5])
    wn = WN(6, 10, n_layers=4, n_channels=64)
    for n, p in wn.named_parameters():
        p.data.copy_(syn.hash_uniform("wn_small." + n, tuple(p.shape), -0.3, 0.3))
    z0 = syn.hash_uniform("wn_small.z0", (3, 6, 37), -1, 1)
    ctx = syn.hash_uniform("wn_small.ctx", (3, 10, 37), -1, 1)
    y = wn((z0, ctx), seq_lens=SequenceLength(lens))
    # --- partial conv alone
    pc = common.ConvNorm(8, 12, kernel_size=5, dilation=4, use_partial_padding=True, use_weight_norm=True)
    for n, p in pc.named_parameters():
        p.data.copy_(syn.hash_uniform("pc." + n, tuple(p.shape), -0.5, 0.5))
    xin = syn.hash_uniform("pc.x", (3, 8, 37), -1, 1)
    mask = SequenceLength(lens).mask.unsqueeze(1).float()
    ypc = pc(xin, mask)
    # --- affine coupling with each scaling fn
    aff = {}
    for fn in ("tanh", "exp", "sigmoid"):
        layer = AffineTransformationLayer(12, 10, 2, affine_model="wavenet", scaling_fn=fn, n_channels=32,
                                          use_partial_padding=True)
        for n, p in layer.named_parameters():
            p.data.copy_(syn.hash_uniform("aff." + n, tuple(p.shape), -0.3, 0.3))
        zin = syn.hash_uniform("aff.z", (3, 12, 37), -1, 1)
        zo, ls = layer(zin, ctx, seq_lens=SequenceLength(lens))
        zi = layer(zo, ctx, inverse=True, seq_lens=SequenceLength(lens))
        aff[fn + "_z"], aff[fn + "_log_s"], aff[fn + "_inv"] = zo, ls, zi
    # --- invertible convs
    inv = {}
    for cls, tag in ((Invertible1x1ConvLUS, "lus"), (DataInitializedInvertible1x1Conv, "whiten")):
        m = cls(12)
        sd = syn.synthetic_state_dict(n_flows=2, n_mel_channels=6, n_group_size=2, tag="inv12")
        pre = "flows.1.invtbl_conv." if tag == "lus" else "flows.0.invtbl_conv."
        m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        zin = syn.hash_uniform("inv.z", (3, 12, 37), -2, 2)
        m.eval()
        zo, ld = m(zin) if tag == "lus" else m(zin, lens=SequenceLength(lens))
        zi = m(zo, inverse=True)
        inv[tag + "_z"], inv[tag + "_logdet"], inv[tag + "_inv"] = zo, ld, zi
    plain = Invertible1x1Conv(12)
    plain.conv.weight.data.copy_(syn.hash_uniform("inv.plainW", (12, 12, 1), -0.6, 0.6) + torch.eye(12)[..., None])
    zo, ld = plain(zin)
    inv["plain_z"], inv["plain_logdet"], inv["plain_inv"] = zo, ld, plain(zo, inverse=True)
    # data-dependent whitening init (common.py:569-591)
    w = DataInitializedInvertible1x1Conv(12)
    w.train()
    zdata = syn.hash_uniform("inv.init", (3, 12, 37), -2, 2) * syn.hash_uniform("inv.scale", (1, 12, 1), 0.2, 2.0)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        zo, ld = w(zdata, lens=SequenceLength(lens))
    inv["init_mean"], inv["init_upper"], inv["init_diag"], inv["init_z"] = w.input_mean, w.upper.data, w.upper_diag.data, zo
    # --- splines
    x = syn.hash_uniform("spl.x", (50, 7), -0.2, 1.2)
    wt = syn.hash_uniform("spl.w", (50, 7, 32), -2, 2)
    vt = syn.hash_uniform("spl.v", (50, 7, 33), -2, 2)
    yq, lj = splines.unbounded_piecewise_quadratic_transform(x, wt, vt)
    xq, _ = splines.unbounded_piecewise_quadratic_transform(yq, wt, vt, inverse=True)
    xl = x.clamp(0, 1)
    yl, ljl = splines.piecewise_linear_transform(xl, wt)
    xli, ljli = splines.piecewise_linear_inverse_transform(yl, wt)
    # --- masked BN (train + eval)
    bn = MaskedBatchNorm1d(8)
    bn.weight.data.copy_(syn.hash_uniform("bn.w", (8,), 0.5, 1.5))
    bn.bias.data.copy_(syn.hash_uniform("bn.b", (8,), -0.5, 0.5))
    bn.train()
    ybn = bn(xin, mask)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    bn.eval()
    ybn_eval = bn(xin, mask)
    # --- flow loss
    zz = syn.hash_uniform("loss.z", (3, 12, 37), -2, 2)
    lsl = [syn.hash_uniform(f"loss.ls{i}", (3, 6, 37), -1, 1) for i in range(3)]
    ldl = [syn.hash_uniform(f"loss.ld{i}", (), -1, 1) for i in range(3)]
    l, lp = compute_flow_loss(zz, [t.clone() for t in ldl], lsl, lens.sum(), 12, mask, 0.8)
    # --- attention
    att = ConvAttention(n_mel_channels=80, n_text_channels=24, n_att_channels=80)
    for n, p in att.named_parameters():
        p.data.copy_(syn.hash_uniform("att." + n, tuple(p.shape), -0.2, 0.2))
    q_in = syn.hash_uniform("att.q", (3, 80, 37), -1, 1)
    k_in = syn.hash_uniform("att.k", (3, 24, 11), -1, 1)
    in_lens = torch.tensor([11, 7, 3])
    prior = syn.hash_uniform("att.prior", (3, 37, 11), 0.0, 1.0)
    amask = (torch.arange(11)[None] < in_lens[:, None])[..., None] == 0
    a, alp = att(q_in, k_in, lens, amask, key_lens=in_lens, attn_prior=prior)
    a_np, alp_np = att(q_in, k_in, lens, amask, key_lens=in_lens, attn_prior=None)
    txt_enc = syn.hash_uniform("att.txt", (3, 24, 11), -1, 1)
    ctx_att = torch.bmm(txt_enc, a.squeeze(1).transpose(1, 2))
    npz("ops.npz", wn_y=y, pc_y=ypc, **aff, **inv, spl_yq=yq, spl_lj=lj, spl_xq=xq, spl_yl=yl, spl_ljl=ljl,
        spl_xli=xli, spl_ljli=ljli, bn_y=ybn, bn_rm=rm, bn_rv=rv, bn_y_eval=ybn_eval, loss=l, loss_prior=lp,
        att=a, att_logprob=alp, att_noprior=a_np, att_logprob_noprior=alp_np, att_ctx=ctx_att)


def spline_step_case():
    """One FlowStep with use_spline=True (decoders.py:51-61) at a narrow channel count, train mode."""
    from decoders import FlowStep
    from common import SequenceLength
    torch.manual_seed(2)
    lens = torch.tensor([21, 9])
    fs = FlowStep(8, 10, 2, mode="LUS", use_partial_padding=True, use_spline=True, use_bn=True)
    sd = fs.state_dict()
    for k in sd:
        if sd[k].dtype == torch.float32 and sd[k].numel() > 1 and not k.endswith("invtbl_conv.p") \
                and "lower_diag" not in k and "running_var" not in k:
            sd[k] = syn.hash_uniform("splstep." + k, tuple(sd[k].shape), -0.3, 0.3)
    sd["invtbl_conv.upper_diag"] = syn.hash_uniform("splstep.ud", (8,), 0.7, 1.3)
    sd["invtbl_conv.p"] = torch.eye(8)[syn.hash_permutation("splstep.p", 8)]
    sd["invtbl_conv.upper"] = torch.triu(sd["invtbl_conv.upper"], 1)
    sd["invtbl_conv.lower"] = torch.tril(sd["invtbl_conv.lower"], -1)
    fs.load_state_dict(sd)
    fs.train()
    z = syn.hash_uniform("splstep.z", (2, 8, 21), -3.5, 3.5)
    ctx = syn.hash_uniform("splstep.ctx", (2, 10, 21), -1, 1)
    zo, ld, ls = fs(z, ctx, seq_lens=SequenceLength(lens))
    rm = fs.coupling_tfn.param_predictor.in_layers[0].bn.running_mean.clone()
    fs.eval()
    zo_eval, _, ls_eval = fs(z, ctx, seq_lens=SequenceLength(lens))
    zi = fs(zo_eval, ctx, inverse=True, seq_lens=SequenceLength(lens))
    keys = [[k, list(v.shape)] for k, v in sd.items()]
    with open(os.path.join(HERE, "spline_step_keys.json"), "w") as f:
        json.dump(keys, f)
    npz("spline_step.npz", z=zo, log_det=ld, log_s=ls, rm0=rm, z_eval=zo_eval, log_s_eval=ls_eval, z_inv=zi)


def frontend_case():
    _stub_librosa()
    from audio_processing import TacotronSTFT
    for sr, tag in ((22050, "22k"), (16000, "16k")):
        stft = TacotronSTFT(1024, 256, 1024, 80, sr, 0.0, 8000.0)
        n = 256 * 24
        t = torch.arange(n) / sr
        y = 0.5 * syn.hash_uniform("audio" + tag, (2, n), -1, 1)
        for f, a in ((220.0, 0.3), (1333.0, 0.2), (5200.0, 0.1)):
            y = y + a * torch.sin(2 * math.pi * f * t)[None]
        y = y.clamp(-1, 1)
        mag, _ = stft.stft_fn.transform(y)
        mel = stft.mel_spectrogram(y)
        npz(f"frontend_{tag}.npz", mag=mag[:, ::8], mel=mel, meta=np.asarray([sr, n]))


def radam_case():
    """Reference radam.RAdam (radam.py:45-142) for 9 steps -- across the N_sma >= 5 switch -- with the gradient clipping
    Lightning applies before every step (configs/RADMMM_train_config.yaml:7-8: clip_grad_norm_(params, 1.0)) and the shipped
    weight decay (configs/RADMMM_model_config.yaml:64)."""
    from radam import RAdam
    shapes = {"a": (37, 19), "b": (1024,), "c": (5, 7, 3), "d": (1,)}
    params = [torch.nn.Parameter(syn.hash_uniform("radam.p." + k, s, -1, 1)) for k, s in shapes.items()]
    opt = RAdam(params, lr=1e-3, weight_decay=1e-6)
    traj, norms = [], []
    for step in range(9):
        for (k, s), p in zip(shapes.items(), params):
            p.grad = syn.hash_uniform(f"radam.g{step}." + k, s, -1, 1) * (3.0 if step % 2 == 0 else 0.01)
        norms.append(torch.nn.utils.clip_grad_norm_(params, 1.0).item())
        opt.step()
        traj.append(torch.cat([p.detach().flatten() for p in params]).clone())
    st = opt.state[params[0]]
    npz("radam.npz", traj=torch.stack(traj), norms=np.asarray(norms), exp_avg_a=st["exp_avg"], exp_avg_sq_a=st["exp_avg_sq"],
        step=np.asarray(st["step"]))


def alignment_maps():
    """Seeded soft attention maps (B, 1, T1, T2) with ragged lengths: softmax over the text axis of smooth logits around a
    diagonal plus noise -- one sharp enough that many probabilities underflow to exactly 0 (log = -inf ties)."""
    B, T1, T2 = 4, 96, 40
    in_lens = torch.tensor([40, 23, 31, 7])
    out_lens = torch.tensor([96, 57, 80, 5])          # the last one has fewer frames than text positions
    t1 = torch.arange(T1, dtype=torch.float32)[None, :, None]
    t2 = torch.arange(T2, dtype=torch.float32)[None, None, :]
    sharp = torch.tensor([0.05, 0.4, 6.0, 0.2])[:, None, None]
    centre = t1 * (in_lens[:, None, None].float() / out_lens[:, None, None].float())
    logits = -sharp * (t2 - centre) ** 2 + 2.0 * syn.hash_uniform("align.noise", (B, T1, T2))
    logprob = logits.unsqueeze(1).contiguous()
    mask = t2 >= in_lens[:, None, None]
    attn = torch.softmax(logits.masked_fill(mask, -float("inf")), dim=2).unsqueeze(1).contiguous()
    return attn, logprob, in_lens, out_lens


def alignment_case():
    """Reference alignment.mas_width1 (numba, alignment.py:31-59) through the loop of binarize_attention
    (tts_lightning_modules.py:270-284), and reference loss.AttentionCTCLoss (loss.py:112-140) forward + gradient."""
    from alignment import mas_width1 as mas
    from loss import AttentionCTCLoss
    attn, logprob, in_lens, out_lens = alignment_maps()
    a = attn.numpy()
    hard = torch.zeros_like(attn)
    for b in range(attn.shape[0]):
        hard[b, 0, :out_lens[b], :in_lens[b]] = torch.tensor(mas(a[b, 0, :out_lens[b], :in_lens[b]]))
    lp = logprob.clone().requires_grad_(True)
    cost, each = AttentionCTCLoss()(lp, in_lens, out_lens, return_all=True)
    cost.backward()
    npz("alignment.npz", attn=attn, logprob=logprob, in_lens=in_lens, out_lens=out_lens, hard=hard, ctc_cost=cost.detach(),
        ctc_each=torch.stack([c.detach() for c in each]), ctc_grad=lp.grad)


def _seeded_state(module, tag, scale=0.3):
    """Overwrite every PARAMETER with a hashed tensor (buffers -- spectral-norm u / v -- keep their constructed values)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            base = syn.hash_uniform(f"{tag}.{name}", tuple(p.shape), -1, 1) * scale
            if name.endswith("weight_g") or name.endswith(".1.weight"):
                base = base.abs() + 0.5
            p.copy_(base)
    # converge the spectral-norm power iteration (u, v buffers) the way training would: a freshly constructed module has
    # random u / v, i.e. an arbitrary "sigma" and recurrent weights of norm >> 1 -- a chaotic LSTM no fixture can pin
    module.train()
    for m in module.modules():
        if isinstance(m, torch.nn.LSTM):
            for _ in range(50):
                for hook in m._forward_pre_hooks.values():
                    hook(m, None)
    module.eval()
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def encoder_case():
    """Reference common.Encoder (common.py:423-500, spectral-normed LSTM as configs/RADMMM_model_config.yaml:15) and
    common.ConvLSTMLinear (common.py:240-330) in eval mode on a ragged batch -- the per-utterance loops are the reference's
    path for B > 1 -- with input and parameter gradients."""
    import common as C
    torch.manual_seed(0)
    lens = torch.tensor([29, 11, 20])
    out = {}
    enc = C.Encoder(3, 64, 5, lstm_norm_fn="spectral").eval()
    sd = _seeded_state(enc, "enc")
    x = syn.hash_uniform("enc.x", (3, 64, 29)).requires_grad_(True)
    y = enc(x, lens)
    g = syn.hash_uniform("enc.g", tuple(y.shape))
    (y * g).sum().backward()
    out.update({"enc_sd." + k: v for k, v in sd.items()})
    out.update(enc_x=x.detach(), enc_y=y.detach(), enc_g=g, enc_dx=x.grad,
               enc_dv0=enc.convolutions[0][0].conv.weight_v.grad, enc_dgamma2=enc.convolutions[2][1].weight.grad,
               enc_dwhh=enc.lstm.weight_hh_l0_orig.grad, enc_dwih_r=enc.lstm.weight_ih_l0_reverse.grad)
    cll = C.ConvLSTMLinear(in_dim=24, out_dim=2, n_layers=2, n_channels=32, kernel_size=3, p_dropout=0.1).eval()
    sd2 = _seeded_state(cll, "cll")
    c = syn.hash_uniform("cll.x", (3, 24, 29)).requires_grad_(True)
    z = cll(c, C.SequenceLength(lens))
    g2 = syn.hash_uniform("cll.g", tuple(z.shape))
    (z * g2).sum().backward()
    out.update({"cll_sd." + k: v for k, v in sd2.items()})
    out.update(cll_x=c.detach(), cll_y=z.detach(), cll_g=g2, cll_dx=c.grad, cll_dv1=cll.convolutions[1].conv.weight_v.grad,
               cll_dwhh_r=cll.bilstm.weight_hh_l0_reverse_orig.grad, cll_ddense=cll.dense.weight.grad, lens=lens)
    import attribute_predictors as AP
    dap = AP.ConvLSTMLinearDAP(n_speaker_dim=4, in_dim=64, out_dim=1, reduction_factor=4, n_backbone_layers=2, n_hidden=32,
                               kernel_size=3, p_dropout=0.1, log_target=True).eval()
    sd3 = _seeded_state(dap, "dap")
    te = syn.hash_uniform("dap.txt", (3, 64, 29)).requires_grad_(True)
    spk = syn.hash_uniform("dap.spk", (3, 4))
    tgt = syn.hash_uniform("dap.tgt", (3, 1, 29), 0.0, 4.0)
    res = dap(tgt, te, spk, C.SequenceLength(lens))
    g3 = syn.hash_uniform("dap.g", tuple(res["x_hat"].shape))
    (res["x_hat"] * g3).sum().backward()
    out.update({"dap_sd." + k: v for k, v in sd3.items()})
    out.update(dap_txt=te.detach(), dap_spk=spk, dap_tgt=tgt, dap_xhat=res["x_hat"].detach(), dap_x=res["x"].detach(), dap_g=g3,
               dap_dtxt=te.grad, dap_dbott=dap.bottleneck_layer.projection_fn.conv.weight_v.grad)
    npz("encoder.npz", **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["keys", "ops", "spline", "frontend", "small", "full", "radam", "alignment", "encoder"]
    if "keys" in which:
        state_dict_keys()
    if "ops" in which:
        op_cases()
    if "spline" in which:
        spline_step_case()
    if "frontend" in which:
        frontend_case()
    if "small" in which:
        decoder_case("decoder_small.npz", n_flows=2, batch=2, frames=128)
    if "full" in which:
        decoder_case("decoder_full.npz", n_flows=8, batch=2, frames=96)
    if "radam" in which:
        radam_case()
    if "alignment" in which:
        alignment_case()
    if "encoder" in which:
        encoder_case()
