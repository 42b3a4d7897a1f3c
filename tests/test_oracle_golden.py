"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz, made by make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import flow as of
from oracle import frontend as ofe
from oracle import spline as osp
from radmmm_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def g(name):
    return {k: torch.from_numpy(v) if v.dtype.kind in "fiub" else v for k, v in np.load(os.path.join(GOLD, name)).items()}


def close(a, b, atol, rtol=0.0):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * b.abs().max().item(), f"max-abs err {err:.3e}"


LENS = torch.tensor([37, 20, 5])


def _wn_small_sd():
    shapes = {"start.bias": (64,), "start.weight_g": (64, 1, 1), "start.weight_v": (64, 16, 1),
              "end.weight": (12, 64, 1), "end.bias": (12,)}
    for i in range(4):
        shapes[f"in_layers.{i}.conv.bias"] = (64,)
        shapes[f"in_layers.{i}.conv.weight_g"] = (64, 1, 1)
        shapes[f"in_layers.{i}.conv.weight_v"] = (64, 64, 5)
        shapes[f"res_skip_layers.{i}.bias"] = (64,)
        shapes[f"res_skip_layers.{i}.weight_g"] = (64, 1, 1)
        shapes[f"res_skip_layers.{i}.weight_v"] = (64, 64, 1)
    return {k: syn.hash_uniform("wn_small." + k, s, -0.3, 0.3) for k, s in shapes.items()}


def test_wn_small():
    gd = g("ops.npz")
    z0 = syn.hash_uniform("wn_small.z0", (3, 6, 37), -1, 1)
    ctx = syn.hash_uniform("wn_small.ctx", (3, 10, 37), -1, 1)
    y = of.wn_forward(_wn_small_sd(), "", z0, ctx, LENS, 4)
    close(y, gd["wn_y"], 2e-5)


def test_partial_conv():
    gd = g("ops.npz")
    sd = {k: syn.hash_uniform("pc." + k, s, -0.5, 0.5) for k, s in
          {"conv.bias": (12,), "conv.weight_g": (12, 1, 1), "conv.weight_v": (12, 8, 5)}.items()}
    x = syn.hash_uniform("pc.x", (3, 8, 37), -1, 1)
    w = of.weight_norm_weight(sd["conv.weight_g"], sd["conv.weight_v"])
    close(of.partial_conv1d(x, w, sd["conv.bias"], LENS, 4), gd["pc_y"], 2e-6)


@pytest.mark.parametrize("fn", ["tanh", "exp", "sigmoid"])
def test_affine_coupling(fn):
    gd = g("ops.npz")
    shapes = {"start.bias": (32,), "start.weight_g": (32, 1, 1), "start.weight_v": (32, 16, 1),
              "end.weight": (12, 32, 1), "end.bias": (12,)}
    for i in range(2):
        shapes[f"in_layers.{i}.conv.bias"] = (32,)
        shapes[f"in_layers.{i}.conv.weight_g"] = (32, 1, 1)
        shapes[f"in_layers.{i}.conv.weight_v"] = (32, 32, 5)
        shapes[f"res_skip_layers.{i}.bias"] = (32,)
        shapes[f"res_skip_layers.{i}.weight_g"] = (32, 1, 1)
        shapes[f"res_skip_layers.{i}.weight_v"] = (32, 32, 1)
    sd = {"affine_param_predictor." + k: syn.hash_uniform("aff.affine_param_predictor." + k, s, -0.3, 0.3)
          for k, s in shapes.items()}
    z = syn.hash_uniform("aff.z", (3, 12, 37), -1, 1)
    ctx = syn.hash_uniform("wn_small.ctx", (3, 10, 37), -1, 1)
    zo, ls = of.affine_coupling(sd, "", z, ctx, LENS, 2, fn)
    close(zo, gd[fn + "_z"], 1e-5)
    close(ls, gd[fn + "_log_s"], 1e-5)
    close(of.affine_coupling(sd, "", zo, ctx, LENS, 2, fn, inverse=True), gd[fn + "_inv"], 1e-4)


def test_invertible_convs():
    gd = g("ops.npz")
    sd = syn.synthetic_state_dict(n_flows=2, n_mel_channels=6, n_group_size=2, tag="inv12")
    z = syn.hash_uniform("inv.z", (3, 12, 37), -2, 2)
    for tag, pre, mode in (("lus", "flows.1.invtbl_conv.", "LUS"), ("whiten", "flows.0.invtbl_conv.", "whiten")):
        zo, ld = of.inv1x1_forward(sd, pre, z, mode)
        close(zo, gd[tag + "_z"], 1e-5)
        close(ld, gd[tag + "_logdet"], 1e-6)
        close(of.inv1x1_inverse(sd, pre, zo, mode), gd[tag + "_inv"], 2e-5)
        # analytic log-det == slogdet of the assembled matrix
        w = of.lus_weight(sd, pre) if mode == "LUS" else of.whiten_weight(sd, pre)
        close(torch.linalg.slogdet(w.double())[1], ld, 1e-5)
    zdata = syn.hash_uniform("inv.init", (3, 12, 37), -2, 2) * syn.hash_uniform("inv.scale", (1, 12, 1), 0.2, 2.0)
    mean, upper, diag = of.whitening_init(zdata, LENS)
    close(mean, gd["init_mean"], 1e-6)
    close(upper, gd["init_upper"], 2e-4, 1e-4)
    close(diag, gd["init_diag"], 2e-4, 1e-4)


def test_splines():
    gd = g("ops.npz")
    x = syn.hash_uniform("spl.x", (50, 7), -0.2, 1.2)
    wt = syn.hash_uniform("spl.w", (50, 7, 32), -2, 2)
    vt = syn.hash_uniform("spl.v", (50, 7, 33), -2, 2)
    y, lj = osp.quadratic_spline(x, wt, vt)
    close(y, gd["spl_yq"], 1e-6)
    close(lj, gd["spl_lj"], 1e-5)
    xi, _ = osp.quadratic_spline(gd["spl_yq"], wt, vt, inverse=True)
    close(xi, gd["spl_xq"], 2e-6)
    yl, ljl = osp.linear_spline(x.clamp(0, 1), wt)
    close(yl, gd["spl_yl"], 1e-6)
    close(ljl, gd["spl_ljl"], 1e-5)
    xli, ljli = osp.linear_spline_inverse(gd["spl_yl"], wt)
    close(xli, gd["spl_xli"], 1e-6)
    close(ljli, gd["spl_ljli"], 1e-5)


def test_masked_bn():
    gd = g("ops.npz")
    x = syn.hash_uniform("pc.x", (3, 8, 37), -1, 1)
    mask = of.length_mask(LENS, 37)[:, None].float()
    sd = {"weight": syn.hash_uniform("bn.w", (8,), 0.5, 1.5), "bias": syn.hash_uniform("bn.b", (8,), -0.5, 0.5),
          "running_mean": torch.zeros(8), "running_var": torch.ones(8), "num_batches_tracked": torch.tensor(0)}
    close(osp.masked_batchnorm(sd, "", x, mask, True, update=True), gd["bn_y"], 1e-5)
    close(sd["running_mean"], gd["bn_rm"], 1e-6)
    close(sd["running_var"], gd["bn_rv"], 1e-6)
    close(osp.masked_batchnorm(sd, "", x, mask, False), gd["bn_y_eval"], 1e-5)


def test_flow_loss():
    gd = g("ops.npz")
    z = syn.hash_uniform("loss.z", (3, 12, 37), -2, 2)
    lsl = [syn.hash_uniform(f"loss.ls{i}", (3, 6, 37), -1, 1) for i in range(3)]
    ldl = [syn.hash_uniform(f"loss.ld{i}", (), -1, 1) for i in range(3)]
    l, lp = of.flow_loss(z, ldl, lsl, LENS, 0.8)
    close(l, gd["loss"], 1e-6)
    close(lp, gd["loss_prior"], 1e-6)


def test_attention():
    gd = g("ops.npz")
    shapes = {"key_proj.0.conv": (48, 24, 3), "key_proj.2.conv": (80, 48, 1), "query_proj.0.conv": (160, 80, 3),
              "query_proj.2.conv": (80, 160, 1), "query_proj.4.conv": (80, 80, 1)}
    sd = {}
    for k, s in shapes.items():
        sd[k + ".bias"] = syn.hash_uniform("att." + k + ".bias", (s[0],), -0.2, 0.2)
        sd[k + ".weight_g"] = syn.hash_uniform("att." + k + ".weight_g", (s[0], 1, 1), -0.2, 0.2)
        sd[k + ".weight_v"] = syn.hash_uniform("att." + k + ".weight_v", s, -0.2, 0.2)
    q_in = syn.hash_uniform("att.q", (3, 80, 37), -1, 1)
    k_in = syn.hash_uniform("att.k", (3, 24, 11), -1, 1)
    in_lens = torch.tensor([11, 7, 3])
    prior = syn.hash_uniform("att.prior", (3, 37, 11), 0.0, 1.0)
    q, k = ofe.attention_projections(sd, "", q_in, k_in)
    a, alp = ofe.soft_attention(q, k, in_lens, prior)
    close(a, gd["att"], 1e-6)
    close(alp, gd["att_logprob"], 1e-5)
    a2, alp2 = ofe.soft_attention(q, k, in_lens, None)
    close(a2, gd["att_noprior"], 1e-6)
    close(alp2, gd["att_logprob_noprior"], 1e-6)
    txt = syn.hash_uniform("att.txt", (3, 24, 11), -1, 1)
    close(ofe.attend(txt, a), gd["att_ctx"], 1e-5)


def test_spline_flow_step():
    """FlowStep(use_spline=True): LUS 1x1 conv + quadratic-spline coupling, train (batch stats + running-stat
    update) then eval (running stats) then inverse -- decoders.py:51-61,72-80."""
    gd = g("spline_step.npz")
    keys = json.load(open(os.path.join(GOLD, "spline_step_keys.json")))
    sd = {}
    for k, shape in keys:
        name = "flows.0." + k
        if k.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0)
        elif "running_var" in k or "lower_diag" in k:
            sd[name] = torch.ones(shape)
        elif k.endswith("invtbl_conv.p"):
            sd[name] = torch.eye(8)[syn.hash_permutation("splstep.p", 8)]
        else:
            sd[name] = syn.hash_uniform("splstep." + k, tuple(shape), -0.3, 0.3)
    sd["flows.0.invtbl_conv.upper_diag"] = syn.hash_uniform("splstep.ud", (8,), 0.7, 1.3)
    lens = torch.tensor([21, 9])
    cfg = of.DecoderConfig(n_conv_layers_per_step=2)
    z = syn.hash_uniform("splstep.z", (2, 8, 21), -3.5, 3.5)
    ctx = syn.hash_uniform("splstep.ctx", (2, 10, 21), -1, 1)
    zc, ld = of.inv1x1_forward(sd, "flows.0.invtbl_conv.", z, "LUS")
    close(ld, gd["log_det"], 1e-6)
    # compared on valid frames only: beyond each length the FiLM/BN stack amplifies fp32 round-off into O(1e-3)
    # differences in the reference itself (its fp32 and an fp64 evaluation disagree there at the same level).
    m = of.length_mask(lens, 21)[:, None]
    zo, ls = osp.spline_coupling(sd, "flows.0.coupling_tfn.", zc, ctx, lens, cfg, training=True, update_bn=True)
    close(zo * m, gd["z"] * m, 5e-5)
    close(ls * m, gd["log_s"] * m, 5e-4)
    close(sd["flows.0.coupling_tfn.param_predictor.in_layers.0.bn.running_mean"], gd["rm0"], 1e-6)
    zo, ls = osp.spline_coupling(sd, "flows.0.coupling_tfn.", zc, ctx, lens, cfg, training=False)
    close(zo * m, gd["z_eval"] * m, 5e-5)
    close(ls * m, gd["log_s_eval"] * m, 5e-4)
    zi = osp.spline_coupling(sd, "flows.0.coupling_tfn.", gd["z_eval"], ctx, lens, cfg, inverse=True, training=False)
    zi = of.inv1x1_inverse(sd, "flows.0.invtbl_conv.", zi, "LUS")
    close(zi * m, gd["z_inv"] * m, 5e-4)


@pytest.mark.parametrize("tag,sr", [("22k", 22050), ("16k", 16000)])
def test_frontend(tag, sr):
    gd = g(f"frontend_{tag}.npz")
    n = 256 * 24
    t = torch.arange(n) / sr
    y = 0.5 * syn.hash_uniform("audio" + tag, (2, n), -1, 1)
    for f, a in ((220.0, 0.3), (1333.0, 0.2), (5200.0, 0.1)):
        y = y + a * torch.sin(2 * math.pi * f * t)[None]
    y = y.clamp(-1, 1)
    mag = ofe.stft_magnitude(y)
    close(mag[:, ::8], gd["mag"], 2e-4)
    close(ofe.mel_spectrogram(y, sr=sr), gd["mel"], 1e-4)
    # fp64 FFT truth agrees with the dense fp32 DFT to fp32 round-off
    close(ofe.stft_magnitude(y, dense=False)[:, ::8], gd["mag"], 1e-3)
    # restated librosa filterbank vs torchaudio's Slaney filterbank (independent implementation)
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(513, 0.0, 8000.0, 80, sr, norm="slaney", mel_scale="slaney").T
    close(torch.from_numpy(ofe.slaney_mel_basis(sr, 1024, 80, 0.0, 8000.0)), fb, 1e-6)


def _decoder_case(fname, n_flows, batch, frames):
    gd = g(fname)
    cfg = of.DecoderConfig.radmmm()
    cfg.n_flows = n_flows
    sd = syn.synthetic_state_dict(n_flows=n_flows)
    bt = syn.synthetic_batch(batch, frames, tag=fname)
    return gd, cfg, sd, bt


@pytest.mark.parametrize("fname,n_flows,batch,frames", [("decoder_small.npz", 2, 2, 128), ("decoder_full.npz", 8, 2, 96)])
def test_decoder_forward_loss_grads_inverse(fname, n_flows, batch, frames):
    gd, cfg, sd, bt = _decoder_case(fname, n_flows, batch, frames)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and
              not any(s in k for s in ("invtbl_conv.p", "lower_diag", "input_mean"))}
    sdp = dict(sd)
    sdp.update(params)
    lstm = of.build_context_lstm(sd, cfg)
    out = of.decoder_forward(sdp, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                             bt["energy_avg"], bt["accent_vecs"], lstm=lstm)
    close(out["z_mel"], gd["z_mel"], 5e-5)
    for i, ls in enumerate(out["log_s_list"]):
        close(ls, gd[f"log_s_{i}"], 2e-5)
    close(torch.stack(out["log_det_W_list"]), gd["log_det"], 1e-5)
    close(out["context_w_spkvec"][:, ::33], gd["context"], 1e-5)
    lens_g = bt["out_lens"] // 2
    n_el = torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor")          # RADMMMLoss.forward, loss.py:520
    loss, prior = of.flow_loss(out["z_mel"], out["log_det_W_list"], out["log_s_list"], lens_g, n_elements=n_el)
    close(loss, gd["loss"], 2e-6)
    close(prior, gd["loss_prior"], 2e-6)
    loss.backward()
    names = [str(n) for n in gd["grad_names"]]
    sums = gd["grad_sums"]
    lstm_grads = {"context_lstm." + n: p.grad for n, p in lstm.named_parameters()}
    for n, row in zip(names, sums):
        gr = lstm_grads[n] if n.startswith("context_lstm.") else params[n].grad
        gr = gr.double().flatten()
        probe = syn.hash_uniform("probe." + n, (gr.numel(),)).double()
        got = torch.tensor([gr.sum(), gr.abs().sum(), (gr * probe).sum()])
        scale = row[1].abs().item() + 1e-12
        assert (got - row).abs().max().item() <= 2e-4 * scale + 1e-9, (n, got, row)
    residual = syn.hash_uniform("residual" + fname, (batch, 160, frames // 2), -1.5, 1.5)
    with torch.no_grad():
        mel = of.decoder_inverse(sd, cfg, residual, out["context_w_spkvec"].detach(), lens_g)
    close(mel, gd["mel_inv"], 2e-4, 1e-5)


def _radam_inputs():
    shapes = {"a": (37, 19), "b": (1024,), "c": (5, 7, 3), "d": (1,)}
    params = [syn.hash_uniform("radam.p." + k, s, -1, 1) for k, s in shapes.items()]
    grads = [[syn.hash_uniform(f"radam.g{step}." + k, s, -1, 1) * (3.0 if step % 2 == 0 else 0.01) for k, s in shapes.items()]
             for step in range(9)]
    return shapes, params, grads


def test_radam_trajectory():
    """oracle.flow.radam_step + clip_grad_norm against 9 steps of the reference radam.RAdam under Lightning-style clipping
    (radam.py:63-142; the trajectory crosses the N_sma >= 5 switch between steps 5 and 6)."""
    gd = g("radam.npz")
    _, params, grads = _radam_inputs()
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for step in range(9):
        gs = [x.clone() for x in grads[step]]
        total = of.clip_grad_norm(gs, 1.0)
        assert abs(float(total) - float(gd["norms"][step])) <= 1e-5 * float(gd["norms"][step])
        of.radam_step(params, gs, m, v, step + 1, lr=1e-3, weight_decay=1e-6)
        close(torch.cat([p.flatten() for p in params]), gd["traj"][step], 1e-7)
    close(m[0], gd["exp_avg_a"], 1e-8)
    close(v[0], gd["exp_avg_sq_a"], 1e-9)
    assert int(gd["step"]) == 9


def test_alignment_oracle_vs_reference():
    """oracle.alignment against outputs of the unmodified reference (numba alignment.mas_width1 looped as
    TTSModel.binarize_attention does; loss.AttentionCTCLoss forward + gradient) on ragged maps that include exact-zero
    probabilities (-inf ties), fewer frames than text positions (the extra opt[0, 0] = 1) and an infeasible CTC target."""
    from oracle import alignment as oa
    gd = g("alignment.npz")
    hard = oa.binarize_attention(gd["attn"], gd["in_lens"], gd["out_lens"])
    assert torch.equal(hard, gd["hard"])
    assert (gd["attn"] == 0).sum() > 1000 and gd["hard"][3, 0, 0].sum() == 2
    lp = gd["logprob"].clone().requires_grad_(True)
    cost, each = oa.attention_ctc_loss(lp, gd["in_lens"], gd["out_lens"])
    cost.backward()
    close(cost.detach(), gd["ctc_cost"], 1e-6)
    close(torch.stack([e.detach() for e in each]), gd["ctc_each"], 1e-6)
    assert float(gd["ctc_each"][3]) == 0.0
    close(lp.grad, gd["ctc_grad"], 1e-7)


def _sub(gd, prefix):
    return {k[len(prefix):]: v for k, v in gd.items() if k.startswith(prefix)}


def test_encoder_oracle_vs_reference():
    """oracle.encoders (per-utterance loops) against the unmodified reference Encoder / ConvLSTMLinear outputs."""
    from oracle import encoders as oe
    gd = g("encoder.npz")
    with torch.no_grad():
        y = oe.encoder_forward(_sub(gd, "enc_sd."), gd["enc_x"], gd["lens"])
        z = oe.conv_lstm_linear_forward(_sub(gd, "cll_sd."), gd["cll_x"], gd["lens"])
    close(y, gd["enc_y"], 2e-6)
    close(z, gd["cll_y"], 2e-6)


def test_batched_encoder_convs_equal_per_utterance_loop():
    """The product's batched conv banks (masked input, closed-form partial-conv ratio, masked instance norm) are plain torch
    ops, so their equality with the per-utterance loop is checked here on the CPU; the LSTM half needs the GPU
    (tests/test_gpu_parity_full.py)."""
    from oracle import encoders as oe
    from radmmm_b200 import encoders as pe
    gd = g("encoder.npz")
    enc = pe.Encoder(3, 64, 5, lstm_norm_fn="spectral").eval()
    enc.load_state_dict(_sub(gd, "enc_sd."), strict=True)
    lens = gd["lens"]
    with torch.no_grad():
        got = enc._convs(gd["enc_x"], lens.long())
        ref = oe.encoder_convs(_sub(gd, "enc_sd."), gd["enc_x"], lens)
    for b, r in enumerate(ref):
        close(got[b, :, :r.shape[0]].t(), r, 2e-5)
        assert float(got[b, :, r.shape[0]:].abs().sum()) == 0.0
    cll = pe.ConvLSTMLinear(in_dim=24, out_dim=2, n_layers=2, n_channels=32, kernel_size=3, p_dropout=0.1).eval()
    cll.load_state_dict(_sub(gd, "cll_sd."), strict=True)
