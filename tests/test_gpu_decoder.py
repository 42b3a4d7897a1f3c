"""GPU parity of the fused flow step and the full decoder against the oracle and the reference-made fixtures."""
import os

import pytest
import torch

from oracle import flow as of
from radmmm_b200 import _native as N
from radmmm_b200 import synthetic as syn
from tests.gpu_util import DEV, close, err, gold

pytestmark = pytest.mark.gpu

# tolerance of each contraction mode relative to the fp32 reference (stated in DESIGN.md "precision modes")
Z_TOL = {"fp32": 5e-5, "bf16x3": 5e-4, "bf16": 0.25}
GRAD_TOL = {"fp32": 3e-4, "bf16x3": 2e-3, "bf16": 6e-2}


def _small_layer(fn="tanh", H=128, C=12, D=10, L=3):
    from radmmm_b200 import common
    layer = common.AffineTransformationLayer(C, D, L, affine_model="wavenet", scaling_fn=fn, n_channels=H,
                                             use_partial_padding=True)
    sd = {}
    for k, v in layer.state_dict().items():
        lo, hi = (-0.3, 0.3)
        if k.endswith("weight_g"):
            lo, hi = 0.5, 1.5
        if k.startswith("affine_param_predictor.end."):      # keep tanh(a) away from saturation (fp32 cancellation)
            lo, hi = -0.02, 0.02
        sd[k] = syn.hash_uniform("small." + k, tuple(v.shape), lo, hi)
    layer.load_state_dict(sd)
    return layer, sd


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("lens", [[37, 20, 5], [37, 37, 37], [1, 2, 3]])
def test_affine_layer_small(precision, lens):
    """AffineTransformationLayer (WN width 128) forward / inverse / all gradients vs oracle autograd; ragged lengths
    including sequences shorter than the dilated receptive field."""
    layer, sd = _small_layer()
    layer.precision = precision
    layer = layer.to(DEV)
    B, C, T, D = 3, 12, 37, 10
    lens_t = torch.tensor(lens)
    z = syn.hash_uniform("sl.z", (B, C, T), -1.5, 1.5)
    ctx = syn.hash_uniform("sl.ctx", (B, D, T), -1, 1)
    mask = of.length_mask(lens_t, T)[:, None].float()
    from radmmm_b200.common import SequenceLength
    zg = z.to(DEV).requires_grad_(True)
    cg = (ctx * mask).to(DEV).requires_grad_(True)
    zo, ls = layer(zg, cg, seq_lens=SequenceLength(lens_t.to(DEV), T))
    # oracle (fp64)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    zc, cc = z.double().requires_grad_(True), (ctx * mask).double().requires_grad_(True)
    zo_ref, ls_ref = of.affine_coupling(sdd, "", zc, cc, lens_t, 3, "tanh")
    tol = Z_TOL[precision]
    m = mask.double()
    close(zo.cpu().double() * m, zo_ref.detach() * m, tol * 4, what="z (valid frames)")
    close(ls.cpu().double() * m, ls_ref.detach() * m, tol * 4, what="log_s (valid frames)")
    if precision == "fp32":      # beyond the lengths the reference computes defined garbage; the fp32 path reproduces it
        close(zo, zo_ref.detach(), 1e-4, what="z (all frames)")
        close(ls, ls_ref.detach(), 1e-4, what="log_s (all frames)")
    # inverse
    zi = layer(zo.detach(), cg.detach(), inverse=True, seq_lens=SequenceLength(lens_t.to(DEV), T))
    close(zi.cpu().double() * m, z.double() * m, tol * 8, what="inverse round trip")
    # gradients (incoming gradients masked, as the flow loss does)
    g1 = syn.hash_uniform("sl.g1", (B, C, T)) * mask
    g2 = syn.hash_uniform("sl.g2", (B, C // 2, T)) * mask
    ((zo * g1.to(DEV)).sum() + (ls * g2.to(DEV)).sum()).backward()
    ((zo_ref * g1.double()).sum() + (ls_ref * g2.double()).sum()).backward()
    gt = GRAD_TOL[precision]
    close(zg.grad, zc.grad, gt * max(1.0, zc.grad.abs().max().item()), what="dz")
    close(cg.grad.cpu().double() * m, cc.grad * m, gt * max(1.0, cc.grad.abs().max().item()), what="dcontext")
    for name, p in layer.named_parameters():
        ref = sdd[name].grad
        close(p.grad, ref, gt * max(1e-3, ref.abs().max().item()), what="grad " + name)


def _decoder(n_flows, precision):
    from radmmm_b200 import decoders
    dec = decoders.RADMMMFlow(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2,
                              n_mel_channels=80, n_flows=n_flows)
    dec.load_state_dict(syn.synthetic_state_dict(n_flows=n_flows))
    return dec.to(DEV).set_precision(precision).train()


@pytest.mark.parametrize("fname,n_flows,batch,frames", [("decoder_small.npz", 2, 2, 128), ("decoder_full.npz", 8, 2, 96)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_decoder_vs_reference_fixture(fname, n_flows, batch, frames, precision):
    """RADMMMFlow.forward + flow loss + backward + inverse against outputs of the UNMODIFIED reference."""
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    gd = gold(fname)
    dec = _decoder(n_flows, precision)
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(batch, frames, tag=fname).items()}
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    lens_g = (bt["out_lens"] // 2).cpu()
    m = of.length_mask(lens_g, frames // 2)[:, None].double()
    tol = Z_TOL[precision]
    assert out["z_mel"].shape == (batch, 160, frames // 2)
    close(out["z_mel"].cpu().double() * m, gd["z_mel"].double() * m, tol, what="z_mel")
    for i, ls in enumerate(out["log_s_list"]):
        close(ls.cpu().double() * m, gd[f"log_s_{i}"].double() * m, tol, what=f"log_s[{i}]")
    close(torch.stack(out["log_det_W_list"]), gd["log_det"], 1e-5, what="log_det_W")
    # "bf16": the context LSTM runs its recurrence with bf16 W_hh / h on the tensor cores (csrc/lstm_cluster.cu)
    close(out["context_w_spkvec"][:, ::33], gd["context"], 5e-3 if precision == "bf16" else 1e-5, what="context_w_spkvec")
    loss, prior = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], lens_g.to(DEV),
                             n_elements=L.n_elements_like_reference(bt["out_lens"], 2))
    rel = {"fp32": 1e-5, "bf16x3": 5e-5, "bf16": 5e-3}[precision]          # north-star bar: 1e-4 relative
    close(loss, gd["loss"], rel * abs(float(gd["loss"])), what="loss")          # log-det bar: 1e-4 relative
    close(prior, gd["loss_prior"], rel * abs(float(gd["loss_prior"])) * 10, what="loss_prior")
    loss.backward()
    names = [str(n) for n in gd["grad_names"]]
    params = dict(dec.named_parameters())
    gt = GRAD_TOL[precision]
    worst = 0.0
    for n, row in zip(names, gd["grad_sums"]):
        g = params[n].grad.double().flatten().cpu()
        probe = syn.hash_uniform("probe." + n, (g.numel(),)).double()
        got = torch.tensor([g.sum(), g.abs().sum(), (g * probe).sum()])
        scale = abs(float(row[1])) + 1e-12
        e = (got - row).abs().max().item() / scale
        worst = max(worst, e)
        assert e <= gt, f"gradient checksum of {n}: rel err {e:.3e} (got {got.tolist()}, ref {row.tolist()})"
    # inverse with the reference's injected residual
    dec.eval()
    residual = syn.hash_uniform("residual" + fname, (batch, 160, frames // 2), -1.5, 1.5).to(DEV)
    dur = torch.zeros(batch, 4, dtype=torch.long, device=DEV)
    # infer() builds its context from (txt_enc, dur); replay the flow loop on the training context instead, as the
    # golden generator does (decoders.py:227-246)
    from radmmm_b200.decoders import _Lens
    with torch.no_grad():
        stack = dec.exit_steps.copy()
        mel = residual[:, len(stack) * 2:]
        rest = residual[:, :len(stack) * 2]
        ctx = out["context_w_spkvec"].detach()
        for i, fs in enumerate(reversed(dec.flows)):
            cur = len(dec.flows) - i - 1
            mel = fs(mel, ctx, inverse=True, seq_lens=_Lens(lens_g.to(DEV)))
            if stack and cur == stack[-1]:
                stack.pop()
                mel = torch.cat((rest[:, len(stack) * 2:], mel), 1)
                rest = rest[:, :len(stack) * 2]
        mel = dec.fold(mel)
    mm = of.length_mask(bt["out_lens"].cpu(), frames)[:, None].double()
    close(mel.cpu().double() * mm, gd["mel_inv"].double() * mm, 1e-3 if precision != "bf16" else 0.5, what="inverse mel")


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_decoder_roundtrip_full_size(precision):
    """Size-independent property at the benchmark's shape (B=8, T=800): infer-path inverse of the forward output
    recovers the mel; log-determinant terms are finite; z beyond the lengths never leaks into valid frames."""
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.decoders import _Lens
    dec = _decoder(8, precision).eval()
    B, T = 8, 800
    bt = {k: v.to(DEV) for k, v in syn.synthetic_batch(B, T, tag="rt").items()}
    with torch.no_grad():
        out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
                  energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        lens_g = bt["out_lens"] // 2
        z = out["z_mel"]
        assert torch.isfinite(z).all() and all(torch.isfinite(ls).all() for ls in out["log_s_list"])
        stack = dec.exit_steps.copy()
        mel = z[:, len(stack) * 2:]
        rest = z[:, :len(stack) * 2]
        for i, fs in enumerate(reversed(dec.flows)):
            cur = len(dec.flows) - i - 1
            mel = fs(mel, out["context_w_spkvec"], inverse=True, seq_lens=_Lens(lens_g))
            if stack and cur == stack[-1]:
                stack.pop()
                mel = torch.cat((rest[:, len(stack) * 2:], mel), 1)
                rest = rest[:, :len(stack) * 2]
        mel = dec.fold(mel)
        mm = of.length_mask(bt["out_lens"].cpu(), T)[:, None].double()
        tol = 2e-3 if precision == "bf16x3" else 0.5
        close(mel.cpu().double() * mm, bt["mel"].cpu().double() * mm, tol, what="forward->inverse round trip")
        # changing only padded frames of the input must not change valid frames of the output
        mel2 = bt["mel"].clone()
        pad = ~of.length_mask(bt["out_lens"].cpu(), T)[:, None].to(DEV)
        mel2 = torch.where(pad.expand_as(mel2), torch.full_like(mel2, 3.0), mel2)
        out2 = dec(mel2, bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], T), f0=bt["f0"],
                   energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        mg = of.length_mask(lens_g.cpu(), T // 2)[:, None].to(DEV)
        assert torch.equal(out2["z_mel"] * mg, z * mg)


def test_infer_api():
    """RADMMMFlow.infer: shapes, sigma=0 determinism, residual injection (decoders.py:207-248)."""
    dec = _decoder(2, "bf16x3").eval()
    B, T2 = 2, 9
    dur = torch.tensor([[3, 2, 4, 1, 2, 5, 3, 2, 2], [2, 2, 2, 2, 2, 2, 2, 2, 0]], device=DEV)
    out_lens = dur.sum(1)
    T = int(out_lens.max())
    txt = syn.hash_uniform("inf.txt", (B, 520, T2)).to(DEV)
    spk = syn.hash_uniform("inf.spk", (B, 16)).to(DEV)
    f0 = syn.hash_uniform("inf.f0", (B, T), 4.4, 6.4).to(DEV)
    en = syn.hash_uniform("inf.en", (B, T), 0.5, 1.0).to(DEV)
    with torch.no_grad():
        a = dec.infer(spk, txt, 0.0, dur=dur, f0=f0, energy_avg=en, out_lens=out_lens)["mel"]
        b = dec.infer(spk, txt, 0.0, dur=dur, f0=f0, energy_avg=en, out_lens=out_lens)["mel"]
        c = dec.infer(spk, txt, 0.7, dur=dur, f0=f0, energy_avg=en, out_lens=out_lens)["mel"]
    assert a.shape == (B, 80, T // 2 * 2)
    assert torch.equal(a, b) and torch.isfinite(c).all() and not torch.equal(a, c)


_LSTM_CASES = [(3, [37, 20, 5], 20, 16), (2, [9, 9], 20, 16), (11, [40, 33, 31, 30, 25, 17, 16, 9, 3, 2, 1], 20, 16),
               # the decoder's real width (hidden 528 = 16 CTAs x 33 units, no padding) and a width that leaves the last CTAs
               # of the cluster partly / completely empty (RADTTS variant: 524)
               (8, [24, 24, 21, 17, 12, 9, 4, 1], 40, 528), (3, [11, 7, 2], 24, 524),
               # more than one 8-sequence tile per launch, more than one launch (chunks of 32 / 64 sequences)
               (19, [13 - (i % 13) for i in range(19)], 20, 40), (70, [1 + (7 * i) % 12 for i in range(70)], 12, 16),
               # hidden <= 132 runs on clusters of 4 CTAs in bf16 mode (the attribute predictors' hidden 128; 132 fills all
               # four CTAs; 100 leaves the last one a third full); 133 is the first size back on clusters of 16
               (8, [30, 30, 26, 19, 12, 8, 3, 1], 24, 128), (5, [17, 9, 9, 4, 2], 20, 132), (11, [21 - 2 * i for i in range(11)], 20, 100),
               (3, [12, 7, 5], 20, 133)]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("batch,lens,n_in,hid", _LSTM_CASES)
def test_context_lstm_vs_torch(precision, batch, lens, n_in, hid):
    """Context bi-LSTM kernels (forward + all gradients) vs torch's packed nn.LSTM in fp64 on the CPU: the fp32 cooperative
    recurrence ("fp32", "bf16x3") and the cluster-resident tensor-core recurrence ("bf16": bf16 W_hh / h, fp32 state)."""
    from radmmm_b200.lstm import context_lstm
    T = max(lens)
    ref = torch.nn.LSTM(n_in, hid, 1, batch_first=True, bidirectional=True)
    wmax = min(0.4, 1.5 / hid ** 0.5)                  # nn.LSTM initialises with U(-1/sqrt(H), 1/sqrt(H))
    for n, p in ref.named_parameters():
        p.data.copy_(syn.hash_uniform("lstm." + n, tuple(p.shape), -wmax, wmax))
    x = syn.hash_uniform("lstm.x", (batch, T, n_in), -1, 1)
    lens_t = torch.tensor(lens)
    mask = of.length_mask(lens_t, T)[..., None].float()
    gout = syn.hash_uniform("lstm.g", (batch, T, 2 * hid)) * mask
    ref64 = torch.nn.LSTM(n_in, hid, 1, batch_first=True, bidirectional=True).double()
    ref64.load_state_dict(ref.state_dict())
    xc = (x * mask).double().requires_grad_(True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xc, lens_t, batch_first=True, enforce_sorted=False)
    yo, _ = ref64(packed)
    yo, _ = torch.nn.utils.rnn.pad_packed_sequence(yo, batch_first=True, total_length=T)
    (yo * gout.double()).sum().backward()
    mine = torch.nn.LSTM(n_in, hid, 1, batch_first=True, bidirectional=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    xg = (x * mask).to(DEV).requires_grad_(True)
    y = context_lstm(mine, xg, lens_t.to(DEV), precision)
    tol = {"fp32": 2e-5, "bf16x3": 2e-4, "bf16": 1.5e-2}[precision]
    close(y, yo.detach(), tol, what="lstm out")
    (y * gout.to(DEV)).sum().backward()
    close(xg.grad.cpu().double() * mask.double(), xc.grad * mask.double(), tol * 5, what="lstm dx")
    for (n, p), (_, q) in zip(mine.named_parameters(), ref64.named_parameters()):
        close(p.grad, q.grad, tol * 10 * max(1.0, q.grad.abs().max().item()), what="lstm grad " + n)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_spline_flow_step_vs_reference(precision):
    """FlowStep(use_spline=True): LUS 1x1 conv + FiLM parameter net + quadratic-spline kernels against the reference
    fixture (train-mode batch statistics incl. the running-stat update, eval mode, inverse) and oracle gradients."""
    import json
    from oracle import spline as osp
    from radmmm_b200 import decoders
    from radmmm_b200.common import SequenceLength
    gd = gold("spline_step.npz")
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "spline_step_keys.json")))
    sd = {}
    for k, shape in keys:
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0)
        elif "running_var" in k or "lower_diag" in k:
            sd[k] = torch.ones(shape)
        elif k.endswith("invtbl_conv.p"):
            sd[k] = torch.eye(8)[syn.hash_permutation("splstep.p", 8)]
        else:
            sd[k] = syn.hash_uniform("splstep." + k, tuple(shape), -0.3, 0.3)
    sd["invtbl_conv.upper_diag"] = syn.hash_uniform("splstep.ud", (8,), 0.7, 1.3)
    fs = decoders.FlowStep(8, 10, 2, mode="LUS", use_partial_padding=True, use_spline=True, use_bn=True)
    assert [k for k, _ in keys] == list(fs.state_dict().keys())
    fs.load_state_dict(sd)
    fs.coupling_tfn.precision = precision
    fs = fs.to(DEV).train()
    lens = torch.tensor([21, 9])
    z = syn.hash_uniform("splstep.z", (2, 8, 21), -3.5, 3.5)
    ctx = syn.hash_uniform("splstep.ctx", (2, 10, 21), -1, 1)
    m = of.length_mask(lens, 21)[:, None].double()
    tol = 1e-4 if precision == "fp32" else 2e-3
    zg = z.to(DEV).requires_grad_(True)
    seq = SequenceLength(lens.to(DEV), 21)
    zo, ld, ls = fs(zg, ctx.to(DEV), seq_lens=seq)
    close(ld, gd["log_det"], 1e-6)
    close(zo.cpu().double() * m, gd["z"].double() * m, tol, what="spline step z (train)")
    close(ls.cpu().double() * m, gd["log_s"].double() * m, tol * 10, what="spline step log_s (train)")
    close(fs.coupling_tfn.param_predictor.in_layers[0].bn.running_mean, gd["rm0"], 1e-5, what="running mean")
    # gradients vs oracle autograd (fp64) on the same training forward
    g1 = syn.hash_uniform("splstep.g1", (2, 8, 21)) * m.float()
    g2 = syn.hash_uniform("splstep.g2", (2, 1, 21)) * m.float()
    ((zo * g1.to(DEV)).sum() + (ls * g2.to(DEV)).sum()).backward()
    sdd = {"flows.0." + k: (v.double().requires_grad_(True) if v.dtype == torch.float32 and v.numel() > 1 and
                            not k.endswith(".p") and "running" not in k and "lower_diag" not in k else v.double() if v.is_floating_point() else v)
           for k, v in sd.items()}
    zc = z.double().requires_grad_(True)
    cfg = of.DecoderConfig(n_conv_layers_per_step=2)
    zm, _ = of.inv1x1_forward(sdd, "flows.0.invtbl_conv.", zc, "LUS")
    zr, lr = osp.spline_coupling(sdd, "flows.0.coupling_tfn.", zm, ctx.double(), lens, cfg, training=True)
    ((zr * g1.double()).sum() + (lr * g2.double()).sum()).backward()
    gt = 2e-3 if precision == "fp32" else 2e-2
    close(zg.grad, zc.grad, gt * max(1.0, zc.grad.abs().max().item()), what="spline step dz")
    for name in ("coupling_tfn.param_predictor.end.weight", "coupling_tfn.param_predictor.in_layers.1.hidden_conv.conv.weight_v",
                 "coupling_tfn.param_predictor.in_layers.0.cond_conv.conv.weight_g", "coupling_tfn.param_predictor.in_layers.0.bn.weight",
                 "invtbl_conv.upper"):
        ref = sdd["flows.0." + name].grad
        mine = dict(fs.named_parameters())[name].grad
        close(mine, ref, gt * max(1e-3, ref.abs().max().item()), what="grad " + name)
    fs.eval()
    with torch.no_grad():
        zo2, _, ls2 = fs(z.to(DEV), ctx.to(DEV), seq_lens=seq)
        close(zo2.cpu().double() * m, gd["z_eval"].double() * m, tol, what="spline step z (eval)")
        close(ls2.cpu().double() * m, gd["log_s_eval"].double() * m, tol * 10, what="spline step log_s (eval)")
        zi = fs(gd["z_eval"].to(DEV), ctx.to(DEV), inverse=True, seq_lens=seq)
        close(zi.cpu().double() * m, gd["z_inv"].double() * m, 5e-3, what="spline step inverse")
