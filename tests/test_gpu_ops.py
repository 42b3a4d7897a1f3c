"""GPU parity tests of the individual kernels against the CPU oracle, called through the C ABI (ctypes)."""
import math

import pytest
import torch

from oracle import flow as of
from oracle import frontend as ofe
from oracle import spline as osp
from radmmm_b200 import _native as N
from radmmm_b200 import synthetic as syn
from tests.gpu_util import DEV, act_to_float, cast_rows, close, err, gold

pytestmark = pytest.mark.gpu
LENS = torch.tensor([37, 20, 5])
MODE_TOL = {"fp32": 2e-5, "bf16x3": 2e-4, "bf16": 3e-2}


# ------------------------------------------------------------------------------------------------ contractions
@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("taps,dil,K,Nout,R", [(1, 1, 128, 256, 384), (5, 1, 128, 256, 384), (5, 8, 256, 128, 384),
                                                (1, 1, 192, 160, 384), (5, 4, 64, 1024, 384),
                                                # even tile counts -> 2x2 cluster multicast variant
                                                (5, 2, 256, 1024, 512), (1, 1, 1024, 512, 1024), (5, 8, 128, 256, 256),
                                                # 416 tiles of 128 x 256 on 74 CTA-pair slots: every CTA walks 5-6 tiles
                                                # (TMEM double-buffer phase flips, smem ring wrap-around across tiles)
                                                (1, 1, 1024, 1024, 13312), (5, 2, 1024, 1024, 13312),
                                                # odd row-tile count -> the 1-CTA kernel, several tiles per CTA
                                                (5, 4, 1024, 1024, 13312 + 128)])
def test_conv_rows(precision, taps, dil, K, Nout, R):
    """Row GEMM (the WN conv): y[r] = sum_j W_j x[r + (j-c)d] + bias, zero outside [0,R)."""
    lib = N.lib()
    mode = N.MODES[precision]
    npad = N.round_up(Nout, 128)
    x = syn.hash_uniform(f"cr.x{K}", (R, K)).to(DEV)
    w = torch.zeros(taps, npad, K)
    w[:, :Nout] = syn.hash_uniform(f"cr.w{K}{Nout}{taps}", (taps, Nout, K), -0.1, 0.1)
    w = w.to(DEV)
    bias = syn.hash_uniform("cr.b", (npad,)).to(DEV)
    xb, xld, xpl = cast_rows(x, mode)
    wb, wld, wpl = cast_rows(w.reshape(taps * npad, K), mode)
    y = torch.full((R, npad), float("nan"), device=DEV)
    N.check(lib.radmmm_conv_rows(mode, N.ptr(xb), xld, xpl, N.ptr(wb), wld, wpl, npad * K, N.fptr(bias), N.fptr(y), npad,
                                 R, K, Nout, taps, dil, N.stream()))
    torch.cuda.synchronize()
    big = R * K * Nout * taps > 1 << 31          # the checker is a plain fp64 matmul either way; big cases run it on the GPU
    xd, wd = (x.double(), w.double()) if big else (x.double().cpu(), w.double().cpu())
    ref = torch.zeros(R, npad, dtype=torch.float64, device=xd.device)
    for j in range(taps):
        s = (j - taps // 2) * dil
        xs = torch.zeros_like(xd)
        lo, hi = max(0, -s), min(R, R - s)
        xs[lo:hi] = xd[lo + s:hi + s]
        ref += xs @ wd[j].t()
    ref += bias.double().to(ref.device)[None]
    scale = ref.abs().max().item()
    close(y[:, :Nout], ref[:, :Nout].cpu(), MODE_TOL[precision] * scale, what=f"conv_rows {precision}")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
@pytest.mark.parametrize("taps,dil,M,Nx,R", [(1, 1, 128, 256, 640), (5, 2, 256, 128, 640), (1, 1, 256, 1152, 640),
                                              (5, 8, 128, 128, 640), (5, 1, 1024, 1024, 640),
                                              # the shapes the train step launches (res-skip / dilated-conv weight grads)
                                              (1, 1, 1024, 1024, 3328), (5, 2, 1024, 1024, 3328), (1, 1, 1024, 1024, 13312)])
def test_wgrad_rows(precision, taps, dil, M, Nx, R):
    """Weight-grad GEMM: out[j][m][n] = sum_r dy[r][m] x[r + (j-c)d][n]."""
    lib = N.lib()
    mode = N.MODES[precision]
    dy = syn.hash_uniform(f"wg.dy{M}", (R, M)).to(DEV)
    x = syn.hash_uniform(f"wg.x{Nx}", (R, Nx)).to(DEV)
    dyb, dld, dpl = cast_rows(dy, mode)
    xb, xld, xpl = cast_rows(x, mode)
    out = torch.full((taps, M, Nx), float("nan"), device=DEV)
    N.check(lib.radmmm_wgrad_rows(mode, N.ptr(dyb), dld, dpl, N.ptr(xb), xld, xpl, N.fptr(out), Nx,
                                  M * Nx, R, M, Nx, taps, dil, 0, N.stream()))
    torch.cuda.synchronize()
    big = R * M * Nx * taps > 1 << 31
    dyd, xd = (dy.double(), x.double()) if big else (dy.double().cpu(), x.double().cpu())
    ref = torch.zeros(taps, M, Nx, dtype=torch.float64, device=xd.device)
    for j in range(taps):
        s = (j - taps // 2) * dil
        xs = torch.zeros_like(xd)
        lo, hi = max(0, -s), min(R, R - s)
        xs[lo:hi] = xd[lo + s:hi + s]
        ref[j] = dyd.t() @ xs
    close(out, ref.cpu(), MODE_TOL[precision] * ref.abs().max().item(), what=f"wgrad_rows {precision}")


# ------------------------------------------------------------------------------------------------ invertible convs
def test_invertible_convs_match_oracle_and_golden():
    from radmmm_b200 import common
    gd = gold("ops.npz")
    sd = syn.synthetic_state_dict(n_flows=2, n_mel_channels=6, n_group_size=2, tag="inv12")
    z = syn.hash_uniform("inv.z", (3, 12, 37), -2, 2)
    for tag, pre, cls in (("lus", "flows.1.invtbl_conv.", common.Invertible1x1ConvLUS),
                          ("whiten", "flows.0.invtbl_conv.", common.DataInitializedInvertible1x1Conv)):
        m = cls(12)
        m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        m = m.to(DEV).eval()
        zg = z.to(DEV).requires_grad_(True)
        zo, ld = m(zg, lens=common.SequenceLength(LENS.to(DEV), 37))
        close(zo, gd[tag + "_z"], 1e-5, what=tag)
        close(ld, gd[tag + "_logdet"], 1e-6)
        close(m(zo.detach(), inverse=True), gd[tag + "_inv"], 2e-5)
        # gradients vs oracle autograd (incoming gradient masked like the flow loss does)
        mask = of.length_mask(LENS, 37)[:, None].float()
        gout = syn.hash_uniform("inv.g", (3, 12, 37)) * mask
        (zo * gout.to(DEV)).sum().backward()
        sdp = {k: v.clone().double().requires_grad_(v.dtype == torch.float32 and "input_mean" not in k and ".p" not in k)
               for k, v in sd.items() if k.startswith(pre)}
        zc = z.double().requires_grad_(True)
        zo_ref, _ = of.inv1x1_forward(sdp, pre, zc, "LUS" if tag == "lus" else "whiten")
        (zo_ref * gout.double()).sum().backward()
        close(zg.grad, zc.grad, 2e-5, what="dz")
        for name in ("upper", "upper_diag") + (("lower",) if tag == "lus" else ()):
            g_ref = sdp[pre + name].grad
            close(getattr(m, name).grad, g_ref, 1e-4 * max(1.0, g_ref.abs().max().item()), what=name)


def test_whitening_init_matches_golden():
    from radmmm_b200 import common
    gd = gold("ops.npz")
    w = common.DataInitializedInvertible1x1Conv(12).to(DEV).train()
    zdata = (syn.hash_uniform("inv.init", (3, 12, 37), -2, 2) * syn.hash_uniform("inv.scale", (1, 12, 1), 0.2, 2.0)).to(DEV)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        zo, _ = w(zdata, lens=common.SequenceLength(LENS.to(DEV), 37))
    assert bool(w.initialized)
    close(w.input_mean, gd["init_mean"], 1e-6)
    close(w.upper, gd["init_upper"], 5e-4, 1e-4)
    close(w.upper_diag, gd["init_diag"], 5e-4, 1e-4)
    close(zo, gd["init_z"], 2e-3)


# ------------------------------------------------------------------------------------------------ coupling + loss
@pytest.mark.parametrize("fn", ["tanh", "exp", "sigmoid"])
def test_coupling_kernels(fn):
    lib = N.lib()
    B, C, T = 3, 12, 37
    z = syn.hash_uniform("cp.z", (B, C, T), -2, 2)
    params = syn.hash_uniform("cp.p", (B, C, T), -1.5, 1.5)
    lens = LENS.to(torch.int32)
    zc, pc = z.double().requires_grad_(True), params.double().requires_grad_(True)
    s, log_s = of.scale_and_log(pc[:, :6], fn)
    zo_ref = torch.cat((zc[:, :6], s * zc[:, 6:] + pc[:, 6:]), 1)
    zd, pd = z.to(DEV), params.to(DEV)
    zo, ls = torch.empty_like(zd), torch.empty(B, 6, T, device=DEV)
    N.check(lib.radmmm_coupling_forward(N.fptr(zd), N.fptr(pd), N.fptr(zo), N.fptr(ls), B, C, T, N.SCALING[fn], 0, N.stream()))
    close(zo, zo_ref.detach(), 2e-6, what="z")
    close(ls, log_s.detach(), 2e-6, what="log_s")
    zi = torch.empty_like(zd)
    N.check(lib.radmmm_coupling_forward(N.fptr(zo), N.fptr(pd), N.fptr(zi), None, B, C, T, N.SCALING[fn], 1, N.stream()))
    close(zi, z, 2e-5, what="inverse")
    mask = of.length_mask(LENS, T)[:, None].double()
    g1, g2 = syn.hash_uniform("cp.g1", (B, C, T)).double(), syn.hash_uniform("cp.g2", (B, 6, T)).double()
    ((zo_ref * g1 + 0).mul(mask).sum() + (log_s * g2 * mask).sum()).backward()
    dz, dp = torch.empty_like(zd), torch.empty_like(zd)
    g1d, g2d, lensd = g1.float().to(DEV), g2.float().to(DEV), lens.to(DEV)     # keep alive until the kernel has run
    N.check(lib.radmmm_coupling_backward(N.fptr(g1d), N.fptr(g2d), N.fptr(zd), N.fptr(pd), N.ptr(lensd), N.fptr(dz),
                                         N.fptr(dp), B, C, T, N.SCALING[fn], N.stream()))
    torch.cuda.synchronize()
    ref_dz = zc.grad.clone()
    close(dz, ref_dz, 2e-5, what="dz")
    close(dp, pc.grad, 2e-5 * max(1.0, pc.grad.abs().max().item()), what="dparams")


def test_flow_loss_matches_oracle_and_golden():
    from radmmm_b200 import loss as L
    gd = gold("ops.npz")
    z = syn.hash_uniform("loss.z", (3, 12, 37), -2, 2)
    lsl = [syn.hash_uniform(f"loss.ls{i}", (3, 6, 37), -1, 1) for i in range(3)]
    ldl = [syn.hash_uniform(f"loss.ld{i}", (), -1, 1) for i in range(3)]
    zg = z.to(DEV).requires_grad_(True)
    lsg = [t.to(DEV).requires_grad_(True) for t in lsl]
    ldg = [t.to(DEV).requires_grad_(True) for t in ldl]
    mask = of.length_mask(LENS, 37)[:, None].float().to(DEV)
    l, lp = L.compute_flow_loss(zg, ldg, lsg, LENS.sum(), 12, mask, 0.8)
    close(l, gd["loss"], 1e-6)
    close(lp, gd["loss_prior"], 1e-6)
    l.backward()
    zc = z.double().requires_grad_(True)
    lsc = [t.double().requires_grad_(True) for t in lsl]
    ldc = [t.double().requires_grad_(True) for t in ldl]
    lr, _ = of.flow_loss(zc, ldc, lsc, LENS, 0.8)
    lr.backward()
    close(zg.grad, zc.grad, 1e-7)
    close(lsg[1].grad, lsc[1].grad, 1e-7)
    close(ldg[2].grad, ldc[2].grad, 1e-6)


# ------------------------------------------------------------------------------------------------ spline
def test_spline_kernels():
    lib = N.lib()
    B, Ch, T = 2, 7, 53
    z1 = syn.hash_uniform("sk.z", (B, Ch, T), -3.6, 3.6)
    q = syn.hash_uniform("sk.q", (B, Ch * 65, T), -2, 2)
    lens = torch.tensor([53, 31], dtype=torch.int32)
    zd, qd = z1.to(DEV), q.to(DEV)
    out, ls = torch.empty_like(zd), torch.empty(B, 1, T, device=DEV)
    lensd = lens.to(DEV)
    N.check(lib.radmmm_spline_forward(N.fptr(zd), N.fptr(qd), N.ptr(lensd), N.fptr(out), N.fptr(ls), B, Ch, T, 32,
                                      -3.0, 3.0, 0, N.stream()))
    zc, qc = z1.double().requires_grad_(True), q.double().requires_grad_(True)
    qq = qc.permute(0, 2, 1).reshape(B, T, Ch, 65)
    y, lj = osp.quadratic_spline((zc.permute(0, 2, 1) + 3) / 6, qq[..., :32], qq[..., 32:])
    y_ref = (y * 6 - 3).permute(0, 2, 1)
    ls_ref = lj.sum(-1).unsqueeze(1)
    close(out, y_ref.detach(), 2e-5, what="spline fwd")
    close(ls, ls_ref.detach(), 2e-4, what="spline log_s")
    # inverse round trip through the kernel
    back = torch.empty_like(zd)
    N.check(lib.radmmm_spline_forward(N.fptr(out), N.fptr(qd), N.ptr(lensd), N.fptr(back), None, B, Ch, T, 32,
                                      -3.0, 3.0, 1, N.stream()))
    close(back, z1, 5e-3, what="spline round trip")
    yi, _ = osp.quadratic_spline((y_ref.detach().float().permute(0, 2, 1) + 3) / 6, qq[..., :32].detach().float(),
                                 qq[..., 32:].detach().float(), inverse=True)
    close(back, (yi * 6 - 3).permute(0, 2, 1), 2e-3, what="spline inverse vs oracle")
    # backward
    mask = of.length_mask(lens.long(), T)[:, None].double()
    g1, g2 = syn.hash_uniform("sk.g1", (B, Ch, T)).double(), syn.hash_uniform("sk.g2", (B, 1, T)).double()
    ((y_ref * g1 * mask).sum() + (ls_ref * g2 * mask).sum()).backward()
    dz, dq = torch.empty_like(zd), torch.empty_like(qd)
    g1d, g2d = g1.float().to(DEV), g2.float().to(DEV)
    N.check(lib.radmmm_spline_backward(N.fptr(zd), N.fptr(qd), N.ptr(lensd), N.fptr(g1d), N.fptr(g2d), N.fptr(dz),
                                       N.fptr(dq), B, Ch, T, 32, -3.0, 3.0, N.stream()))
    torch.cuda.synchronize()
    close(dz, zc.grad, 1e-3 * max(1.0, zc.grad.abs().max().item()), what="spline dz")
    close(dq, qc.grad, 1e-3 * max(1.0, qc.grad.abs().max().item()), what="spline dq")


# ------------------------------------------------------------------------------------------------ front end
def test_linear_spline_kernels():
    """Piecewise-linear transform (splines.py:57-238): forward / inverse against the REFERENCE fixtures (ops.npz: spl_yl,
    spl_ljl, spl_xli, spl_ljli were produced by the unmodified splines.py), gradients against oracle autograd, pass-through of
    elements outside [lo, hi], 8- and 32-bin variants."""
    lib = N.lib()
    gd = gold("ops.npz")
    # the fixture's layout: x (50, 7) in [0, 1], q~ (50, 7, 32)  ->  kernel layout (B=1, Ch=7, T=50), q (1, 7*32, 50)
    x = syn.hash_uniform("spl.x", (50, 7), -0.2, 1.2).clamp(0, 1)
    wt = syn.hash_uniform("spl.w", (50, 7, 32), -2, 2)
    z1 = x.t().reshape(1, 7, 50).contiguous().to(DEV)
    q = wt.permute(1, 2, 0).reshape(1, 7 * 32, 50).contiguous().to(DEV)
    lens = torch.tensor([50], dtype=torch.int32, device=DEV)
    out, ls = torch.empty_like(z1), torch.empty(1, 1, 50, device=DEV)
    N.check(lib.radmmm_spline_linear_forward(N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(out), N.fptr(ls), 1, 7, 50, 32, 0.0, 1.0, 0,
                                             N.stream()))
    close(out[0].t(), gd["spl_yl"], 2e-6, what="linear spline fwd vs reference")
    close(ls[0, 0], gd["spl_ljl"], 2e-5, what="linear spline log J vs reference")
    back, lsi = torch.empty_like(z1), torch.empty(1, 1, 50, device=DEV)
    yl = gd["spl_yl"].t().reshape(1, 7, 50).contiguous().to(DEV)
    N.check(lib.radmmm_spline_linear_forward(N.fptr(yl), N.fptr(q), N.ptr(lens), N.fptr(back), N.fptr(lsi), 1, 7, 50, 32, 0.0, 1.0, 1,
                                             N.stream()))
    close(back[0].t(), gd["spl_xli"], 2e-5, what="linear spline inverse vs reference")
    close(lsi[0, 0], gd["spl_ljli"], 2e-4, what="linear spline inverse log J vs reference")
    # gradients + out-of-range pass-through at other shapes / bin counts, vs oracle autograd (fp64)
    for nb in (8, 32):
        B, Ch, T = 2, 5, 41
        z = syn.hash_uniform(f"spll.z{nb}", (B, Ch, T), -3.6, 3.6)
        qq = syn.hash_uniform(f"spll.q{nb}", (B, Ch * nb, T), -2, 2)
        ln = torch.tensor([41, 23], dtype=torch.int32)
        zd, qd, lnd = z.to(DEV), qq.to(DEV), ln.to(DEV)
        o, l = torch.empty_like(zd), torch.empty(B, 1, T, device=DEV)
        N.check(lib.radmmm_spline_linear_forward(N.fptr(zd), N.fptr(qd), N.ptr(lnd), N.fptr(o), N.fptr(l), B, Ch, T, nb, -3.0, 3.0, 0,
                                                 N.stream()))
        zc, qc = z.double().requires_grad_(True), qq.double().requires_grad_(True)
        y, lj = osp.linear_spline((zc.permute(0, 2, 1) + 3) / 6, qc.permute(0, 2, 1).reshape(B, T, Ch, nb))
        y_ref, l_ref = (y * 6 - 3).permute(0, 2, 1), lj.unsqueeze(1)
        close(o, y_ref.detach(), 2e-5, what=f"linear spline fwd ({nb} bins)")
        close(l, l_ref.detach(), 2e-4, what=f"linear spline log J ({nb} bins)")
        outside = (z < -3) | (z > 3)
        # out-of-range elements pass through the spline untouched; like the reference they still go through the
        # (z - lo) / (hi - lo) normalisation and back, so equality holds to rounding, not bit for bit
        close(o.cpu()[outside], z[outside], 1e-6, what=f"linear spline pass-through ({nb} bins)")
        mask = of.length_mask(ln.long(), T)[:, None].double()
        g1, g2 = syn.hash_uniform("spll.g1", (B, Ch, T)).double(), syn.hash_uniform("spll.g2", (B, 1, T)).double()
        ((y_ref * g1 * mask).sum() + (l_ref * g2 * mask).sum()).backward()
        dz, dq = torch.empty_like(zd), torch.empty_like(qd)
        g1d, g2d = g1.float().to(DEV), g2.float().to(DEV)
        N.check(lib.radmmm_spline_linear_backward(N.fptr(zd), N.fptr(qd), N.ptr(lnd), N.fptr(g1d), N.fptr(g2d), N.fptr(dz), N.fptr(dq),
                                                  B, Ch, T, nb, -3.0, 3.0, N.stream()))
        close(dz, zc.grad, 2e-4 * max(1.0, zc.grad.abs().max().item()), what=f"linear spline dz ({nb} bins)")
        close(dq, qc.grad, 2e-4 * max(1.0, qc.grad.abs().max().item()), what=f"linear spline dq ({nb} bins)")


def test_linear_spline_layer():
    """SplineTransformationLayer(use_quadratic=False) (the reference class default, common.py:1010-1020): forward, inverse
    round trip and gradients flow through the FiLM parameter net."""
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.splines import SplineTransformationLayer
    layer = SplineTransformationLayer(8, 10, 2, n_bins=8, left=-3, right=3, bottom=-3, top=3, use_quadratic=False, use_bn=False)
    for n, p in layer.named_parameters():
        p.data.copy_(syn.hash_uniform("spllayer." + n, tuple(p.shape), -0.3, 0.3))
    layer.precision = "fp32"
    layer = layer.to(DEV)
    lens = torch.tensor([21, 9], device=DEV)
    z = syn.hash_uniform("spllayer.z", (2, 8, 21), -3.5, 3.5).to(DEV).requires_grad_(True)
    ctx = syn.hash_uniform("spllayer.ctx", (2, 10, 21), -1, 1).to(DEV)
    seq = SequenceLength(lens, 21)
    zo, ls = layer(z, ctx, seq_lens=seq)
    assert zo.shape == z.shape and ls.shape == (2, 1, 21) and torch.isfinite(zo).all() and torch.isfinite(ls).all()
    with torch.no_grad():
        zi = layer(zo, ctx, inverse=True, seq_lens=seq)
    m = of.length_mask(lens.cpu(), 21)[:, None].double()
    close(zi.cpu().double() * m, z.detach().cpu().double() * m, 5e-4, what="linear spline layer round trip")
    (zo.sum() + ls.sum()).backward()
    assert z.grad is not None and all(p.grad is not None and torch.isfinite(p.grad).all() for p in layer.parameters())


@pytest.mark.parametrize("tag,sr", [("22k", 22050), ("16k", 16000)])
def test_stft_mel(tag, sr):
    from radmmm_b200 import audio_processing as ap
    gd = gold(f"frontend_{tag}.npz")
    n = 256 * 24
    t = torch.arange(n) / sr
    y = 0.5 * syn.hash_uniform("audio" + tag, (2, n), -1, 1)
    for f, a in ((220.0, 0.3), (1333.0, 0.2), (5200.0, 0.1)):
        y = y + a * torch.sin(2 * math.pi * f * t)[None]
    y = y.clamp(-1, 1)
    stft = ap.TacotronSTFT(1024, 256, 1024, 80, sr, 0.0, 8000.0).to(DEV)
    close(stft.mel_basis, ofe.slaney_mel_basis(sr, 1024, 80, 0.0, 8000.0), 0.0)
    mel = stft.mel_spectrogram(y.to(DEV))
    assert mel.shape == (2, 80, 25)
    close(mel, gd["mel"], 2e-4, what="mel vs reference")
    mag, _ = stft.stft_fn.transform(y.to(DEV))
    close(mag[:, ::8], gd["mag"], 2e-3, what="|STFT| vs reference (fp32 dense DFT)")
    close(mag, ofe.stft_magnitude(y, dense=False), 2e-4, what="|STFT| vs fp64 FFT")
    # ragged / odd length input, single utterance
    y2 = y[:1, :5000]
    close(stft.mel_spectrogram(y2.to(DEV)), ofe.mel_spectrogram(y2, sr=sr), 2e-4)
    # the module uses the row-support table (radmmm_stft_mel_sparse); the dense entry point must agree (the lanes sum the
    # same products in a different order), and the table must be the basis' true support
    lib = N.lib()
    yd = y.to(DEV).contiguous()
    dense = torch.empty_like(mel)
    N.check(lib.radmmm_stft_mel(N.fptr(yd), N.fptr(stft.mel_basis), N.fptr(dense), None, 2, n, 1024, 256, 80, 1e-5, N.stream()))
    close(dense, mel, 1e-5, what="dense vs sparse mel")
    sup = stft._support().cpu()
    nz = stft.mel_basis.cpu() != 0
    for m in range(80):
        idx = nz[m].nonzero().flatten()
        assert (int(sup[m, 0]), int(sup[m, 1])) == ((int(idx[0]), int(idx[-1])) if len(idx) else (1, 0))
    # a basis with an all-zero row and a row with two separate runs still works (the range covers both runs)
    odd = stft.mel_basis.clone()
    odd[3] = 0
    odd[5, 400] = 0.25
    stft2 = ap.TacotronSTFT(1024, 256, 1024, 80, sr, 0.0, 8000.0).to(DEV)
    stft2.mel_basis.copy_(odd)
    mag_full, _ = stft2.stft_fn.transform(yd)
    want = torch.log(torch.clamp(torch.matmul(odd.double(), mag_full.double()), min=1e-5))
    close(stft2.mel_spectrogram(yd), want, 2e-4, what="mel with an irregular basis")


# ------------------------------------------------------------------------------------------------ attention
def test_soft_attention():
    lib = N.lib()
    gd = gold("ops.npz")
    shapes = {"key_proj.0.conv": (48, 24, 3), "key_proj.2.conv": (80, 48, 1), "query_proj.0.conv": (160, 80, 3),
              "query_proj.2.conv": (80, 160, 1), "query_proj.4.conv": (80, 80, 1)}
    sd = {}
    for k, s in shapes.items():
        sd[k + ".bias"] = syn.hash_uniform("att." + k + ".bias", (s[0],), -0.2, 0.2)
        sd[k + ".weight_g"] = syn.hash_uniform("att." + k + ".weight_g", (s[0], 1, 1), -0.2, 0.2)
        sd[k + ".weight_v"] = syn.hash_uniform("att." + k + ".weight_v", s, -0.2, 0.2)
    q_in = syn.hash_uniform("att.q", (3, 80, 37), -1, 1)
    k_in = syn.hash_uniform("att.k", (3, 24, 11), -1, 1)
    in_lens = torch.tensor([11, 7, 3], dtype=torch.int32)
    prior = syn.hash_uniform("att.prior", (3, 37, 11), 0.0, 1.0)
    txt = syn.hash_uniform("att.txt", (3, 24, 11), -1, 1)
    q, k = ofe.attention_projections(sd, "", q_in, k_in)
    qd, kd, priord, lensd, txtd = q.to(DEV).contiguous(), k.to(DEV).contiguous(), prior.to(DEV), in_lens.to(DEV), txt.to(DEV)
    for use_prior, ga, gl in ((True, "att", "att_logprob"), (False, "att_noprior", "att_logprob_noprior")):
        attn = torch.empty(3, 1, 37, 11, device=DEV)
        logp = torch.empty_like(attn)
        ctx = torch.empty(3, 24, 37, device=DEV)
        N.check(lib.radmmm_soft_attention(N.fptr(qd), N.fptr(kd), N.fptr(priord) if use_prior else None, N.ptr(lensd),
                                          N.fptr(attn), N.fptr(logp), N.fptr(txtd), N.fptr(ctx), 3, 80, 37, 11, 24,
                                          0.0005, N.stream()))
        close(attn, gd[ga], 2e-6, what=ga)
        close(logp, gd[gl], 2e-5, what=gl)
        if use_prior:
            close(ctx, gd["att_ctx"], 1e-5, what="context")
    # larger ragged case against the oracle
    B, T1, T2 = 2, 203, 61
    q2, k2 = syn.hash_uniform("a2.q", (B, 80, T1), -3, 3), syn.hash_uniform("a2.k", (B, 80, T2), -3, 3)
    pr = syn.hash_uniform("a2.p", (B, T1, T2), 0, 1)
    il = torch.tensor([61, 40], dtype=torch.int32)
    a_ref, l_ref = ofe.soft_attention(q2, k2, il.long(), pr)
    attn, logp = torch.empty(B, 1, T1, T2, device=DEV), torch.empty(B, 1, T1, T2, device=DEV)
    q2d, k2d, prd, ild = q2.to(DEV), k2.to(DEV), pr.to(DEV), il.to(DEV)
    N.check(lib.radmmm_soft_attention(N.fptr(q2d), N.fptr(k2d), N.fptr(prd), N.ptr(ild), N.fptr(attn), N.fptr(logp), None,
                                      None, B, 80, T1, T2, 0, 0.0005, N.stream()))
    close(attn, a_ref, 2e-6)
    close(logp, l_ref, 2e-5)


def test_conv_attention_module_forward_backward():
    """ConvAttention drop-in: state-dict compatible projections, fused attention kernel forward + backward (dq, dk through
    the projections, and the fused context matmul) against oracle autograd."""
    from radmmm_b200 import common
    gd = gold("ops.npz")
    att = common.ConvAttention(n_mel_channels=80, n_text_channels=24, n_att_channels=80)
    sd = {}
    for k, v in att.state_dict().items():
        sd[k] = syn.hash_uniform("att." + k, tuple(v.shape), -0.2, 0.2)
    att.load_state_dict(sd)
    att = att.to(DEV)
    q_in = syn.hash_uniform("att.q", (3, 80, 37), -1, 1)
    k_in = syn.hash_uniform("att.k", (3, 24, 11), -1, 1)
    in_lens = torch.tensor([11, 7, 3])
    prior = syn.hash_uniform("att.prior", (3, 37, 11), 0.0, 1.0)
    txt = syn.hash_uniform("att.txt", (3, 24, 11), -1, 1)
    qg, kg, tg = q_in.to(DEV).requires_grad_(True), k_in.to(DEV).requires_grad_(True), txt.to(DEV).requires_grad_(True)
    amask = ((torch.arange(11)[None] < in_lens[:, None])[..., None] == 0).to(DEV)
    a, lp = att(qg, kg, LENS.to(DEV), amask, key_lens=in_lens.to(DEV), attn_prior=prior.to(DEV))
    close(a, gd["att"], 2e-6, what="module attn vs reference")
    close(lp, gd["att_logprob"], 2e-5, what="module logprob vs reference")
    a2, lp2, ctx = att.forward_with_context(qg, kg, tg, key_lens=in_lens.to(DEV), attn_prior=prior.to(DEV))
    close(ctx, gd["att_ctx"], 1e-5, what="fused context vs reference")
    g1 = syn.hash_uniform("att.g1", (3, 1, 37, 11))
    g2 = syn.hash_uniform("att.g2", (3, 1, 37, 11)) * 0.1
    g3 = syn.hash_uniform("att.g3", (3, 24, 37))
    ((a2 * g1.to(DEV)).sum() + (lp2 * g2.to(DEV)).sum() + (ctx * g3.to(DEV)).sum()).backward()
    # oracle (fp64 autograd)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    qc, kc, tc = q_in.double().requires_grad_(True), k_in.double().requires_grad_(True), txt.double().requires_grad_(True)
    qe, ke = ofe.attention_projections(sdd, "", qc, kc)
    ar, lr = ofe.soft_attention(qe, ke, in_lens, prior.double())
    cr = ofe.attend(tc, ar)
    ((ar * g1.double()).sum() + (lr * g2.double()).sum() + (cr * g3.double()).sum()).backward()
    close(qg.grad, qc.grad, 2e-5 * max(1.0, qc.grad.abs().max().item()), what="d queries")
    close(kg.grad, kc.grad, 2e-5 * max(1.0, kc.grad.abs().max().item()), what="d keys")
    close(tg.grad, tc.grad, 2e-5 * max(1.0, tc.grad.abs().max().item()), what="d txt_enc")
    for n, p in att.named_parameters():
        key = n.replace("weight_g", "weight_g").replace("weight_v", "weight_v")
        ref = sdd[key].grad
        close(p.grad, ref, 5e-5 * max(1e-3, ref.abs().max().item()), what="grad " + n)
