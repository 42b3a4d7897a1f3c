"""Emulate tensor-core operand rounding inside the oracle to size the parity budget (precision modes).
Only the WN convolutions (the tensor-core GEMMs) are rounded; everything else stays fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import flow as of
from radmmm_b200 import synthetic as syn

real_conv = F.conv1d
MODE = "fp32"

def split(x, dt, n):
    parts, r = [], x
    for _ in range(n):
        h = r.to(dt).to(torch.float32)
        parts.append(h); r = r - h
    return parts

def conv_emul(x, w, b=None, **kw):
    if MODE == "fp32" or w.shape[0] < 100 and w.shape[1] < 100:
        return real_conv(x, w, b, **kw)
    dt, n = {"bf16": (torch.bfloat16, 1), "bf16x3": (torch.bfloat16, 2), "bf16x6": (torch.bfloat16, 3),
             "fp16": (torch.float16, 1), "fp16x3": (torch.float16, 2)}[MODE]
    xs, ws = split(x, dt, n), split(w, dt, n)
    y = 0
    for i, xi in enumerate(xs):
        for j, wj in enumerate(ws):
            if i + j < n:
                y = y + real_conv(xi.double(), wj.double(), None, **kw)
    y = y.float()
    return y + b.view(1, -1, 1) if b is not None else y

F.conv1d = conv_emul
cfg = of.DecoderConfig.radmmm()
sd = syn.synthetic_state_dict()
bt = syn.synthetic_batch(2, 96, tag="decoder_full.npz")
lstm = of.build_context_lstm(sd, cfg)
res = {}
for MODE in ("fp32", "bf16", "fp16", "bf16x3", "fp16x3", "bf16x6"):
    with torch.no_grad():
        out = of.decoder_forward(sd, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                                 bt["energy_avg"], bt["accent_vecs"], lstm=lstm)
        lens_g = bt["out_lens"] // 2
        loss, _ = of.flow_loss(out["z_mel"], out["log_det_W_list"], out["log_s_list"], lens_g)
        m = of.length_mask(lens_g, 48)[:, None]
        res[MODE] = (out["z_mel"] * m, sum((ls * m).sum() for ls in out["log_s_list"]), loss)
        residual = syn.hash_uniform("residual", (2, 160, 48), -1.5, 1.5)
        mel = of.decoder_inverse(sd, cfg, residual, out["context_w_spkvec"], lens_g)
        mm = of.length_mask(bt["out_lens"], 96)[:, None]
        res[MODE] += (mel * mm,)
    z0, ls0, l0, mel0 = res["fp32"]
    z, ls, l, mel = res[MODE]
    print(f"{MODE:8s} z max-abs {float((z-z0).abs().max()):.3e}  sum log_s rel {float(((ls-ls0)/ls0).abs()):.3e}  "
          f"loss rel {float(((l-l0)/l0).abs()):.3e}  inverse mel max-abs {float((mel-mel0).abs().max()):.3e}")
