#!/usr/bin/env python
"""In-kernel timeline of the tensor-core contraction kernels of ONE flow step (forward + backward) at the benchmark shape
(radmmm_debug_trace): where a launch's microseconds go -- set-up, first operands, main loop, epilogue, teardown.
Diagnostic only.  Usage: python tools/gemm_timeline.py [batch] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from radmmm_b200 import _native as N  # noqa: E402
from radmmm_b200 import common, synthetic as syn  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    os.environ["RADMMM_B200_SIDE_STREAM"] = "0"          # one stream: launches run one after the other, in issue order
    dev = "cuda"
    lib = N.lib()
    C, D, H, L, Tp = 160, 1056, 1024, 4, frames // 2
    layer = common.AffineTransformationLayer(C, D, L, affine_model="wavenet", scaling_fn="tanh", n_channels=H, use_partial_padding=True)
    pre = "flows.0.coupling_tfn."
    layer.load_state_dict({k[len(pre):]: v for k, v in syn.synthetic_state_dict(n_flows=1).items() if k.startswith(pre)})
    layer.precision = "bf16"
    layer = layer.to(dev)
    bt = syn.synthetic_batch(batch, frames, tag="bench.rank0")
    lens = (bt["out_lens"] // 2).to(dev)
    z = torch.randn(batch, C, Tp, device=dev, requires_grad=True)
    ctx = torch.randn(batch, D, Tp, device=dev)
    seq = common.SequenceLength(lens, Tp)
    for _ in range(2):
        zo, ls = layer(z, ctx, seq_lens=seq)
        (zo.sum() + ls.sum()).backward()
    torch.cuda.synchronize()
    max_ctas, max_launches, slots = 160, 64, 10
    buf = torch.zeros(max_launches * max_ctas * slots, dtype=torch.int64, device=dev)
    lib.radmmm_debug_trace(buf.data_ptr(), max_ctas, max_launches)
    zo, ls = layer(z, ctx, seq_lens=seq)
    (zo.sum() + ls.sum()).backward()
    torch.cuda.synchronize()
    lib.radmmm_debug_trace(None, 0, 0)
    t = buf.cpu().reshape(max_launches, max_ctas, slots)
    clk = 1.965          # GHz (clocks.max.sm; the bench runs at max clocks at this load)
    names = ["setup", "->1st TMA", "->1st operands", "main loop", "->acc visible", "epilogue", "teardown"]
    print(f"B={batch} T={frames}: per launch, median over CTAs, microseconds at {clk} GHz; span = globaltimer first entry -> last exit")
    print("launch ctas   " + " ".join(f"{n:>14}" for n in names) + "     total   span_us  entry_spread_us")
    for i in range(max_launches):
        a = t[i]
        used = a[:, 0] != 0
        if not used.any():
            continue
        a = a[used].double()
        seg = torch.stack([a[:, 1] - a[:, 0], a[:, 2] - a[:, 1], a[:, 3] - a[:, 1], a[:, 4] - a[:, 3], a[:, 5] - a[:, 4],
                           a[:, 6] - a[:, 5], a[:, 7] - a[:, 6]], 1) / (clk * 1e3)
        lead = a[:, 3] != 0           # only the pair leaders run the MMA warp
        med = [float(seg[:, 0].median()), float(seg[:, 1].median()), float(seg[lead, 2].median()) if lead.any() else 0.0,
               float(seg[lead, 3].median()) if lead.any() else 0.0, float((a[lead, 5] - a[lead, 4]).median() / (clk * 1e3)) if lead.any() else 0.0,
               float(seg[:, 5].median()), float(seg[:, 6].median())]
        total = float(((a[:, 7] - a[:, 0]) / (clk * 1e3)).median())
        span = float(a[:, 9].max() - a[:, 8].min()) / 1e3
        spread = float(a[:, 8].max() - a[:, 8].min()) / 1e3
        print(f"{i:5d} {int(used.sum()):5d}   " + " ".join(f"{m:14.2f}" for m in med) + f"  {total:8.2f}  {span:8.2f}  {spread:8.2f}")


if __name__ == "__main__":
    main()
