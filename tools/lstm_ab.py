#!/usr/bin/env python
"""A/B of the context bi-LSTM (models/radmmm.py:137-146) at the benchmark shapes: our kernels (radmmm_b200.lstm.context_lstm:
input-projection contraction + persistent recurrence + gradient contractions) against torch's packed nn.LSTM (cuDNN, fp32,
TF32 off) -- forward and forward+backward, eager calls, CUDA events.  Usage: python tools/lstm_ab.py [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from radmmm_b200.lstm import context_lstm  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    Tp = (int(sys.argv[1]) if len(sys.argv) > 1 else 800) // 2
    n_in, hid = 1060, 528
    lstm = torch.nn.LSTM(n_in, hid, 1, batch_first=True, bidirectional=True).to(dev)
    print(f"context bi-LSTM {n_in} -> 2 x {hid}, T' = {Tp} grouped frames, lengths U[T'/2, T']")
    for B in (8, 32, 64):
        g = torch.Generator().manual_seed(B)
        lens = (Tp // 2 + torch.randint(0, Tp // 2 + 1, (B,), generator=g)).clamp(max=Tp)
        lens[0] = Tp
        lens = torch.sort(lens, descending=True)[0]
        x = torch.randn(B, Tp, n_in, device=dev, requires_grad=True)
        gout = torch.randn(B, Tp, 2 * hid, device=dev)
        lens_d = lens.to(dev)

        def ours(bwd, prec):
            y = context_lstm(lstm, x, lens_d, prec)
            if bwd:
                (y * gout).sum().backward()

        def cudnn(bwd):
            packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=True)
            lstm.flatten_parameters()
            y, _ = lstm(packed)
            y, _ = torch.nn.utils.rnn.pad_packed_sequence(y, batch_first=True, total_length=Tp)
            if bwd:
                (y * gout).sum().backward()

        for prec in ("bf16", "bf16x3"):
            print(f"B={B:3d} ours[{prec:6s}]  fwd {timeit(lambda: ours(False, prec)):8.3f} ms   fwd+bwd {timeit(lambda: ours(True, prec)):8.3f} ms")
        print(f"B={B:3d} cuDNN packed  fwd {timeit(lambda: cudnn(False)):8.3f} ms   fwd+bwd {timeit(lambda: cudnn(True)):8.3f} ms")


if __name__ == "__main__":
    main()
