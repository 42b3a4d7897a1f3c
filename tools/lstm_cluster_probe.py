#!/usr/bin/env python
"""Phase timers of the cluster-resident LSTM recurrence (csrc/lstm_cluster.cu, radmmm_debug_trace with max_launches < 0):
where a time step's microseconds go -- waiting for the exchange, mat-vec, barriers, gate math, pushes.  Also times the kernels
with CUDA events (untraced build).  Diagnostic.  Usage: python tools/lstm_cluster_probe.py [batch] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from radmmm_b200 import _native as N  # noqa: E402


def main():
    lib = N.lib()
    dev = "cuda"
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = (int(sys.argv[2]) if len(sys.argv) > 2 else 800) // 2
    H = 528
    R = N.rows(B, T)
    xproj = torch.randn(R, 8 * H, device=dev) * 0.1
    whf = torch.randn(4 * H, H, device=dev) * 0.03
    whr = torch.randn(4 * H, H, device=dev) * 0.03
    lens = torch.linspace(T, T // 2, B).to(torch.int32).to(dev)
    out = torch.zeros(B, T, 2 * H, device=dev)
    gates = torch.empty(R, 8 * H, device=dev)
    cst = torch.empty(R, 2 * H, device=dev)
    dout = torch.randn(B, T, 2 * H, device=dev) * 0.1
    dg = torch.zeros(R, 8 * H, device=dev)
    ws = torch.empty(lib.radmmm_lstm_workspace_bytes(B, H), dtype=torch.uint8, device=dev)

    def fwd(mode):
        N.check(lib.radmmm_lstm_forward(mode, N.fptr(xproj), N.fptr(whf), N.fptr(whr), N.ptr(lens), B, T, H, N.fptr(out),
                                        N.fptr(gates), N.fptr(cst), N.ptr(ws), N.stream()))

    def bwd(mode):
        N.check(lib.radmmm_lstm_backward(mode, N.fptr(dout), N.fptr(gates), N.fptr(cst), N.fptr(whf), N.fptr(whr), N.ptr(lens), B, T, H,
                                         N.fptr(dg), N.ptr(ws), N.stream()))

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5 * 1e3

    for name, mode in (("cluster bf16", N.MODE_BF16), ("cooperative fp32", N.MODE_F32)):
        tf, tb = timeit(lambda: fwd(mode)), timeit(lambda: bwd(mode))
        print(f"B={B} T'={T} {name:17s}: forward {tf:8.1f} us = {tf / T:5.2f} us/step   backward {tb:8.1f} us = {tb / T:5.2f} us/step")
    clk = 1965.0     # MHz
    buf = torch.zeros(32 * 8, dtype=torch.int64, device=dev)
    lib.radmmm_debug_trace(buf.data_ptr(), 0, -1)
    for name, fn, labels in (("forward", fwd, ["wait h", "mat-vec", "cp.async+bar", "gates", "barrier", "push"]),
                             ("backward", bwd, ["wait dh", "cp.async", "gate grads", "barrier", "mat-vec+push", "-"])):
        buf.zero_()
        fn(N.MODE_BF16)
        torch.cuda.synchronize()
        t = buf.cpu().reshape(32, 8).double() / clk / T          # us per step per CTA
        print(f"{name}: us per step, thread 0 of each CTA (traced build: a little slower than the numbers above)")
        print("   " + " ".join(f"{l:>13}" for l in labels[:6]) + "         total")
        for tag, rows in (("mean", t.mean(0)), ("min ", t.min(0).values), ("max ", t.max(0).values)):
            print(f"{tag}" + " ".join(f"{float(v):13.3f}" for v in rows[:6]) + f"   {float(rows[:6].sum()):8.3f}")
    lib.radmmm_debug_trace(None, 0, -1)
    # what a step's pointwise phase is made of: parts disabled one at a time (results are wrong, timings are what matters)
    for bits, what in ((1, "no saved-state stores"), (2, "no transcendentals"), (4, "no cp.async prefetch"), (7, "none of the three")):
        os.environ["RADMMM_B200_LSTM_PROBE"] = str(bits)
        tf, tb = timeit(lambda: fwd(N.MODE_BF16)), timeit(lambda: bwd(N.MODE_BF16))
        print(f"probe {bits} ({what:24s}): forward {tf / T:5.2f} us/step   backward {tb / T:5.2f} us/step")
    os.environ["RADMMM_B200_LSTM_PROBE"] = "0"


if __name__ == "__main__":
    main()
