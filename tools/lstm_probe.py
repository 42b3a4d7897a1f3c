#!/usr/bin/env python
"""Where does a step of the persistent LSTM recurrence go?  Times radmmm_lstm_forward (H=528, B=8, T'=400) with parts of
the step disabled (RADMMM_B200_LSTM_PROBE bits: 1 no mat-vec, 2 no exchange reload, 4 no grid barrier).  Diagnostic."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from radmmm_b200 import _native as N  # noqa: E402


MODE = N.MODE_BF16 if (len(sys.argv) > 1 and sys.argv[1] == 'bf16') else N.MODE_F32


def main():
    lib = N.lib()
    dev = "cuda"
    B, T, H = 8, 400, 528
    R = N.rows(B, T)
    xproj = torch.randn(R, 8 * H, device=dev) * 0.1
    whf = torch.randn(4 * H, H, device=dev) * 0.03
    whr = torch.randn(4 * H, H, device=dev) * 0.03
    lens = torch.tensor([400, 380, 360, 330, 300, 280, 250, 210], dtype=torch.int32, device=dev)
    out = torch.zeros(B, T, 2 * H, device=dev)
    gates = torch.empty(R, 8 * H, device=dev)
    cst = torch.empty(R, 2 * H, device=dev)
    ws = torch.empty(lib.radmmm_lstm_workspace_bytes(B, H), dtype=torch.uint8, device=dev)

    def run():
        N.check(lib.radmmm_lstm_forward(MODE, N.fptr(xproj), N.fptr(whf), N.fptr(whr), N.ptr(lens), B, T, H, N.fptr(out),
                                        N.fptr(gates), N.fptr(cst), N.ptr(ws), N.stream()))

    for bits in (0, 1, 2, 3, 4, 5, 6, 7):
        os.environ["RADMMM_B200_LSTM_PROBE"] = str(bits)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"probe={bits} (matvec={'off' if bits & 1 else 'on '} reload={'off' if bits & 2 else 'on '} "
              f"barrier={'off' if bits & 4 else 'on '}): {ms * 1e3:8.1f} us  = {ms * 1e3 / T:6.2f} us/step")
    os.environ["RADMMM_B200_LSTM_PROBE"] = "0"
    # backward recurrence (no probe bits): gates / cell states from the forward pass above, random incoming gradient
    run()
    dout = torch.randn(B, T, 2 * H, device=dev) * 0.1
    dg = torch.zeros(R, 8 * H, device=dev)

    def run_bwd():
        N.check(lib.radmmm_lstm_backward(MODE, N.fptr(dout), N.fptr(gates), N.fptr(cst), N.fptr(whf), N.fptr(whr), N.ptr(lens), B, T, H,
                                         N.fptr(dg), N.ptr(ws), N.stream()))

    run_bwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run_bwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"backward: {ms * 1e3:8.1f} us  = {ms * 1e3 / T:6.2f} us/step")


if __name__ == "__main__":
    main()
