#!/usr/bin/env python
"""Which pipeline bounds the tcgen05 contraction kernel?  Times the dilated k=5 conv shape (R rows x 1024 x 5*1024)
with parts of the kernel disabled (RADMMM_B200_TC_PROBE bits: 1 no epilogue, 2 no MMA, 4 no TMA) and with the CTA-pair
variant on/off.  Diagnostic only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from radmmm_b200 import _native as N  # noqa: E402


def main():
    lib = N.lib()
    dev = "cuda"
    H = 1024
    for R in (3328, 3328 * 4):
        x = (torch.randn(R, H, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(5 * H, H, device=dev) * 0.02).to(torch.bfloat16)
        y = torch.empty(R, H, device=dev)
        flops = 2.0 * R * H * 5 * H

        def run(taps):
            N.check(lib.radmmm_conv_rows(N.MODE_BF16, x.data_ptr(), H, R * H, w.data_ptr(), H, 5 * H * H, H * H, None,
                                         y.data_ptr(), H, R, H, H, taps, 2, N.stream()))

        for pair in ("1", "0"):
            os.environ["RADMMM_B200_TC_PAIR"] = pair
            for bits in (0, 1, 2, 3, 4, 5, 6, 7):
                os.environ["RADMMM_B200_TC_PROBE"] = str(bits)
                for taps in (5, 1):
                    for _ in range(3):
                        run(taps)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    n = 20
                    e0.record()
                    for _ in range(n):
                        run(taps)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / n
                    print(f"R={R} pair={pair} probe={bits} (epi={'off' if bits & 1 else 'on '} mma={'off' if bits & 2 else 'on '} "
                          f"tma={'off' if bits & 4 else 'on '}) taps={taps}: {us:7.1f} us  "
                          f"{flops * taps / 5 / us / 1e6:7.1f} TFLOP/s")
    os.environ["RADMMM_B200_TC_PROBE"] = "0"
    os.environ["RADMMM_B200_TC_PAIR"] = "1"


if __name__ == "__main__":
    main()
