"""Launch the 1x1 invertible conv kernels alone at the benchmark shape (for `ncu -k regex:inv1x1`)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radmmm_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
lib = N.lib()
B, C, Tp = 8, 160, 400
z, W, zo = torch.randn(B, C, Tp, device=dev), torch.randn(C, C, device=dev), torch.empty(B, C, Tp, device=dev)
st = N.stream()
for _ in range(6):
    N.check(lib.radmmm_inv1x1(N.fptr(z), N.fptr(W), None, None, N.fptr(zo), B, C, C, Tp, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(2_000_000)
e0.record()
for _ in range(20):
    N.check(lib.radmmm_inv1x1(N.fptr(z), N.fptr(W), None, None, N.fptr(zo), B, C, C, Tp, st))
e1.record()
e1.synchronize()
print("inv1x1 us/launch (20 back to back):", e0.elapsed_time(e1) * 1e3 / 20)
