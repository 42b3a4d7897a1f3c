"""Helpers shared by the GPU parity tests."""
import os

import numpy as np
import torch

from radmmm_b200 import _native as N

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def gold(name):
    return {k: torch.from_numpy(v) if v.dtype.kind in "fiub" else v for k, v in np.load(os.path.join(GOLD, name)).items()}


def err(a, b):
    return (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max().item()


def close(a, b, atol, rtol=0.0, what=""):
    b = torch.as_tensor(b).double().cpu()
    e = err(a, b)
    lim = atol + rtol * b.abs().max().item()
    assert e <= lim, f"{what} max-abs err {e:.3e} > {lim:.3e}"
    return e


def cast_rows(x: torch.Tensor, mode: int):
    """fp32 matrix -> (buffer, ld, plane_stride) in the act format of `mode` (via radmmm_cast_rows)."""
    lib = N.lib()
    x = x.contiguous().float()
    n = x.numel()
    if mode == N.MODE_F32:
        return x, x.shape[1], n
    buf = torch.empty((2 if mode == N.MODE_BF16X3 else 1) * n, dtype=torch.bfloat16, device=x.device)
    N.check(lib.radmmm_cast_rows(mode, N.fptr(x), n, N.ptr(buf), n, N.stream()))
    return buf, x.shape[1], n


def act_to_float(buf, shape, mode):
    n = int(np.prod(shape))
    if mode == N.MODE_F32:
        return buf.reshape(shape)
    out = buf[:n].float()
    if mode == N.MODE_BF16X3:
        out = out + buf[n:2 * n].float()
    return out.reshape(shape)
