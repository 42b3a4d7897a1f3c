"""Hard alignment on the GPU (reference: alignment.py:31-59 ``mas_width1``; caller tts_lightning_modules.py:270-284
``TTSModel.binarize_attention``).

The reference copies the (B, 1, T1, T2) soft attention to the host, runs one numba Viterbi per utterance and copies the
result back -- a device->host sync directly before the decoder call of every step after ``binarization_start_iter``.  Here
the whole batch is ONE kernel launch on the current stream (one CTA per utterance, csrc/alignment.cu), no host round trip,
lengths read on the device.

``binarize_attention(attn, in_lens, out_lens)`` is the drop-in for the method body; ``mas_width1`` / ``mas`` keep the
reference's single-map signature (numpy in -> numpy out, tensor in -> tensor out) for callers that use it directly.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N


def _as_lens(lens, device) -> torch.Tensor:
    if not torch.is_tensor(lens):
        lens = torch.as_tensor(np.asarray(lens))
    return lens.to(device=device, dtype=torch.int32).contiguous()


@torch.no_grad()
def binarize_attention(attn: torch.Tensor, in_lens, out_lens, is_log: bool = False) -> torch.Tensor:
    """tts_lightning_modules.py:270-284.  attn: (B, 1, max_mel_len, max_text_len) soft attention on the GPU; returns the 0/1
    map of every utterance's ``attn[b, 0, :out_len, :in_len]`` (zeros elsewhere), no gradient.  ``is_log``: ``attn`` already
    holds log-probabilities (skips the logarithm the reference takes)."""
    if not attn.is_cuda:
        raise RuntimeError("radmmm_b200.alignment: expected a CUDA tensor (the kernels have no CPU path)")
    if attn.dim() != 4 or attn.shape[1] != 1:
        raise ValueError(f"binarize_attention: expected (B, 1, T1, T2), got {tuple(attn.shape)}")
    lib = N.lib()
    a = attn.detach().float().contiguous()
    b, _, t1, t2 = a.shape
    out = torch.empty_like(a)
    if a.numel() == 0:
        return out
    with N.on_device_of(a):
        il, ol = _as_lens(in_lens, a.device), _as_lens(out_lens, a.device)
        need = lib.radmmm_mas_workspace_bytes(b, t1, t2)
        ws = torch.empty(need, dtype=torch.uint8, device=a.device) if need > 0 else None
        N.check(lib.radmmm_mas_width1(N.fptr(a), N.ptr(il), N.ptr(ol), N.fptr(out), b, t1, t2, int(is_log),
                                      N.ptr(ws), need, N.stream()))
    return out


def mas_width1(attn_map):
    """alignment.py:31-59 for one (mel frames x text positions) map.  numpy in -> numpy out (as the reference), CUDA tensor in
    -> CUDA tensor out."""
    as_numpy = not torch.is_tensor(attn_map)
    a = torch.as_tensor(np.asarray(attn_map, dtype=np.float32)) if as_numpy else attn_map
    if as_numpy:
        a = a.cuda()
    t1, t2 = a.shape
    out = binarize_attention(a.reshape(1, 1, t1, t2), [t2], [t1])[0, 0]
    return out.cpu().numpy() if as_numpy else out


mas = mas_width1
