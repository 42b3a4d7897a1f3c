"""ctypes binding of libradmmm_b200.so (the C ABI declared in include/radmmm_b200.h).

There is NO fallback: if the library is missing or an entry point fails, a RuntimeError is raised.  Build the
library with ``python -m radmmm_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RADMMM_B200_LIB") or os.path.join(_HERE, "libradmmm_b200.so")   # env override: A/B builds
MAX_LAYERS = 8
ABI_VERSION = 3
ROW_GAP = 16
MODE_F32, MODE_BF16, MODE_BF16X3 = 0, 1, 2
MODES = {"fp32": MODE_F32, "bf16": MODE_BF16, "bf16x3": MODE_BF16X3}
SCALING = {"tanh": 0, "exp": 1, "sigmoid": 2, "translate": 3}

_fp = C.c_void_p        # every device pointer crosses the boundary as void*


class FlowDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("B", C.c_int32), ("C", C.c_int32), ("Tp", C.c_int32), ("D", C.c_int32),
        ("H", C.c_int32), ("L", C.c_int32), ("scaling_fn", C.c_int32), ("training", C.c_int32),
        ("reserved", C.c_int32),
        ("lens", _fp),
        ("start_g", _fp), ("start_v", _fp), ("start_b", _fp),
        ("in_g", _fp * MAX_LAYERS), ("in_v", _fp * MAX_LAYERS), ("in_b", _fp * MAX_LAYERS),
        ("rs_g", _fp * MAX_LAYERS), ("rs_v", _fp * MAX_LAYERS), ("rs_b", _fp * MAX_LAYERS),
        ("end_w", _fp), ("end_b", _fp),
        ("W", _fp), ("W_inv", _fp), ("W_T", _fp), ("mean", _fp),
        ("prepared", _fp),
        ("ctx_rows", _fp),
        ("workspace", _fp),
        ("side_stream", _fp),
    ]


class FlowGrads(C.Structure):
    _fields_ = [
        ("start_g", _fp), ("start_v", _fp), ("start_b", _fp),
        ("in_g", _fp * MAX_LAYERS), ("in_v", _fp * MAX_LAYERS), ("in_b", _fp * MAX_LAYERS),
        ("rs_g", _fp * MAX_LAYERS), ("rs_v", _fp * MAX_LAYERS), ("rs_b", _fp * MAX_LAYERS),
        ("end_w", _fp), ("end_b", _fp),
        ("W", _fp),
    ]


_i, _ll, _f, _sz = C.c_int, C.c_longlong, C.c_float, C.c_size_t
_P = C.POINTER
# name -> (restype, argtypes); mirrors include/radmmm_b200.h one to one (tests check the symbol list against the header)
SIGNATURES = {
    "radmmm_abi_version": (_i, []),
    "radmmm_last_error": (C.c_char_p, []),
    "radmmm_launch_count": (_ll, []),
    "radmmm_debug_trace": (None, [_fp, _i, _i]),
    "radmmm_profile_enable": (None, [_i]),
    "radmmm_profile_collect": (_i, [_i, _P(C.c_int), _P(C.c_double), _P(C.c_double)]),
    "radmmm_sizeof_flow_desc": (_sz, []),
    "radmmm_sizeof_flow_grads": (_sz, []),
    "radmmm_rows": (_i, [_i, _i]),
    "radmmm_pitch": (_i, [_i]),
    "radmmm_flow_prepared_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "radmmm_flow_workspace_bytes": (_sz, [_i] * 8),
    "radmmm_flow_backward_scratch_bytes": (_sz, [_i] * 7),
    "radmmm_context_rows_bytes": (_sz, [_i] * 4),
    "radmmm_flow_prepare": (_i, [_P(FlowDesc), _fp]),
    "radmmm_context_rows": (_i, [_i, _fp, _fp, _i, _i, _i, _fp, _fp]),
    "radmmm_context_rows_backward": (_i, [_fp, _fp, _i, _i, _i, _fp, _i, _fp]),
    "radmmm_flow_forward": (_i, [_P(FlowDesc), _fp, _fp, _fp, _fp, _fp, _fp]),
    "radmmm_flow_inverse": (_i, [_P(FlowDesc), _fp, _fp, _fp, _fp, _fp]),
    "radmmm_flow_backward": (_i, [_P(FlowDesc), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _P(FlowGrads), _fp, _fp]),
    "radmmm_inv1x1": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp]),
    "radmmm_inv1x1_wgrad": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp]),
    "radmmm_coupling_forward": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _fp]),
    "radmmm_coupling_backward": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp]),
    "radmmm_masked_sum": (_i, [_fp, _fp, _i, _i, _i, _i, _fp, _fp]),
    "radmmm_masked_sum_backward": (_i, [_fp, _fp, _i, _i, _i, _i, _fp, _f, _fp, _fp]),
    "radmmm_conv_rows": (_i, [_i, _fp, _ll, _ll, _fp, _ll, _ll, _ll, _fp, _fp, _ll, _i, _i, _i, _i, _i, _fp]),
    "radmmm_wgrad_rows": (_i, [_i, _fp, _ll, _ll, _fp, _ll, _ll, _fp, _ll, _ll, _i, _i, _i, _i, _i, _i, _fp]),
    "radmmm_radam_chunk_elems": (_i, []),
    "radmmm_radam_step": (_i, [_fp, _fp, _fp, _i, _fp, _fp, _fp]),
    "radmmm_lstm_workspace_bytes": (_sz, [_i, _i]),
    "radmmm_lstm_forward": (_i, [_i, _fp, _fp, _fp, _fp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp]),
    "radmmm_lstm_backward": (_i, [_i, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp, _fp, _fp]),
    "radmmm_cast_rows": (_i, [_i, _fp, _ll, _fp, _ll, _fp]),
    "radmmm_spline_forward": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _i, _fp]),
    "radmmm_spline_backward": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _fp]),
    "radmmm_spline_linear_forward": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _i, _fp]),
    "radmmm_spline_linear_backward": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _fp]),
    "radmmm_stft_mel": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _f, _fp]),
    "radmmm_mel_support": (_i, [_fp, _i, _i, _fp, _fp]),
    "radmmm_stft_mel_sparse": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _f, _fp]),
    "radmmm_soft_attention": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _f, _fp]),
    "radmmm_soft_attention_backward_workspace_bytes": (_ll, [_i, _i, _i]),
    "radmmm_soft_attention_backward": (_i, [_fp] * 12 + [_i, _i, _i, _i, _i, _f, _fp, _ll, _fp]),
    "radmmm_mas_workspace_bytes": (_ll, [_i, _i, _i]),
    "radmmm_mas_width1": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp, _ll, _fp]),
    "radmmm_attention_ctc": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _fp]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the native library (once).  Raises if it has not been built -- there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"radmmm_b200: native library {LIB_PATH} is missing. Build it with `python -m radmmm_b200.build` "
                "(needs nvcc; there is no CPU or eager fallback).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.radmmm_abi_version() != ABI_VERSION:
            raise RuntimeError("radmmm_b200: ABI version mismatch between the Python binding and the library")
        if handle.radmmm_sizeof_flow_desc() != C.sizeof(FlowDesc) or handle.radmmm_sizeof_flow_grads() != C.sizeof(FlowGrads):
            raise RuntimeError("radmmm_b200: struct layout mismatch between the Python binding and the library")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().radmmm_last_error()
        raise RuntimeError(f"radmmm_b200 native call failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous, on CUDA."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("radmmm_b200: expected a CUDA tensor (the kernels have no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("radmmm_b200: expected a contiguous tensor")
    return t.data_ptr()


def fptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is not None and t.dtype != torch.float32:
        raise RuntimeError(f"radmmm_b200: expected float32, got {t.dtype}")
    return ptr(t)


def stream() -> int:
    """Current stream of the CURRENT device.  The entry points that take tensors (RADMMMFlow.forward / infer, FlowStep,
    the standalone modules via ``on_device_of``) switch to the tensors' device first; backward passes run on autograd's
    per-device threads, which have the right device set already."""
    return torch.cuda.current_stream().cuda_stream


def on_device_of(t: torch.Tensor):
    """Context manager: make ``t``'s device current (a no-op when it already is)."""
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        return torch.cuda.device(t.device)
    return contextlib.nullcontext()


def rows(batch: int, tp: int) -> int:
    return (batch * (tp + ROW_GAP) + 255) // 256 * 256


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b
