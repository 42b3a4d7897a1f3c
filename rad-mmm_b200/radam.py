"""Fused multi-tensor RAdam with global-norm gradient clipping (reference: radam.py:45-142 ``RAdam`` and Lightning's
``gradient_clip_val: 1.0`` / ``gradient_clip_algorithm: norm``, configs/RADMMM_train_config.yaml:7-8).

Drop-in for ``radam.RAdam`` -- same constructor arguments, the same per-parameter state (``step``, ``exp_avg``,
``exp_avg_sq``), the same arithmetic -- but one step is three kernel launches for ALL parameters instead of a Python loop
of ~12 ATen kernels per tensor with fp32 copies, has no host synchronisation and can be captured in a CUDA graph
(``GraphedTrainStep(after_backward=optimizer.step)``).  ``max_grad_norm`` folds what Lightning does before
``optimizer.step()`` -- ``torch.nn.utils.clip_grad_norm_(parameters, max_grad_norm)`` -- into the same pass (leave it
``None`` when the trainer clips itself).

Parameters and gradients must be fp32 CUDA tensors (the decoder's are).  The kernels read the gradients wherever they live;
with the decoder's gradient arena (``radmmm_b200.graphs.grad_arena`` / ``ddp.BucketedGradReducer``) the pointers never change
and the device tables are built once.
"""
from __future__ import annotations

import struct
from typing import List, Optional

import torch
from torch.optim.optimizer import Optimizer

from . import _native as N


class RAdam(Optimizer):
    """radam.py:45-142.  ``step`` counts per parameter in the reference are all equal; here one counter per group lives on
    the device (and is mirrored on the host for ``state_dict``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, max_grad_norm: Optional[float] = None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._tables = {}          # group index -> dict(recs, chunk_tensor, chunk_off, n_chunks, ptrs, cfg, cfg_host, dstate)

    # ------------------------------------------------------------------------------------------------ tables
    def _params_with_grad(self, group) -> List[torch.nn.Parameter]:
        return [p for p in group["params"] if p.grad is not None]

    def _ensure_state(self, group, params):
        """exp_avg / exp_avg_sq of a group live in two flat buffers; the per-parameter state tensors are views of them."""
        missing = [p for p in params if "exp_avg" not in self.state[p]]
        if not missing:
            return
        n = sum((p.numel() + 63) // 64 * 64 for p in missing)
        dev = missing[0].device
        flat_m, flat_v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        off = 0
        for p in missing:
            st = self.state[p]
            st.setdefault("step", 0)
            st["exp_avg"] = flat_m[off:off + p.numel()].view_as(p)
            st["exp_avg_sq"] = flat_v[off:off + p.numel()].view_as(p)
            off += (p.numel() + 63) // 64 * 64

    def _build(self, gi: int, group, params):
        lib = N.lib()
        dev = params[0].device
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                raise RuntimeError("radmmm_b200.radam.RAdam: parameters and gradients must be fp32 CUDA tensors")
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("radmmm_b200.radam.RAdam: parameters and gradients must be contiguous")
        self._ensure_state(group, params)
        chunk = lib.radmmm_radam_chunk_elems()
        ptrs = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                      p.numel()) for p in params)
        recs = b"".join(struct.pack("<QQQQq", *r) for r in ptrs)
        ct, co = [], []
        for i, p in enumerate(params):
            for o in range(0, p.numel(), chunk):
                ct.append(i)
                co.append(o)
        old = self._tables.get(gi)
        tab = {
            "ptrs": ptrs, "n_chunks": len(ct), "n_params": sum(p.numel() for p in params),
            "recs": torch.frombuffer(bytearray(recs), dtype=torch.uint8).to(dev),
            "chunk_tensor": torch.tensor(ct, dtype=torch.int32, device=dev),
            "chunk_off": torch.tensor(co, dtype=torch.int64, device=dev),
            "dstate": old["dstate"] if old else torch.zeros(8, dtype=torch.float64, device=dev),
            "cfg": old["cfg"] if old else torch.zeros(8, dtype=torch.float64, device=dev),
            "cfg_host": None,
        }
        if not old:       # resume: the device step counter starts from the loaded per-parameter step
            tab["dstate"][1] = float(max((self.state[p].get("step", 0) for p in params), default=0))
        self._tables[gi] = tab
        return tab

    # ------------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = N.lib()
        capturing = torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            params = self._params_with_grad(group)
            if not params:
                continue
            tab = self._tables.get(gi)
            if tab is None or (not capturing and tab["ptrs"] != tuple(
                    (p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr() if "exp_avg" in self.state[p] else 0,
                     self.state[p]["exp_avg_sq"].data_ptr() if "exp_avg_sq" in self.state[p] else 0, p.numel()) for p in params)):
                if capturing:
                    raise RuntimeError("radmmm_b200.radam.RAdam: run one eager step before capturing (device tables are built on "
                                       "the first step and need stable parameter / gradient addresses)")
                tab = self._build(gi, group, params)
            cfg_host = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                        float(group["weight_decay"]), float(group["max_grad_norm"] or 0.0))
            if tab["cfg_host"] != cfg_host:                    # hyper-parameters live on the device (a captured graph reads them)
                if capturing:
                    raise RuntimeError("radmmm_b200.radam.RAdam: hyper-parameters changed during capture")
                tab["cfg"][:6].copy_(torch.tensor(cfg_host, dtype=torch.float64))
                tab["cfg_host"] = cfg_host
            with torch.cuda.device(params[0].device):
                N.check(lib.radmmm_radam_step(N.ptr(tab["recs"]), N.ptr(tab["chunk_tensor"]), N.ptr(tab["chunk_off"]), tab["n_chunks"],
                                              N.ptr(tab["dstate"]), N.ptr(tab["cfg"]), N.stream()))
            if not capturing:
                for p in params:
                    self.state[p]["step"] = self.state[p].get("step", 0) + 1
        return loss

    def load_state_dict(self, state_dict):
        """The loaded state tensors are new objects: forget the device tables (rebuilt, with the loaded step count, at the next
        ``step``)."""
        super().load_state_dict(state_dict)
        self._tables = {}

    def sync_step_counts(self):
        """After CUDA-graph replays (which advance the device counter only): copy it into the per-parameter ``step`` entries
        so that ``state_dict()`` is exact.  One device-to-host read."""
        for gi, group in enumerate(self.param_groups):
            tab = self._tables.get(gi)
            if tab is None:
                continue
            n = int(tab["dstate"][1].item())
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = n

    def last_grad_norm(self, group: int = 0) -> torch.Tensor:
        """Total gradient norm of the last step (device scalar, fp64) -- what clip_grad_norm_ returns."""
        return self._tables[group]["dstate"][5]

    def bytes_per_step(self) -> int:
        """Algorithmic HBM traffic of one step: 4 B (norm pass) + 16 B read + 12 B written per parameter."""
        return sum(t["n_params"] for t in self._tables.values()) * 32
