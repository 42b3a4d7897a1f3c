"""Data-parallel gradient exchange for the flow decoder: one flat bucket per flow step, all-reduced over NCCL
(NVLink 5 / NVSwitch) as soon as that step's backward has produced its gradients.

Replaces Lightning's ``strategy: ddp`` for this path (configs/RADMMM_train_config.yaml:28).
The flow has no cross-sample operation, so the gradient sum is the ONLY exchange step: flows run backward 7 -> 0, so
bucket i's all-reduce overlaps the backward of flows i-1 .. 0.  FlowStepFunction.backward returns gradients that are
views into the bucket, and autograd's AccumulateGrad adopts them without a copy (``.grad`` is ``None`` before
backward), i.e. gradients are written straight into bucket storage.

Works with any ``torch.distributed`` backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist


def default_bucket_key(name: str) -> str:
    """Bucket assignment by parameter name: 'flows.<i>.coupling_tfn...' -> one bucket per flow step; everything else
    -> 'rest'.  The invertible-1x1-conv parameters (flows.<i>.invtbl_conv.*, 3 x 160 x 160 floats) go to 'rest' as well:
    their matrix is assembled once per step ahead of the flow chain (RADMMMFlow._prepare_weights_async), so their
    gradients only materialise at the very END of the backward pass -- in the flow's own bucket they would hold its
    all-reduce back until nothing is left to overlap it with."""
    parts = name.split(".")
    if len(parts) > 2 and parts[0] == "flows" and parts[1].isdigit() and parts[2] != "invtbl_conv":
        return "flow%s" % parts[1]
    return "rest"


class BucketedGradReducer:
    """All-reduce (mean) of parameter gradients in flat buckets, overlapped with backward.

    usage:
        reducer = BucketedGradReducer(model)        # after the model is on its device
        loss.backward(); reducer.finish()           # .grad of every parameter now holds the cross-rank mean
    """

    ALIGN = 64      # elements

    def __init__(self, module: torch.nn.Module, bucket_key: Callable[[str], str] = default_bucket_key,
                 process_group=None, average: bool = True, force_single: bool = False):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_initialized() and not force_single) else 1
        self.average = average
        self.buckets: Dict[str, dict] = {}
        self._handles: List = []
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            b = self.buckets.setdefault(bucket_key(name), {"params": [], "names": []})
            b["params"].append(p)
            b["names"].append(name)
        for key, b in self.buckets.items():
            # every slot starts on a 256-byte boundary: the kernels write gradients with 16-byte vector stores / atomics
            b["offsets"] = []
            n = 0
            for p in b["params"]:
                b["offsets"].append(n)
                n += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            p0 = b["params"][0]
            b["flat"] = torch.zeros(n, dtype=p0.dtype, device=p0.device)
            b["views"] = [b["flat"][off:off + p.numel()].view_as(p) for p, off in zip(b["params"], b["offsets"])]
            b["pending"] = len(b["params"])
            b["launched"] = False
            b["streams"] = {}
            for p, v in zip(b["params"], b["views"]):
                p.register_post_accumulate_grad_hook(self._make_hook(key, v))
        self.reset()

    # gradient buffers a custom backward can write into directly (FlowStepFunction does)
    def grad_view(self, param: torch.nn.Parameter) -> Optional[torch.Tensor]:
        return self._view_of.get(id(param))

    def fresh_view(self, param: torch.Tensor) -> Optional[torch.Tensor]:
        """A NEW tensor object viewing ``param``'s slot of its bucket (autograd adopts it as ``.grad`` without a copy
        because nothing else references the object).  Returns None -- the caller then writes into a private tensor and
        autograd accumulates normally -- when the parameter already holds a gradient (``zero_grad(set_to_none=False)``,
        gradient accumulation, a second backward): that gradient lives in the very slot a view would alias, and the
        kernels OVERWRITE their output."""
        slot = self._slot_of.get(param.data_ptr())
        if slot is None:
            return None
        flat, off, shape, owner = slot
        if owner.grad is not None:
            return None
        n = 1
        for d in shape:
            n *= d
        return flat[off:off + n].view(shape)

    def install(self):
        """Route FlowStepFunction's parameter gradients into the buckets."""
        from . import common
        common.set_grad_sink(self.fresh_view)
        return self

    def reset(self):
        self._view_of = {}
        self._slot_of = {}
        for b in self.buckets.values():
            b["pending"] = len(b["params"])
            b["launched"] = False
            b["streams"] = {}
            for p, v, off in zip(b["params"], b["views"], b["offsets"]):
                self._view_of[id(p)] = v
                self._slot_of[p.data_ptr()] = (b["flat"], off, tuple(p.shape), p)
        self._handles = []

    def _make_hook(self, key: str, view: torch.Tensor):
        def hook(param: torch.Tensor):
            b = self.buckets[key]
            if b["launched"]:          # a new backward began without finish() (single-process eager use): re-arm
                b["launched"] = False
                b["pending"] = len(b["params"])
                b["streams"] = {}
            g = param.grad
            if g.data_ptr() != view.data_ptr():      # gradient did not land in the bucket: copy it in and alias
                view.copy_(g)
                param.grad = view
            if g.is_cuda:
                # the stream this gradient was finalised on.  AccumulateGrad nodes keep the stream of their FIRST use (the
                # hooks keep them alive), so parameters of one bucket can be finalised on different streams; the collective
                # must wait for all of them, not only for the stream the last hook happens to run on
                cs = torch.cuda.current_stream(g.device)
                b["streams"][cs.cuda_stream] = cs
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)
        return hook

    def _join_streams(self, b: dict):
        """Make the current stream wait for every stream a gradient of this bucket was finalised on."""
        if not b["streams"]:
            return
        dev = b["flat"].device
        cur = torch.cuda.current_stream(dev)
        for key, s in b["streams"].items():
            if key != cur.cuda_stream:
                ev = torch.cuda.Event()
                ev.record(s)
                cur.wait_event(ev)
        b["streams"] = {}

    def _launch(self, b: dict):
        b["launched"] = True
        self._join_streams(b)
        if self.world > 1:
            op = dist.ReduceOp.SUM
            if self.average:
                if dist.get_backend(self.group) == "nccl":
                    op = dist.ReduceOp.AVG              # averaged inside the collective: no extra pass over the bucket
                else:
                    b["flat"].div_(self.world)
            self._handles.append(dist.all_reduce(b["flat"], op=op, group=self.group, async_op=True))

    def finish(self):
        """Wait for every outstanding all-reduce (call after ``backward``), then re-arm for the next step.  EVERY bucket
        is reduced every step: parameters that received no gradient contribute zeros (all ranks must issue the same
        collectives; buckets completed during backward go first, in backward order, the rest here in bucket order --
        ranks whose sets of unused parameters differ per step are not supported, as with DDP's static graph)."""
        for b in self.buckets.values():
            if not b["launched"]:
                for p, v in zip(b["params"], b["views"]):
                    if p.grad is None:
                        v.zero_()
                        p.grad = v
                self._launch(b)
        for h in self._handles:
            h.wait()
        self.reset()

    def bytes_per_step(self) -> int:
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets.values())
