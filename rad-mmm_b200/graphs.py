"""CUDA-graph capture of the decoder train step.

The eager module path (radmmm_b200.decoders.RADMMMFlow called from a Lightning loop) issues ~1000 kernel launches per
step through Python / ctypes / autograd, and at B=8 x T=800 the host needs about as long to enqueue them as the B200
needs to run them.  ``GraphedTrainStep`` captures ONE whole step -- weight preparation, context LSTM, the 8 flow steps,
the flow NLL and the complete backward pass, including the side-stream forks -- into a CUDA graph with static input
buffers; every later step is ``copy inputs -> replay``.  Sequence lengths are device data, so one graph serves every
batch that is padded to the captured (batch, frames) shape.

Gradients land in ONE persistent arena per decoder (flat buckets of ``radmmm_b200.ddp.BucketedGradReducer``, allocated
outside any graph): the backward kernels write straight into it, ``p.grad`` is a view of it, and every graph captured
for that decoder -- ``GraphedTrainStepPool`` keeps one per frame bucket -- writes the same memory, so an optimizer
always reads the gradients of the step that ran last.  Do not set ``.grad`` to None between steps; zeroing is
unnecessary because a replay overwrites every gradient.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import loss as L
from .common import SequenceLength


def grad_arena(decoder, reducer=None):
    """The decoder's persistent gradient arena: ``reducer`` if given (multi-GPU: the caller's BucketedGradReducer, whose
    buckets are also the all-reduce buffers), else a private single-process BucketedGradReducer created once per decoder."""
    from .ddp import BucketedGradReducer
    if reducer is not None:
        decoder._radmmm_grad_arena = reducer
        return reducer
    arena = getattr(decoder, "_radmmm_grad_arena", None)
    if arena is None:
        arena = BucketedGradReducer(decoder, process_group=None, force_single=True).install()
        decoder._radmmm_grad_arena = arena
    return arena

_INPUT_KEYS = ("mel", "spk_vecs", "context", "out_lens", "f0", "energy_avg", "accent_vecs")


class GraphedTrainStep:
    """decoder forward + flow NLL + backward as one replayable CUDA graph.

    ``example`` is a dict with the keys of ``radmmm_b200.synthetic.synthetic_batch`` (mel (B,80,T), spk_vecs (B,16),
    context (B,n_text,T), out_lens (B), f0 (B,T), energy_avg (B,T), accent_vecs (B,n_acc)).  For a decoder whose
    whitening layer is already initialised only shapes and dtypes matter; for a FRESH decoder the warm-up steps run the
    reference's data-dependent initialisation (common.py:569-591) on ``example``, so pass a real first batch.
    ``reducer``: the ``BucketedGradReducer`` of a multi-GPU run (its ``finish`` -- the bucketed all-reduce -- is captured
    after ``backward``); single-process runs get a private gradient arena.  ``after_backward`` (optional) is called inside
    the captured region after that.  The loss follows ``RADMMMLoss.forward`` (loss.py:518-528):
    ``n_elements = floor(sum(out_lens) / n_group_size)``.  ``extra_loss(static_inputs)`` (optional) returns a scalar -- or
    ``(list of scalars, join)`` when it computes them on side streams; ``join()`` must make the current stream wait for
    those -- that is added to the flow loss before the single backward pass -- the text encoder and attribute predictors of the reference's
    joint training (tts_lightning_modules.py:643-686); pass their parameters as ``extra_params`` (and a ``reducer`` built over
    a module that contains them when running on several GPUs).  The returned loss is the flow loss alone.
    """

    def __init__(self, decoder, example: Dict[str, torch.Tensor], sigma: float = 1.0, warmup: int = 3,
                 after_backward: Optional[Callable[[], None]] = None, reducer=None,
                 extra_loss: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None, extra_params=()):
        dev = next(decoder.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("radmmm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.decoder, self.sigma, self.after_backward = decoder, sigma, after_backward
        self.extra_loss = extra_loss
        self.arena = grad_arena(decoder, reducer)
        self.params = [p for p in decoder.parameters() if p.requires_grad] + [p for p in extra_params if p.requires_grad]
        self.static = {k: example[k].detach().to(dev).clone() for k in _INPUT_KEYS if example.get(k) is not None}
        self.frames = int(self.static["mel"].shape[2])
        self.group = decoder.n_group_size
        # warm-up on a side stream (allocator pools, lazily created events / attributes, data-dependent init)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _step(self) -> torch.Tensor:
        st, dec = self.static, self.decoder
        # every captured / warm-up step starts without gradients, so autograd ADOPTS the arena views the backward
        # kernels wrote into (no copy, no accumulate) and ``p.grad`` ends up aliasing the persistent arena
        for p in self.params:
            p.grad = None
        extra_parts, extra_join = None, None
        if self.extra_loss is not None:
            # BEFORE the decoder: a callable that works on side streams (it returns (list of scalar losses, join)) then
            # overlaps the whole decoder forward, and its backward -- enqueued on the same side streams -- the decoder backward
            r = self.extra_loss(st)
            extra_parts, extra_join = r if isinstance(r, tuple) else ([r], None)
        out = dec(st["mel"], st["spk_vecs"], st["context"], SequenceLength(st["out_lens"], self.frames),
                  f0=st.get("f0"), energy_avg=st.get("energy_avg"), accent_vecs=st.get("accent_vecs"))
        lens_g = torch.div(st["out_lens"], self.group, rounding_mode="floor")
        loss, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], lens_g, self.sigma,
                             n_elements=L.n_elements_like_reference(st["out_lens"], self.group))
        total = loss
        if extra_parts is not None:            # joint training (config 3): encoder / predictor losses share the backward pass
            if extra_join is not None:
                extra_join()
            for part in extra_parts:
                total = total + part
        total.backward()
        self.arena.finish()                    # multi-GPU: the bucketed all-reduce; always: re-arm the arena
        if self.after_backward is not None:
            self.after_backward()
        return loss.detach()

    def prefetch(self, batch: Dict[str, torch.Tensor]) -> None:
        """Start moving the NEXT batch (pinned host tensors of the captured shapes) to the device on a copy stream, into staging
        buffers; the following ``step()`` (no argument) consumes it.  Called right after a replay has been enqueued, the
        transfer overlaps that replay -- what a data loader with ``pin_memory`` + ``non_blocking`` copies does for the eager
        loop.  The H2D copy of a batch costs 0.36 ms at B=8 x T=800 (15 MB over PCIe); without this it sits in front of
        every replay."""
        dev = self.static["mel"].device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(dev))
            self._has_staged = False
        cs = self._copy_stream
        cs.wait_event(self._staging_free)              # the previous step's copy out of the staging buffers is done
        with torch.cuda.stream(cs):
            for k, dst in self._staging.items():
                src = batch[k]
                if src.shape != dst.shape:
                    raise RuntimeError(f"GraphedTrainStep.prefetch: '{k}' has shape {tuple(src.shape)}, the graph was captured "
                                       f"for {tuple(dst.shape)} (pad the batch to the captured shape)")
                dst.copy_(src, non_blocking=True)
            self._staged_ready.record(cs)
        self._has_staged = True

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Copy ``batch`` (host or device tensors of the captured shapes) into the static buffers and replay; with no argument,
        consume the batch a previous ``prefetch`` staged.  Returns the static loss tensor (device; valid until the next
        call)."""
        if batch is None:
            if not getattr(self, "_has_staged", False):
                raise RuntimeError("GraphedTrainStep(): no batch given and none prefetched")
            cur = torch.cuda.current_stream(self.static["mel"].device)
            cur.wait_event(self._staged_ready)
            for k, dst in self.static.items():
                dst.copy_(self._staging[k], non_blocking=True)     # device-to-device, a few microseconds
            self._staging_free.record(cur)
            self._has_staged = False
            self.graph.replay()
            return self.loss
        for k, dst in self.static.items():
            src = batch[k]
            if src.shape != dst.shape:
                raise RuntimeError(f"GraphedTrainStep: '{k}' has shape {tuple(src.shape)}, the graph was captured for "
                                   f"{tuple(dst.shape)} (pad the batch to the captured shape)")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


_INFER_KEYS = ("spk_vec", "txt_enc", "dur", "f0", "energy_avg", "out_lens", "accent_vecs", "residual")


class GraphedInfer:
    """``RADMMMFlow.infer`` (length regulation, context LSTM, 8 inverse flow steps, fold) for one fixed
    (batch, tokens, max_frames) shape as a replayable CUDA graph.

    ``example``: spk_vec (B,16), txt_enc (B,n_text,T2), dur (B,T2) long, f0 / energy_avg (B,max_frames), out_lens (B),
    optional accent_vecs and ``residual`` (B, n_mel*g, max_frames//g: the latent sample; drawn inside the graph from the
    CUDA generator when absent).  Durations and lengths are device data: one graph serves every utterance batch padded
    to the captured shape.  Weights are taken as fixed (inference): the inverse 1x1 matrices are cached and the prepared
    WN weights are not re-derived inside the graph.
    """

    def __init__(self, decoder, example: Dict[str, torch.Tensor], sigma: float = 0.8, warmup: int = 2):
        dev = next(decoder.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("radmmm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.decoder, self.sigma = decoder, float(sigma)
        self.static = {k: example[k].detach().to(dev).clone() for k in _INFER_KEYS if example.get(k) is not None}
        self.max_frames = int(self.static["f0"].shape[1])
        decoder.enable_inverse_cache()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._call()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.mel = self._call()

    def _call(self) -> torch.Tensor:
        st = self.static
        with torch.no_grad():
            return self.decoder.infer(st["spk_vec"], st["txt_enc"], self.sigma, dur=st["dur"], f0=st.get("f0"),
                                      energy_avg=st.get("energy_avg"), out_lens=st["out_lens"],
                                      accent_vecs=st.get("accent_vecs"), residual=st.get("residual"),
                                      max_frames=self.max_frames)["mel"]

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Copy ``batch`` into the static buffers and replay; returns the static mel tensor (B, n_mel, max_frames)."""
        for k, dst in self.static.items():
            src = batch[k]
            if src.shape != dst.shape:
                raise RuntimeError(f"GraphedInfer: '{k}' has shape {tuple(src.shape)}, the graph was captured for "
                                   f"{tuple(dst.shape)}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.mel


_TIME_KEYS = {"mel": 2, "context": 2, "f0": 1, "energy_avg": 1}      # tensors with a time axis, and which axis it is


def pad_batch(batch: Dict[str, torch.Tensor], frames: int) -> Dict[str, torch.Tensor]:
    """Zero-pad the time axis of a decoder batch to ``frames`` (lengths stay as they are).  The decoder's outputs on the
    valid frames do not depend on the padding (everything beyond ``out_lens`` is masked), so a batch padded to a captured
    shape gives the same loss and gradients as the un-padded one."""
    out = dict(batch)
    for k, axis in _TIME_KEYS.items():
        x = batch.get(k)
        if x is None:
            continue
        t = x.shape[axis]
        if t > frames:
            raise ValueError(f"pad_batch: '{k}' has {t} frames, more than the target {frames}")
        if t < frames:
            out[k] = torch.nn.functional.pad(x, (0, frames - t))       # the time axis is the last one in all of them
    return out


class GraphedTrainStepPool:
    """One ``GraphedTrainStep`` per frame bucket, captured on first use: a loop with variable-length batches pads each
    batch up to the next bucket (e.g. 512 / 640 / 768 / 896 frames) and replays that bucket's graph.

    ``step_factory(decoder, example)`` builds the step object (default: GraphedTrainStep); the batch size is fixed by the
    first batch of each bucket.  All buckets share the decoder's gradient arena (see the module docstring), so
    ``optimizer.step()`` after any bucket's replay consumes that replay's gradients.
    """

    def __init__(self, decoder, frame_buckets, step_factory=None, **step_kwargs):
        self.decoder = decoder
        self.buckets = sorted(int(b) for b in frame_buckets)
        if not self.buckets:
            raise ValueError("GraphedTrainStepPool: at least one frame bucket is needed")
        self._factory = step_factory or (lambda dec, ex: GraphedTrainStep(dec, ex, **step_kwargs))
        self._steps: Dict[int, object] = {}

    def bucket_for(self, frames: int) -> int:
        for b in self.buckets:
            if frames <= b:
                return b
        raise ValueError(f"GraphedTrainStepPool: {frames} frames exceed the largest bucket ({self.buckets[-1]})")

    def __call__(self, batch: Dict[str, torch.Tensor]):
        b = self.bucket_for(int(batch["mel"].shape[2]))
        padded = pad_batch(batch, b)
        step = self._steps.get(b)
        if step is None:
            step = self._steps[b] = self._factory(self.decoder, padded)
        return step(padded)
