#!/usr/bin/env python
"""Benchmark of the RAD-MMM flow-decoder train step (BASELINE.json metric: mel-frames/s of a decoder train step).

    python bench.py --gpus N --steps K --warmup W            # ours, one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference itself on the host cores (same config)

A step is one pass of the hot path over one synthetic batch: RADMMMFlow.forward + flow NLL + backward
(+ the gradient all-reduce when N > 1).  Workload at N=1: BASELINE.json configs[1] -- the full-depth RADMMM decoder
(configs/RADMMM_model_config.yaml, 8 flows, 219 M parameters), B=8 utterances padded to T=800 frames with lengths in
[400, 800].  Weak scaling: every rank gets its own batch of the same shape.

The reference arm runs the UNMODIFIED reference modules (``/root/reference`` or the copy staged under ``oracle/_ref`` by
``python -m oracle.stage_ref``) -- ``decoders.RADMMMFlow`` + ``loss.compute_flow_loss`` + backward, fp32, all host
threads, on the SAME config; if neither is present it falls back to the oracle port and says so (``kind: "port"``).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "mel_frames_per_sec_decoder_train_step"
UNIT = "mel-frames/s"
FWD_MFLOP_PER_FRAME = 218.8      # BASELINE.md section 3 (8 flows + context LSTM)
TRAIN_MFLOP_PER_FRAME = 656.4    # 3 x forward (fwd + dgrad + wgrad; recompute not counted)
K5_FLOP_PER_GROUPED_FRAME = 2 * 1024 * 1024 * 5      # one dilated k=5 layer (the dominant kernel), per grouped frame
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at B=8 x T=800, bf16, from the
# committed `ncu --set full` capture named below (re-measured whenever the kernel changes)
K5_DRAM_BYTES_NCU = {"bytes": 17337088, "source": "profiles/r2_ncu_full_k5_final.md"}
MODEL_ARGS = dict(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2, n_mel_channels=80,
                  n_flows=8)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_batch(batch, frames, rank):
    from radmmm_b200 import synthetic as syn
    return syn.synthetic_batch(batch, frames, tag=f"bench.rank{rank}")


# ------------------------------------------------------------------------------------------------------- reference arm
def _reference_step_fn(batch, frames):
    """(step(), valid_frames, kind, description).  The unmodified reference when it can be imported, else the port."""
    from radmmm_b200 import synthetic as syn
    from oracle import ref_import
    bt = syn.synthetic_batch(batch, frames, tag="bench.rank0")          # the batch rank 0 of our arm trains on
    valid = int(bt["out_lens"].sum())
    sd = syn.synthetic_state_dict()
    if ref_import.available():
        root, mods = ref_import.import_reference()
        dec = mods["decoders"].RADMMMFlow(**MODEL_ARGS, n_conv_layers_per_step=4, n_early_size=2, n_early_every=2,
                                          affine_model="wavenet", scaling_fn="tanh", affine_activation="softplus",
                                          use_partial_padding=True)
        dec.load_state_dict(sd, strict=True)                          # flow 0's whitening layer arrives initialised
        dec.train()
        SL, cfl = mods["common"].SequenceLength, mods["loss"].compute_flow_loss

        def step():
            for p in dec.parameters():
                p.grad = None
            out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SL(bt["out_lens"]), f0=bt["f0"],
                      energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
            lens_g = bt["out_lens"] // 2
            mask = (torch.arange(frames // 2)[None] < lens_g[:, None])[:, None].float()
            n_el = torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor")                 # loss.py:520
            loss, _ = cfl(out["z_mel"], list(out["log_det_W_list"]), out["log_s_list"], n_el, out["z_mel"].size(1), mask, 1.0)
            loss.backward()
            return float(loss)
        where = "staged copy oracle/_ref" if root.endswith("_ref") else root
        return step, valid, "reference", f"unmodified reference decoders.RADMMMFlow + loss.compute_flow_loss + backward ({where})"
    from oracle import flow as of
    cfg = of.DecoderConfig.radmmm()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype == torch.float32 and not any(s in k for s in ("invtbl_conv.p", "lower_diag", "input_mean"))}
    sdp = dict(sd)
    sdp.update(params)
    lstm = of.build_context_lstm(sd, cfg)

    def step():
        for p in params.values():
            p.grad = None
        out = of.decoder_forward(sdp, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                                 bt["energy_avg"], bt["accent_vecs"], lstm=lstm)
        loss, _ = of.flow_loss(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2,
                               n_elements=torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor"))
        loss.backward()
        return float(loss)
    return step, valid, "port", "oracle port (oracle/_ref not staged on this box)"


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores, SAME config as our arm (B x T), fp32."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, frames = args.ref_batch or args.batch, args.ref_frames or args.frames
    step, valid, kind, what = _reference_step_fn(batch, frames)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = valid / dt
    sample = f"{what}; {batch} utterances x {frames} frames ({valid} valid frames), fp32, torch CPU, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RADMMM decoder train step (configs/RADMMM_model_config.yaml, 8 flows, 219M params), "
                                   f"B={batch} x T={frames}: forward + flow NLL + backward on the host CPU",
                       "batch_per_gpu": batch, "frames": frames, "valid_frames_per_gpu": valid, "precision": "fp32",
                       "same_config_as_gpu_arm": batch == args.batch and frames == args.frames},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(batch, frames, timeout_s=420.0):
    """Bounded CPU run of the reference on rank 0 (reported baseline, not the target): `bench.py --impl reference` on the
    SAME config in a fresh process, 1 warm-up + 2 timed steps (~20 s of CPU work at B=8 x T=800).  A separate process:
    the CPU run shares nothing with the CUDA process (thread pools, allocator, autograd device threads), and a hard
    timeout keeps the default run bounded."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                              "--batch", str(batch), "--frames", str(frames)],
                             capture_output=True, text=True, timeout=timeout_s, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["sample"] += "; 2 timed steps after 1 warm-up, separate process"
        cb["ms_per_step"] = line["ms_per_step"]
        return cb
    except Exception as exc:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference",
                "sample": f"CPU sample failed: {type(exc).__name__}: {str(exc)[:200]}"}


# ------------------------------------------------------------------------------------------------------- HBM-bound kernels
def hbm_rooflines(dev, pk, batch, frames):
    """Every HBM-bound kernel of the path timed ALONE through the C ABI at the benchmark's shapes: CUDA events around one
    launch, L2 flushed (a 512 MB memset) before every timed launch and the launch queued behind a device-side spin (a
    ~10 us kernel timed on an idle stream measures the host's launch latency instead), median of 7.  achieved = algorithmic bytes / time."""
    from radmmm_b200 import _native as N
    from radmmm_b200 import synthetic as syn
    lib = N.lib()
    B, Tp, C, D, H = batch, frames // 2, 160, 1056, 1024
    R = N.rows(B, Tp)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    st = N.stream()
    lens = torch.full((B,), Tp, dtype=torch.int32, device=dev)
    f32 = lambda *s: torch.randn(*s, device=dev)          # noqa: E731
    out = []

    def timed(name, nbytes, fn, note=""):
        ts = []
        for _ in range(7):
            torch.cuda._sleep(400_000)      # ~0.2 ms of device spin: the host queues the launch below while the GPU is
            flush.zero_()                   # still busy, so the events bracket the kernel and not the host's launch path
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "us": round(ms * 1e3, 2), "algorithmic_bytes": int(nbytes), "achieved": round(gbs, 1),
                    "unit": "GB/s", "peak": pk["hbm_gbs"], "frac": round(gbs / pk["hbm_gbs"], 3), "note": note})

    # weight preparation of one flow step (weight norm + both layouts), bf16
    d = N.FlowDesc()
    d.mode, d.B, d.C, d.Tp, d.D, d.H, d.L = N.MODE_BF16, 1, C, 1, D, H, 4
    sd = {k: v.to(dev) for k, v in syn.synthetic_state_dict(n_flows=1).items() if "coupling_tfn" in k}
    pre = "flows.0.coupling_tfn.affine_param_predictor."
    keep = []

    def P(name):
        t = sd[pre + name].contiguous()
        keep.append(t)
        return N.fptr(t)
    d.start_g, d.start_v, d.start_b = P("start.weight_g"), P("start.weight_v"), P("start.bias")
    for i in range(4):
        d.in_g[i], d.in_v[i], d.in_b[i] = P(f"in_layers.{i}.conv.weight_g"), P(f"in_layers.{i}.conv.weight_v"), P(f"in_layers.{i}.conv.bias")
        d.rs_g[i], d.rs_v[i], d.rs_b[i] = P(f"res_skip_layers.{i}.weight_g"), P(f"res_skip_layers.{i}.weight_v"), P(f"res_skip_layers.{i}.bias")
    d.end_w, d.end_b = P("end.weight"), P("end.bias")
    prepared = torch.zeros(lib.radmmm_flow_prepared_bytes(N.MODE_BF16, C, D, H, 4), dtype=torch.uint8, device=dev)
    d.prepared = N.ptr(prepared)
    n_w = sum(t.numel() for t in keep)
    timed("weight_prep (weight norm + K-major and transposed bf16 layouts, one flow step)", n_w * (4 + 2 * 2),
          lambda: N.check(lib.radmmm_flow_prepare(ctypes.byref(d), st)), "4 B read + 2 x 2 B written per weight")
    # 1x1 invertible conv and its weight gradient
    z, W, zo = f32(B, C, Tp), f32(C, C), torch.empty(B, C, Tp, device=dev)
    timed("inv1x1", 2 * z.numel() * 4, lambda: N.check(lib.radmmm_inv1x1(N.fptr(z), N.fptr(W), None, None, N.fptr(zo), B, C, C, Tp, st)))
    dW = torch.empty(C, C, device=dev)
    timed("inv1x1_wgrad", 2 * z.numel() * 4, lambda: N.check(lib.radmmm_inv1x1_wgrad(N.fptr(z), N.fptr(zo), None, N.ptr(lens), N.fptr(dW), B, C, Tp, st)))
    # coupling tail forward / backward
    prm, ls = f32(B, C, Tp) * 0.1, torch.empty(B, C // 2, Tp, device=dev)
    timed("coupling_fwd", int(3.5 * z.numel() * 4), lambda: N.check(lib.radmmm_coupling_forward(N.fptr(z), N.fptr(prm), N.fptr(zo), N.fptr(ls), B, C, Tp, 0, 0, st)))
    dz, dp = torch.empty_like(z), torch.empty_like(z)
    timed("coupling_bwd", int(5.5 * z.numel() * 4), lambda: N.check(lib.radmmm_coupling_backward(
        N.fptr(zo), N.fptr(ls), N.fptr(z), N.fptr(prm), N.ptr(lens), N.fptr(dz), N.fptr(dp), B, C, Tp, 0, st)))
    # flow-NLL reductions
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    timed("masked_sum (flow NLL)", z.numel() * 4, lambda: N.check(lib.radmmm_masked_sum(N.fptr(z), N.ptr(lens), B, C, Tp, 1, N.ptr(acc), st)))
    coef = torch.ones(1, device=dev)
    timed("masked_sum_bwd", 2 * z.numel() * 4, lambda: N.check(lib.radmmm_masked_sum_backward(N.fptr(z), N.ptr(lens), B, C, Tp, 1, N.fptr(coef), 1.0, N.fptr(dz), st)))
    # conditioning (B, Tp, D) fp32 -> bf16 rows
    ctx = f32(B, Tp, D)
    rows = torch.empty(lib.radmmm_context_rows_bytes(N.MODE_BF16, B, Tp, D), dtype=torch.uint8, device=dev)
    timed("rows_from_btd (context -> rows)", ctx.numel() * 4 + rows.numel(),
          lambda: N.check(lib.radmmm_context_rows(N.MODE_BF16, N.fptr(ctx), N.ptr(lens), B, Tp, D, N.ptr(rows), st)))
    # spline coupling (80 channels x 65 parameters per grouped frame)
    z1, q = f32(B, 80, Tp), f32(B, 80 * 65, Tp)
    z1o, lsp = torch.empty_like(z1), torch.empty(B, 1, Tp, device=dev)
    timed("spline_fwd", (q.numel() + 2 * z1.numel()) * 4, lambda: N.check(lib.radmmm_spline_forward(
        N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(z1o), N.fptr(lsp), B, 80, Tp, 32, -3.0, 3.0, 0, st)))
    dq, dz1 = torch.empty_like(q), torch.empty_like(z1)
    timed("spline_bwd", (2 * q.numel() + 3 * z1.numel()) * 4, lambda: N.check(lib.radmmm_spline_backward(
        N.fptr(z1), N.fptr(q), N.ptr(lens), N.fptr(z1o), N.fptr(lsp), N.fptr(dz1), N.fptr(dq), B, 80, Tp, 32, -3.0, 3.0, st)))
    # soft attention with the fused context matmul (T1 = frames, T2 = frames / 6 tokens)
    T1, T2, Ca, Dt = frames, max(4, frames // 6), 80, 520
    qq, kk, prior, txt = f32(B, Ca, T1), f32(B, Ca, T2), torch.rand(B, T1, T2, device=dev), f32(B, Dt, T2)
    attn, logp, cx = torch.empty(B, 1, T1, T2, device=dev), torch.empty(B, 1, T1, T2, device=dev), torch.empty(B, Dt, T1, device=dev)
    in_lens = torch.full((B,), T2, dtype=torch.int32, device=dev)
    att_bytes = B * ((Ca * (T1 + T2) + 3 * T1 * T2) + Dt * (T1 + T2)) * 4
    timed("soft_attention (+ context matmul)", att_bytes, lambda: N.check(lib.radmmm_soft_attention(
        N.fptr(qq), N.fptr(kk), N.fptr(prior), N.ptr(in_lens), N.fptr(attn), N.fptr(logp), N.fptr(txt), N.fptr(cx), B, Ca, T1, T2, Dt,
        0.0005, st)))
    dattn, dlogp, dcx = f32(B, 1, T1, T2), f32(B, 1, T1, T2), f32(B, Dt, T1)
    dq, dk, dtxt = torch.empty_like(qq), torch.empty_like(kk), torch.empty_like(txt)
    att_ws = torch.empty(lib.radmmm_soft_attention_backward_workspace_bytes(B, T1, T2), dtype=torch.uint8, device=dev)
    timed("soft_attention_backward (+ context matmul)", att_bytes + B * (2 * T1 * T2 + Dt * T1) * 4,
          lambda: N.check(lib.radmmm_soft_attention_backward(
              N.fptr(qq), N.fptr(kk), N.fptr(prior), N.ptr(in_lens), N.fptr(attn), N.fptr(dattn), N.fptr(dlogp), N.fptr(txt),
              N.fptr(dcx), N.fptr(dq), N.fptr(dk), N.fptr(dtxt), B, Ca, T1, T2, Dt, 0.0005, N.ptr(att_ws), att_ws.numel(), st)))
    return out


def frontend_bench(dev, pk, batch, frames):
    """STFT + mel front end (audio_processing.TacotronSTFT.mel_spectrogram): GPU kernel vs the oracle on the host cores."""
    from radmmm_b200 import audio_processing as AP
    from oracle import frontend as ofe
    S = frames * 256
    stft = AP.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(dev)
    y = (torch.rand(batch, S, device=dev) - 0.5).clamp(-1, 1)
    for _ in range(3):
        mel = stft.mel_spectrogram(y, check_range=False)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(7):
        torch.cuda._sleep(400_000)          # the launch is queued while the GPU still spins: the events time the kernel, not
        flush.zero_()                       # the host (the reference's two range asserts, each a device sync, are skipped)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mel = stft.mel_spectrogram(y, check_range=False)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    n_frames = batch * mel.shape[2]
    gbs = n_frames * 1344 / (ms * 1e-3) / 1e9
    yc = y[:2].cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    ofe.mel_spectrogram(yc)
    t0 = time.perf_counter()
    ofe.mel_spectrogram(yc)
    cpu_s = time.perf_counter() - t0
    return {"kernel": "stft_mel", "frames_per_s": n_frames / (ms * 1e-3), "audio_s_per_s": batch * S / 22050 / (ms * 1e-3),
            "us": ms * 1e3, "achieved": gbs, "unit": "GB/s", "peak": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"],
            "algorithmic_bytes_per_frame": 1344,
            "cpu_frames_per_s": 2 * mel.shape[2] / cpu_s, "cpu_note": f"oracle dense-DFT restatement of TacotronSTFT on {os.cpu_count()} host threads, 2 utterances"}


def alignment_bench(dev, batch, frames):
    """Hard alignment after `binarization_start_iter`: batched GPU monotonic alignment search and the batched
    attention CTC loss (forward + gradient) against the reference's own code on the same maps -- numba `mas_width1` looped
    over the batch with the device->host / host->device copies of TTSModel.binarize_attention, and loss.AttentionCTCLoss
    (per-utterance log_softmax + nn.CTCLoss) forward + backward on the same GPU."""
    from radmmm_b200.alignment import binarize_attention
    from radmmm_b200.loss import AttentionCTCLoss
    T1, T2 = frames, 120
    gen = torch.Generator().manual_seed(7)
    out_lens = torch.randint(frames // 2, frames + 1, (batch,), generator=gen)
    out_lens[0] = frames
    in_lens = torch.clamp((out_lens.float() * 0.15).long(), 8, T2)
    in_lens[0] = T2
    t1 = torch.arange(T1, dtype=torch.float32)[None, :, None]
    t2 = torch.arange(T2, dtype=torch.float32)[None, None, :]
    logits = -0.3 * (t2 - t1 * (in_lens[:, None, None].float() / out_lens[:, None, None].float())) ** 2 + torch.randn(batch, T1, T2, generator=gen)
    attn = torch.softmax(logits.masked_fill(t2 >= in_lens[:, None, None], -float("inf")), dim=2).unsqueeze(1).contiguous().to(dev)
    logprob = logits.unsqueeze(1).contiguous().to(dev)
    il, ol = in_lens.to(dev), out_lens.to(dev)

    def timed_us(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    ctc = AttentionCTCLoss()

    def ctc_step():
        lp = logprob.detach().requires_grad_(True)
        ctc(lp, il, ol).backward()
    res = {"shape": f"{batch} utterances, {T1} mel frames x {T2} text positions (ragged)",
           "mas_us": timed_us(lambda: binarize_attention(attn, il, ol)), "ctc_fwd_bwd_us": timed_us(ctc_step)}
    try:
        from oracle import ref_import
        ref_import.import_reference()
        from alignment import mas_width1 as ref_mas
        from loss import AttentionCTCLoss as RefCTC

        def ref_binarize():                                   # tts_lightning_modules.py:270-284
            a = attn.data.cpu().numpy()
            out = torch.zeros_like(attn)
            for b in range(batch):
                out[b, 0, :out_lens[b], :in_lens[b]] = torch.tensor(ref_mas(a[b, 0, :out_lens[b], :in_lens[b]]), device=dev)
            return out
        ref_binarize()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ref_binarize()
        torch.cuda.synchronize()
        res["reference_mas_us"] = (time.perf_counter() - t0) / 5 * 1e6
        rctc = RefCTC()

        def ref_ctc_step():
            lp = logprob.detach().requires_grad_(True)
            rctc(lp, in_lens, out_lens).backward()
        for _ in range(2):
            ref_ctc_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ref_ctc_step()
        torch.cuda.synchronize()
        res["reference_ctc_fwd_bwd_us"] = (time.perf_counter() - t0) / 5 * 1e6
        res["reference_note"] = "unmodified reference code (staged copy): numba MAS on the host with its copies, torch CTC loop on the same GPU; wall clock"
    except Exception as exc:
        res["reference_note"] = f"reference unavailable: {type(exc).__name__}: {exc}"[:200]
    return res


def reference_gpu_bench(dev, batch, frames):
    """Context for a user switching over: the UNMODIFIED reference decoder (staged copy) running the same train step on the
    SAME B200 with stock PyTorch kernels (cuDNN convolutions, cuDNN LSTM), fp32 and bf16 autocast.  Not the baseline the
    metric is defined against (that is the CPU arm) -- an extra line."""
    from oracle import ref_import
    from radmmm_b200 import synthetic as syn
    root, mods = ref_import.import_reference()
    bt = {k: v.to(dev) for k, v in syn.synthetic_batch(batch, frames, tag="bench.rank0").items()}
    valid = int(bt["out_lens"].sum())
    dec = mods["decoders"].RADMMMFlow(**MODEL_ARGS, n_conv_layers_per_step=4, n_early_size=2, n_early_every=2,
                                      affine_model="wavenet", scaling_fn="tanh", affine_activation="softplus",
                                      use_partial_padding=True)
    dec.load_state_dict(syn.synthetic_state_dict(), strict=True)
    dec = dec.to(dev).train()
    SL, cfl = mods["common"].SequenceLength, mods["loss"].compute_flow_loss
    lens_g = bt["out_lens"] // 2
    mask = (torch.arange(frames // 2, device=dev)[None] < lens_g[:, None])[:, None].float()
    n_el = torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor")
    res = {}
    for name, amp in (("fp32", False), ("bf16_autocast", True)):
        def step():
            for p in dec.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SL(bt["out_lens"]), f0=bt["f0"],
                          energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
            loss, _ = cfl(out["z_mel"].float(), [x.float() for x in out["log_det_W_list"]], [x.float() for x in out["log_s_list"]],
                          n_el, out["z_mel"].size(1), mask, 1.0)
            loss.backward()
        try:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                step()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / 5
            res[name] = {"ms_per_step": ms, "value": valid / (ms * 1e-3), "unit": UNIT}
        except Exception as exc:
            res[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    del dec
    torch.cuda.empty_cache()
    res["note"] = ("unmodified reference decoders.RADMMMFlow + compute_flow_loss + backward on this GPU, stock PyTorch kernels, "
                   f"same batch ({batch} x {frames}); eager, no optimizer")
    return res


def precision_errors(dev, precisions):
    """max-abs z error and relative loss error of each contraction mode against the oracle (fp32 CPU) on a small sample:
    the 8-flow decoder, 2 utterances x 96 frames (the shape of tests/golden/decoder_full.npz)."""
    from oracle import flow as of
    from radmmm_b200 import decoders, loss as L, synthetic as syn
    from radmmm_b200.common import SequenceLength
    sd = syn.synthetic_state_dict()
    bt = syn.synthetic_batch(2, 96, tag="bench.parity")
    cfg = of.DecoderConfig.radmmm()
    with torch.no_grad():
        ref = of.decoder_forward(sd, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"], bt["energy_avg"], bt["accent_vecs"])
        n_el = torch.div(bt["out_lens"].sum(), 2, rounding_mode="floor")
        rl, _ = of.flow_loss(ref["z_mel"], ref["log_det_W_list"], ref["log_s_list"], bt["out_lens"] // 2, n_elements=n_el)
    m = of.length_mask(bt["out_lens"] // 2, 48)[:, None]
    res = {}
    for prec in precisions:
        dec = decoders.RADMMMFlow(**MODEL_ARGS)
        dec.load_state_dict(sd)
        dec = dec.to(dev).set_precision(prec).eval()
        b = {k: v.to(dev) for k, v in bt.items()}
        with torch.no_grad():
            out = dec(b["mel"], b["spk_vecs"], b["context"], SequenceLength(b["out_lens"], 96), f0=b["f0"],
                      energy_avg=b["energy_avg"], accent_vecs=b["accent_vecs"])
            l = L.RADMMMFlowLoss(1.0, 2)(out, b["out_lens"])["loss_mel"][0]
        res[prec] = {"z_max_abs_err": float(((out["z_mel"].cpu() - ref["z_mel"]) * m).abs().max()),
                     "loss_rel_err": abs(float(l) - float(rl)) / abs(float(rl))}
        del dec
    return res


def optimizer_bench(dev, pk, build, resident, host, valid_frames, timed, args):
    """fused multi-tensor RAdam + global-norm clip (radmmm_b200.radam.RAdam, configs: lr 1e-3, weight decay 1e-6,
    gradient_clip_val 1.0).  optimizer_ms: the three launches alone on the decoder's 219 M parameters (CUDA events);
    full_step: decoder train step + clip + optimizer captured in ONE CUDA graph, the reference loop's whole per-step device work
    for this path."""
    from radmmm_b200.graphs import GraphedTrainStep
    from radmmm_b200.radam import RAdam
    dec = build("bf16")
    opt = RAdam(dec.parameters(), lr=1e-3, weight_decay=1e-6, max_grad_norm=1.0)
    g = GraphedTrainStep(dec, resident)                       # gradients live in the decoder's arena from here on
    g(resident)
    opt.step()                                                # builds the device tables
    torch.cuda.synchronize()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.step()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    opt_ms = sorted(ts)[len(ts) // 2]
    nbytes = opt.bytes_per_step()
    del g
    gfull = GraphedTrainStep(dec, resident, after_backward=opt.step)
    for _ in range(3):
        gfull(resident)
    ms_full = timed(lambda: gfull(resident), max(5, args.steps // 2))
    ms_full_e2e = timed(lambda: float(gfull(host).cpu()), max(3, args.steps // 4))
    out = {"optimizer_ms": opt_ms, "algorithmic_bytes": nbytes, "achieved": nbytes / (opt_ms * 1e-3) / 1e9, "unit": "GB/s",
           "peak": pk["hbm_gbs"], "frac": nbytes / (opt_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
           "full_step": {"ms_per_step": ms_full, "value": valid_frames / (ms_full * 1e-3), "e2e_ms_per_step": ms_full_e2e,
                         "e2e_value": valid_frames / (ms_full_e2e * 1e-3), "unit": UNIT,
                         "note": "train step + gradient-norm clip + RAdam update of all 219M parameters in one CUDA graph"},
           "note": "radmmm_b200.radam.RAdam: 3 launches (gradient sum of squares, scalar hyper step, fused update); 32 B per "
                   "parameter; L2 flushed before every timed call"}
    del gfull, dec, opt
    torch.cuda.empty_cache()
    return out


def joint_training_bench(dev, dec, resident, reducer, world, timed, valid_frames, frames, args):
    """BASELINE.json config 3: the decoder step JOINTLY with the text encoder and the four attribute predictors
    (configs/RADMMM_{f0,energy,vpred,duration}model_config.yaml: ConvLSTMLinearDAP, in_dim 520, 3 backbone layers, hidden 256,
    k=5, dropout 0.5, spectral-normed bi-LSTM), one backward pass, everything in ONE CUDA graph; on several GPUs the encoder /
    predictor gradients travel in their own bucket next to the decoder's nine.  Frame-level predictors (f0, energy, voiced) read
    the frame-level text encoding (the batch's `context`), the duration predictor the token-level encoder output."""
    import torch.distributed as dist
    from radmmm_b200.ddp import BucketedGradReducer
    from radmmm_b200.encoders import ConvLSTMLinearDAP, Encoder
    from radmmm_b200.graphs import GraphedTrainStep
    B = resident["mel"].shape[0]
    n_tok = max(8, frames // 6)
    mk = lambda: ConvLSTMLinearDAP(n_speaker_dim=16, in_dim=520, out_dim=1, reduction_factor=16, n_backbone_layers=3,   # noqa: E731
                                   n_hidden=256, kernel_size=5, p_dropout=0.5)
    aux = torch.nn.ModuleDict({"encoder": Encoder(3, 520, 5, lstm_norm_fn="spectral"), "f0": mk(), "energy": mk(), "voiced": mk(),
                               "duration": mk()}).to(dev).train()
    if world > 1:
        for p in aux.parameters():
            dist.broadcast(p.data, 0)
    gen = torch.Generator().manual_seed(11)
    txt_emb = torch.randn(B, 520, n_tok, generator=gen).to(dev)
    in_lens = torch.clamp(resident["out_lens"] // 6, min=8, max=n_tok)
    tgt = {k: torch.rand(B, 1, frames, generator=gen).to(dev) for k in ("f0", "energy", "voiced")}
    dur_tgt = (torch.rand(B, 1, n_tok, generator=gen) * 8 + 2).to(dev)
    spk = resident["spk_vecs"]

    def masked_mse(x_hat, x, lens):
        m = (torch.arange(x.shape[2], device=dev)[None, :] < lens[:, None])[:, None].float()
        return (((x_hat - x) * m) ** 2).sum() / m.sum().clamp(min=1)

    for m in aux.modules():
        if hasattr(m, "precision"):
            m.precision = "bf16"               # like the decoder: bf16 contractions, cluster-resident LSTM recurrence
    # four independent chains (three frame-level predictors; encoder -> duration predictor), each on its own stream: their
    # bi-LSTM recurrences are T sequential steps on 32 CTAs each, so they run next to each other and next to the decoder
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]

    def extra_loss(st):
        cur = torch.cuda.current_stream(dev)
        parts = []
        for i, k in enumerate(("f0", "energy", "voiced")):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]), torch.autocast("cuda", dtype=torch.bfloat16):
                r = aux[k](tgt[k], st["context"], spk, st["out_lens"])
                parts.append(masked_mse(r["x_hat"].float(), r["x"], st["out_lens"]))
        streams[3].wait_stream(cur)
        with torch.cuda.stream(streams[3]), torch.autocast("cuda", dtype=torch.bfloat16):
            text_enc = aux["encoder"](txt_emb, in_lens).transpose(1, 2)          # the encoder itself opts out of autocast, as the reference's does
            r = aux["duration"](dur_tgt, text_enc, spk, in_lens)
            parts.append(masked_mse(r["x_hat"].float(), r["x"], in_lens))

        def join():
            for s_ in streams:
                cur.wait_stream(s_)
        return parts, join

    aux_reducer = BucketedGradReducer(aux, bucket_key=lambda name: "aux") if world > 1 else None
    g = GraphedTrainStep(dec, resident, reducer=reducer, extra_loss=extra_loss, extra_params=list(aux.parameters()),
                         after_backward=(aux_reducer.finish if aux_reducer is not None else None))
    for _ in range(3):
        g(resident)
    ms = timed(lambda: g(resident), max(5, args.steps // 2))
    if os.environ.get("RADMMM_BENCH_JOINT_TRACE"):           # diagnostic: where the joint step's time goes (per stream, LSTM kernels)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            g(resident)
            torch.cuda.synchronize()
        path = os.environ["RADMMM_BENCH_JOINT_TRACE"]
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        t0 = min(e["ts"] for e in ev)
        busy = {}
        for e in ev:
            busy.setdefault(e["args"].get("stream"), [0.0, 1e18, 0.0])
            b_ = busy[e["args"].get("stream")]
            b_[0] += e["dur"]; b_[1] = min(b_[1], e["ts"] - t0); b_[2] = max(b_[2], e["ts"] + e["dur"] - t0)
        sys.stderr.write("joint trace: span %.0f us\n" % (max(e["ts"] + e["dur"] for e in ev) - t0))
        for k_, b_ in sorted(busy.items(), key=lambda kv: -kv[1][0])[:12]:
            sys.stderr.write("  stream %s busy %.0f us, first %.0f last %.0f\n" % (k_, b_[0], b_[1], b_[2]))
        for e in ev:
            if "lstm_cl" in e["name"] or e["dur"] > 300:
                sys.stderr.write("  %8.0f +%7.0f  s%s %s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:70]))
        os.remove(path)
    n_aux = sum(p.numel() for p in aux.parameters())
    out = {"ms_per_step": ms, "value": world * valid_frames / (ms * 1e-3), "unit": UNIT, "aux_parameters": n_aux,
           "aux_grad_nonzero": bool(sum(float(p.grad.abs().sum()) for p in aux.parameters() if p.grad is not None) > 0),
           "note": "decoder train step + text encoder (3 conv + bi-LSTM) + f0 / energy / voiced / duration predictors "
                   f"(radmmm_b200.encoders, {n_tok} tokens; predictor convolutions under bf16 autocast like the decoder's contractions, "
                   "bi-LSTMs on the cluster kernels), one backward, one CUDA graph"
                   + ("; encoder / predictor gradients all-reduced in a 10th bucket" if world > 1 else "")}
    del g
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from radmmm_b200 import _native as N
    from radmmm_b200 import decoders, loss as L, synthetic as syn
    from radmmm_b200.common import SequenceLength
    from radmmm_b200.graphs import GraphedInfer, GraphedTrainStep
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver reads the rank count from NCCL's INFO lines); the JSON line is
        # the LAST thing on stdout (printed after the final barrier, see _finish)
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    precision = args.precision
    batch, frames = args.batch, args.frames
    sd = syn.synthetic_state_dict()

    def build(prec):
        d = decoders.RADMMMFlow(**MODEL_ARGS)
        d.load_state_dict(sd)
        return d.to(dev).set_precision(prec).train()

    dec = build(precision)
    reducer = None
    if world > 1:
        from radmmm_b200.ddp import BucketedGradReducer
        reducer = BucketedGradReducer(dec).install()

    host = {k: v.pin_memory() for k, v in make_batch(batch, frames, rank).items()}
    valid_frames = int(host["out_lens"].sum())
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    crit = L.RADMMMFlowLoss(1.0, 2)

    def train_step(d, bt, red=None):
        """What an unmodified Lightning loop executes per step for this path (tts_lightning_modules.py:672-685 + backward)."""
        for p in d.parameters():
            p.grad = None
        out = d(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
                energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        loss = crit(out, bt["out_lens"])["loss_mel"][0]
        loss.backward()
        if red is not None:
            red.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps, tag=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        if tag:
            host_ms[tag] = (time.perf_counter() - t0) * 1e3 / steps      # host enqueue time per step (diagnostic)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- eager module path (what a Lightning loop calls unchanged): secondary number
    def e2e_eager_step():
        bt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return float(train_step(dec, bt, reducer).detach().cpu())

    for _ in range(args.warmup):
        train_step(dec, resident, reducer)
    ms_eager = timed(lambda: train_step(dec, resident, reducer), max(3, args.steps // 2), "train_eager")
    e2e_eager_step()
    ms_e2e_eager = timed(e2e_eager_step, max(3, args.steps // 4))

    # ---- headline path: the same step captured once into a CUDA graph (radmmm_b200.graphs.GraphedTrainStep), replayed
    gstep, graph_error = None, None
    if not args.eager:
        try:
            gstep = GraphedTrainStep(dec, resident, reducer=reducer)
        except Exception as exc:                                   # report, then fall back to the eager numbers
            graph_error = f"{type(exc).__name__}: {exc}"[:300]
            gstep = None
    graph_ok = torch.tensor([1 if gstep is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(graph_ok, op=dist.ReduceOp.MIN)
    if int(graph_ok) == 0:
        gstep = None
    run_step = (lambda: gstep(resident)) if gstep is not None else (lambda: train_step(dec, resident, reducer))

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        run_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = N.lib().radmmm_launch_count()
    ms = timed(run_step, args.steps, 'train')
    launches = N.lib().radmmm_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers in, loss out, through the public API (GraphedTrainStep.__call__ / the module)
    def e2e_step():
        if gstep is not None:
            loss = gstep()                   # consumes the batch the previous iteration prefetched, replays
            gstep.prefetch(host)             # THIS iteration's host->device copy (15 MB, pinned): it travels while the replay runs
            return float(loss.cpu())         # device->host read of the loss (a sync): every iteration has one copy in, one out
        return e2e_eager_step()

    if gstep is not None:
        gstep.prefetch(host)
    e2e_step()
    ms_e2e = timed(e2e_step, max(3, args.steps // 2))

    # ---- N > 1: the in-graph NCCL-averaged gradients against an explicit all_gather + mean of the local gradients
    allreduce_check = None
    if world > 1 and reducer is not None:
        run_step()
        torch.cuda.synchronize()
        got = {k: b["flat"].clone() for k, b in reducer.buckets.items()}
        w0, reducer.world = reducer.world, 1                       # local gradients only: no collective is issued
        train_step(dec, resident, reducer)
        reducer.world = w0
        worst, gmax, per_bucket, worst_param = 0.0, 0.0, {}, None
        for k, b in reducer.buckets.items():
            parts = [torch.empty_like(b["flat"]) for _ in range(world)]
            dist.all_gather(parts, b["flat"].contiguous())
            mean = torch.stack(parts).double().mean(0)
            diff = (got[k].double() - mean).abs()
            per_bucket[k] = float(diff.max())
            if per_bucket[k] > worst:
                worst = per_bucket[k]
                at = int(diff.argmax())
                for name, off, p in zip(b["names"], b["offsets"], b["params"]):
                    if off <= at < off + p.numel():
                        worst_param = name
            gmax = max(gmax, float(mean.abs().max()))
        # run-to-run noise of the local gradients themselves (fp32 atomics in the bias / 1x1 reductions): a second local step
        local1 = {k: b["flat"].clone() for k, b in reducer.buckets.items()}
        reducer.world = 1
        train_step(dec, resident, reducer)
        reducer.world = w0
        noise = max(float((local1[k].double() - b["flat"].double()).abs().max()) for k, b in reducer.buckets.items())
        allreduce_check = {"max_abs_diff": worst, "max_abs_grad": gmax, "buckets": len(reducer.buckets),
                           "per_bucket_max_abs_diff": per_bucket, "worst_parameter": worst_param,
                           "local_run_to_run_max_abs_diff": noise,
                           "bytes_per_step": reducer.bytes_per_step(),
                           "note": "gradients of the captured step (bucketed NCCL AVG inside the CUDA graph) vs all_gather + "
                                   "mean of the ranks' local gradients of the same step, fp64 compare"}

    # ---- inference (BASELINE metric "infer frames/s"): RADMMMFlow.infer on the same shapes, sigma = 0.8
    dec.eval()
    n_tok = max(4, frames // 6)
    lens_dev = resident["out_lens"]
    dur = torch.zeros(batch, n_tok, dtype=torch.long, device=dev)
    base = lens_dev // n_tok
    dur += base[:, None]
    dur[:, 0] += lens_dev - base * n_tok
    txt_enc = syn.hash_uniform("bench.txt", (batch, 520, n_tok)).to(dev)

    def infer_step():
        with torch.no_grad():
            return dec.infer(resident["spk_vecs"], txt_enc, 0.8, dur=dur, f0=resident["f0"],
                             energy_avg=resident["energy_avg"], out_lens=lens_dev)["mel"]

    for _ in range(3):
        infer_step()
    ms_infer_eager = timed(infer_step, max(5, args.steps))
    dec.train()

    # ---- per-kernel roofline of the dominant kernel (dilated k=5 conv forward, tcgen05): events around every launch
    lib = N.lib()
    l0 = lib.radmmm_launch_count()
    train_step(dec, resident, reducer)
    launches_per_eager_step = lib.radmmm_launch_count() - l0
    lib.radmmm_profile_enable(1)
    train_step(dec, resident, reducer)
    n_tags = 32
    cnt = (ctypes.c_int * n_tags)()
    kms = (ctypes.c_double * n_tags)()
    kfl = (ctypes.c_double * n_tags)()
    lib.radmmm_profile_collect(n_tags, cnt, kms, kfl)
    lib.radmmm_profile_enable(0)
    # ---- graphed inference: a failed capture must not be able to disturb the numbers above
    dec.eval()
    ms_infer, infer_graph_error = ms_infer_eager, None
    if not args.eager and world == 1:          # inference does not shard: replicas only, measured on one GPU
        try:
            ex = {"spk_vec": resident["spk_vecs"], "txt_enc": txt_enc, "dur": dur, "f0": resident["f0"],
                  "energy_avg": resident["energy_avg"], "out_lens": lens_dev}
            ginfer = GraphedInfer(dec, ex, sigma=0.8)
            for _ in range(3):
                ginfer(ex)
            ms_infer = timed(lambda: ginfer(ex), max(5, args.steps))
            del ginfer
        except Exception as exc:
            infer_graph_error = f"{type(exc).__name__}: {exc}"[:300]
    dec.train()
    names = {0: "start", 1: "k5_conv_fwd", 2: "res_skip_fwd", 3: "end", 4: "dgrad_end", 5: "dgrad_layer", 6: "dgrad_h0",
             7: "dgrad_z0", 8: "dgrad_ctx", 10: "lstm_projections", 17: "wgrad_1x1", 20: "wgrad_end", 21: "wgrad_k5",
             28: "wgrad_res_skip_grouped"}
    kernels = {names.get(i, f"tag{i}"): {"launches": cnt[i], "ms": round(kms[i], 4),
                                         "executed_tflops": round(kfl[i] / (kms[i] * 1e9), 1) if kms[i] > 0 else None}
               for i in range(n_tags) if cnt[i]}
    gemm_ms = sum(kms[i] for i in range(n_tags))
    pk = peaks()
    valid_grouped = int((host["out_lens"] // 2).sum())
    roof = None
    if cnt[1]:
        per_launch_ms = kms[1] / cnt[1]
        algo_flops = K5_FLOP_PER_GROUPED_FRAME * valid_grouped              # valid frames only (conservative)
        achieved = algo_flops / (per_launch_ms * 1e-3) / 1e12
        peak = pk["bf16_burst"] / (3.0 if precision == "bf16x3" else 1.0)
        same_shape = precision == "bf16" and batch == 8 and frames == 800
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel<EPI_IN> (dilated k=5 conv forward)", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": K5_DRAM_BYTES_NCU["bytes"] if same_shape else None,
                "traffic_note": f"dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                f"({K5_DRAM_BYTES_NCU['source']}); algorithmic bytes A 6.8 MB + W 10.5 MB + out 6.8 MB (the output "
                                "stays in the 126 MB L2)",
                "peak_source": pk["source"] + (", bf16 burst / 3 for the 3-pass split" if precision == "bf16x3" else ", bf16 burst"),
                "per_launch_ms": per_launch_ms, "algorithmic_flops_per_launch": algo_flops,
                "executed_tflops": kfl[1] / (kms[1] * 1e9),
                "timing": "CUDA events on the launching stream around every launch of one eager train step (radmmm_profile_enable)"}

    # ---- rank 0, N = 1 only: parity-grade mode beside the headline, measured errors, HBM-bound kernels, front end, CPU arm
    extras = {}
    if (world == 1 and not args.quick) or args.config3:
        # config 3.  On several GPUs only with --config3 (all ranks take part: the encoder / predictor gradients are a 10th
        # all-reduce bucket); the default multi-GPU line stays the plain decoder step the driver's scaling run expects.
        try:
            extras["joint_training"] = joint_training_bench(dev, dec, resident, reducer, world, timed, valid_frames, frames, args)
        except Exception as exc:
            extras["joint_training"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if world == 1 and not args.quick:
        if gstep is not None:
            del gstep
            gstep = None
        torch.cuda.empty_cache()
        try:
            dec3 = build("bf16x3")
            g3 = GraphedTrainStep(dec3, resident)
            for _ in range(3):
                g3(resident)
            ms3 = timed(lambda: g3(resident), max(5, args.steps // 2))
            ms3_e2e = timed(lambda: float(g3(host).cpu()), max(3, args.steps // 4))
            errs = precision_errors(dev, ["bf16x3", precision] if precision != "bf16x3" else ["bf16x3"])
            v3 = valid_frames / (ms3 * 1e-3)
            extras["parity_mode"] = {
                "precision": "bf16x3", "value": v3, "unit": UNIT, "ms_per_step": ms3, "e2e_value": valid_frames / (ms3_e2e * 1e-3),
                "e2e_ms_per_step": ms3_e2e,
                "step_tensor_roofline": {"achieved_tflops": v3 * TRAIN_MFLOP_PER_FRAME * 1e-6, "peak_tflops": pk["bf16_sustained"] / 3.0,
                                         "frac": v3 * TRAIN_MFLOP_PER_FRAME * 1e-6 / (pk["bf16_sustained"] / 3.0),
                                         "note": "3 tensor-core passes per product (hi*hi + lo*hi + hi*lo): peak = sustained bf16 / 3"},
                "errors_vs_oracle": errs,
                "note": "same step, same graph machinery, contractions in the fp32-grade split-bf16 mode that meets north_star's "
                        "tolerances (mel 1e-3 max-abs, log-det 1e-4 rel); errors_vs_oracle: 8-flow decoder, 2 x 96 frames, vs oracle fp32"}
            del g3, dec3
            torch.cuda.empty_cache()
        except Exception as exc:
            extras["parity_mode"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            extras["optimizer"] = optimizer_bench(dev, pk, build, resident, host, valid_frames, timed, args)
        except Exception as exc:
            extras["optimizer"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            extras["roofline_hbm"] = hbm_rooflines(dev, pk, batch, frames)
        except Exception as exc:
            extras["roofline_hbm"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            extras["frontend"] = frontend_bench(dev, pk, batch, frames)
        except Exception as exc:
            extras["frontend"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            extras["reference_same_gpu"] = reference_gpu_bench(dev, batch, frames)
        except Exception as exc:
            extras["reference_same_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            extras["hard_alignment"] = alignment_bench(dev, batch, frames)
        except Exception as exc:
            extras["hard_alignment"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        _finish(world)
        return
    value = world * valid_frames / (ms * 1e-3)
    e2e_value = world * valid_frames / (ms_e2e * 1e-3)
    step_tflops = value * TRAIN_MFLOP_PER_FRAME * 1e6 / 1e12 / world
    graphed = graph_error is None and not args.eager
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (hi/lo split, fp32-grade)", "fp32": "f32"}[precision], "data": "synthetic",
        "config": {"workload": "RADMMM decoder train step: weight norm of all 219M params + forward + flow NLL + backward "
                               f"(+ grad all-reduce), configs/RADMMM_model_config.yaml (8 flows), B={batch} x T={frames} per GPU, "
                               "replayed as one CUDA graph",
                   "batch_per_gpu": batch, "frames": frames, "valid_frames_per_gpu": valid_frames,
                   "padded_frames_per_gpu": batch * frames, "precision": precision,
                   "l2": "working set (0.9 GB prepared weights + ~1 GB activations per step) exceeds the 126 MB L2",
                   "parallelism": f"dp{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4,
                "note": "through GraphedTrainStep's public calls: every timed iteration copies one full batch from pinned host "
                        "memory (prefetch: on a copy stream, overlapping the replay in flight; consumed by the next iteration "
                        "through a device-to-device copy into the graph's static inputs) and reads the loss back to the host"
                        if graphed else "module API called eagerly; inputs copied from pinned host memory, loss read back, every step"},
        "infer": {"value": world * valid_frames / (ms_infer * 1e-3), "unit": UNIT, "ms_per_call": ms_infer,
                  "eager_ms_per_call": ms_infer_eager, "graph_error": infer_graph_error,
                  "note": "RADMMMFlow.infer (length regulation + context LSTM + 8 inverse flow steps), sigma 0.8, replayed "
                          "as a CUDA graph (radmmm_b200.graphs.GraphedInfer); eager_ms_per_call = the plain module call",
                  "tensor_roofline_frac": (world * valid_frames / (ms_infer * 1e-3)) * FWD_MFLOP_PER_FRAME * 1e6 / 1e12 / world / pk["bf16_sustained"]},
        "gpu_launches": int(launches) if not graphed else int(launches_per_eager_step) * args.steps,
        "graph": {"captured": graphed, "error": graph_error,
                  "note": "one whole train step (weight prep, LSTM, 8 flows, NLL, backward) replayed as a CUDA graph; "
                          "gpu_launches counts the kernels inside the graph x steps"},
        "eager": {"ms_per_step": ms_eager, "e2e_ms_per_step": ms_e2e_eager,
                  "value": world * valid_frames / (ms_eager * 1e-3), "e2e_value": world * valid_frames / (ms_e2e_eager * 1e-3),
                  "host_enqueue_ms_per_step": host_ms.get("train_eager"),
                  "note": "RADMMMFlow.forward + RADMMMFlowLoss + backward called eagerly: the path an UNMODIFIED Lightning loop "
                          "takes (class_path swap only); the headline needs GraphedTrainStep, i.e. a host-code edit"},
        "host_enqueue_ms_per_step": host_ms.get("train"),
        "roofline": roof,
        "step_tensor_roofline": {"achieved_tflops": step_tflops, "peak_tflops": pk["bf16_sustained"],
                                 "frac": step_tflops / pk["bf16_sustained"],
                                 "note": "valid frames x 656.4 MFLOP / step time vs sustained bf16 peak"},
        "contraction_kernels_one_step": kernels, "contraction_ms_one_step": gemm_ms,
    }
    if allreduce_check is not None:
        line["allreduce_check"] = allreduce_check
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline_sample(batch, frames)
    _finish(world, json.dumps(line))


def _finish(world, line=None):
    """Single process: print.  Multi-rank: every rank synchronises and meets at a last barrier, THEN rank 0 prints its JSON
    line (so it is the last line on stdout even with NCCL_DEBUG=INFO) and all leave without running the NCCL / graph
    destructors -- destroy_process_group() can block for minutes while CUDA graphs that captured NCCL collectives are alive."""
    if world <= 1:
        if line is not None:
            print(line)
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if line is not None:
        time.sleep(0.5)
        print(line)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RADMMM_BENCH_PRECISION", "bf16"), choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=800)
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm: utterances (default: --batch, the same config)")
    ap.add_argument("--ref-frames", type=int, default=0, help="reference arm: frames (default: --frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config3", action="store_true", help="also time the joint step with text encoder + attribute predictors (BASELINE config 3) at N > 1")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (no parity-mode / HBM-kernel / front-end / CPU sections)")
    ap.add_argument("--eager", action="store_true", help="skip the CUDA-graph capture; time the eager module path only")
    args = ap.parse_args()
    # safety net: a wedged collective / rendezvous must not hold a GPU box for its whole lease
    watchdog = threading.Timer(float(os.environ.get("RADMMM_BENCH_WATCHDOG_S", "1500")), lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
