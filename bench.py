#!/usr/bin/env python
"""Benchmark of the RAD-MMM flow-decoder train step (BASELINE.json metric: mel-frames/s of a decoder train step).

    python bench.py --gpus N --steps K --warmup W            # ours, one process per GPU (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (CPU oracle port) on host cores

A step is one pass of the hot path over one synthetic batch: RADMMMFlow.forward + flow NLL + backward
(+ the gradient all-reduce when N > 1).  Workload at N=1: BASELINE.json configs[1] -- the full-depth RADMMM decoder
(configs/RADMMM_model_config.yaml, 8 flows, 219 M parameters), B=8 utterances padded to T=800 frames with lengths in
[400, 800].  Weak scaling: every rank gets its own batch of the same shape.
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
K5_DRAM_BYTES_NCU = 17337344                         # measured once with ncu --set full (see roofline.traffic_note)


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
def run_reference(args):
    """The reference algorithm on the host CPU: the oracle port (torch-CPU restatement pinned to the unmodified
    reference by tests/golden; the Python reference itself cannot travel to the GPU box).  A bounded sample of the same
    workload: same model, fewer / shorter utterances."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import flow as of
    from radmmm_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, frames = args.ref_batch, args.ref_frames
    cfg = of.DecoderConfig.radmmm()
    sd = syn.synthetic_state_dict()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype == torch.float32 and not any(s in k for s in ("invtbl_conv.p", "lower_diag", "input_mean"))}
    sdp = dict(sd)
    sdp.update(params)
    lstm = of.build_context_lstm(sd, cfg)
    bt = syn.synthetic_batch(batch, frames, tag="bench.ref")
    valid = int(bt["out_lens"].sum())

    def step():
        for p in params.values():
            p.grad = None
        out = of.decoder_forward(sdp, cfg, bt["mel"], bt["spk_vecs"], bt["context"], bt["out_lens"], bt["f0"],
                                 bt["energy_avg"], bt["accent_vecs"], lstm=lstm)
        loss, _ = of.flow_loss(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2)
        loss.backward()
        return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = valid / dt
    sample = f"oracle port, {batch} utterances x {frames} frames (lengths {bt['out_lens'].tolist()}), fp32, torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RADMMM decoder train step (configs/RADMMM_model_config.yaml, 8 flows, 219M params); "
                                   "bounded CPU sample", "batch": batch, "frames": frames, "valid_frames": valid},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(timeout_s=240.0):
    """Bounded CPU run of the oracle port on rank 0 (reported baseline, not the target).  Runs in a FRESH process
    (`bench.py --impl reference`, two timed steps of 2 x 256 frames): the CPU port shares nothing with the CUDA process
    (thread pools, allocator, autograd device threads), and a hard timeout keeps the default run bounded."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                             capture_output=True, text=True, timeout=timeout_s, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["sample"] += ", 2 timed steps after 1 warm-up, separate process"
        return cb
    except Exception as exc:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                "sample": f"CPU sample failed: {type(exc).__name__}: {str(exc)[:200]}"}


# ------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from radmmm_b200 import _native as N
    from radmmm_b200 import decoders, loss as L, synthetic as syn
    from radmmm_b200.common import SequenceLength
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the single JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False          # keep the (cuDNN) context LSTM in fp32 like the reference default
    precision = args.precision
    batch, frames = args.batch, args.frames

    dec = decoders.RADMMMFlow(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2,
                              n_mel_channels=80, n_flows=8)
    dec.load_state_dict(syn.synthetic_state_dict())
    dec = dec.to(dev).set_precision(precision).train()
    reducer = None
    if world > 1:
        from radmmm_b200.ddp import BucketedGradReducer
        reducer = BucketedGradReducer(dec).install()

    host = {k: v.pin_memory() for k, v in make_batch(batch, frames, rank).items()}
    valid_frames = int(host["out_lens"].sum())
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}

    def train_step(bt):
        # no optimizer in the timed loop: mark the parameters as changed so that every step re-runs the weight norm +
        # re-layout of all 219 M parameters, as it must after a real optimizer step (nothing is served from a cache)
        dec.invalidate_weight_cache()
        for p in dec.parameters():
            p.grad = None
        out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
                  energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        loss, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2)
        loss.backward()
        if reducer is not None:
            reducer.finish()
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

    # ---- eager module path (what a Lightning loop calls): kept as a secondary number -- at this shape the host needs
    #      about as long to enqueue the ~1000 launches of a step as the GPU needs to run them
    def e2e_eager_step():
        bt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return float(train_step(bt).detach().cpu())

    for _ in range(args.warmup):
        train_step(resident)
    eager_steps = max(3, args.steps // 2)
    ms_eager = timed(lambda: train_step(resident), eager_steps, "train_eager")
    e2e_eager_step()
    ms_e2e_eager = timed(e2e_eager_step, max(3, args.steps // 4))

    # ---- headline path: the same step captured once into a CUDA graph (radmmm_b200.graphs.GraphedTrainStep) and
    #      replayed; inputs are copied into the graph's static buffers every step
    gstep, graph_error = None, None
    if not args.eager:
        try:
            from radmmm_b200.graphs import GraphedTrainStep
            gstep = GraphedTrainStep(dec, resident, after_backward=(reducer.finish if reducer is not None else None))
        except Exception as exc:                                   # report, then fall back to the eager numbers
            graph_error = f"{type(exc).__name__}: {exc}"[:300]
            gstep = None
    graph_ok = torch.tensor([1 if gstep is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(graph_ok, op=dist.ReduceOp.MIN)
    if int(graph_ok) == 0:
        gstep = None
    run_step = (lambda: gstep(resident)) if gstep is not None else (lambda: train_step(resident))

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
            return float(gstep(host).cpu())
        return e2e_eager_step()

    e2e_step()
    ms_e2e = timed(e2e_step, max(3, args.steps // 2))
    if gstep is not None:
        for p in dec.parameters():        # leave graph mode: the remaining (eager) sections own their gradients again
            p.grad = None

    # ---- inference (BASELINE metric "infer frames/s"): RADMMMFlow.infer on the same shapes (text tokens ~ T/6,
    #      durations summing to each length), sigma = 0.8
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
    train_step(resident)
    launches_per_eager_step = lib.radmmm_launch_count() - l0
    lib.radmmm_profile_enable(1)
    train_step(resident)
    n_tags = 32
    cnt = (ctypes.c_int * n_tags)()
    kms = (ctypes.c_double * n_tags)()
    kfl = (ctypes.c_double * n_tags)()
    lib.radmmm_profile_collect(n_tags, cnt, kms, kfl)
    lib.radmmm_profile_enable(0)
    # ---- graphed inference LAST: a failed capture must not be able to disturb any other number
    dec.eval()
    ms_infer, infer_graph_error = ms_infer_eager, None
    if not args.eager and world == 1:          # inference does not shard: replicas only, measured on one GPU
        try:
            from radmmm_b200.graphs import GraphedInfer
            ex = {"spk_vec": resident["spk_vecs"], "txt_enc": txt_enc, "dur": dur, "f0": resident["f0"],
                  "energy_avg": resident["energy_avg"], "out_lens": lens_dev}
            ginfer = GraphedInfer(dec, ex, sigma=0.8)
            for _ in range(3):
                ginfer(ex)
            ms_infer = timed(lambda: ginfer(ex), max(5, args.steps))
        except Exception as exc:
            infer_graph_error = f"{type(exc).__name__}: {exc}"[:300]
    dec.train()
    names = {0: "start", 1: "k5_conv_fwd", 2: "res_skip_fwd", 3: "end", 4: "dgrad_end", 5: "dgrad_layer", 6: "dgrad_h0",
             7: "dgrad_z0", 8: "dgrad_ctx", 17: "wgrad_1x1", 21: "wgrad_k5"}
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
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel<EPI_IN> (dilated k=5 conv forward)", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": K5_DRAM_BYTES_NCU if (precision == "bf16" and batch == 8 and frames == 800) else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                "(profiles/r1_ncu_full_k5_s12.md); algorithmic bytes A 6.8 MB + W 10.5 MB + out 6.8 MB",
                "peak_source": pk["source"] + (", bf16 burst / 3 for the 3-pass split" if precision == "bf16x3" else ", bf16 burst"),
                "per_launch_ms": per_launch_ms, "algorithmic_flops_per_launch": algo_flops,
                "executed_tflops": kfl[1] / (kms[1] * 1e9)}

    if rank != 0:
        _finish(world)
        return
    value = world * valid_frames / (ms * 1e-3)
    e2e_value = world * valid_frames / (ms_e2e * 1e-3)
    step_tflops = value * TRAIN_MFLOP_PER_FRAME * 1e6 / 1e12 / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (hi/lo split, fp32-grade)", "fp32": "f32"}[precision], "data": "synthetic",
        "config": {"workload": "RADMMM decoder train step: weight norm of all 219M params + forward + flow NLL + backward "
                               "(+ grad all-reduce), configs/RADMMM_model_config.yaml (8 flows), B=8 x T=800 per GPU, "
                               "replayed as one CUDA graph",
                   "batch_per_gpu": batch, "frames": frames, "valid_frames_per_gpu": valid_frames,
                   "padded_frames_per_gpu": batch * frames, "precision": precision,
                   "l2": "working set (0.9 GB prepared weights + ~1 GB activations per step) exceeds the 126 MB L2",
                   "parallelism": f"dp{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4},
        "infer": {"value": world * valid_frames / (ms_infer * 1e-3), "unit": UNIT, "ms_per_call": ms_infer,
                  "eager_ms_per_call": ms_infer_eager, "graph_error": infer_graph_error,
                  "note": "RADMMMFlow.infer (length regulation + context LSTM + 8 inverse flow steps), sigma 0.8, replayed "
                          "as a CUDA graph (radmmm_b200.graphs.GraphedInfer); eager_ms_per_call = the plain module call",
                  "tensor_roofline_frac": (world * valid_frames / (ms_infer * 1e-3)) * FWD_MFLOP_PER_FRAME * 1e6 / 1e12 / world / pk["bf16_sustained"]},
        "gpu_launches": int(launches) if gstep is None else int(launches_per_eager_step) * args.steps,
        "graph": {"captured": gstep is not None, "error": graph_error,
                  "note": "one whole train step (weight prep, LSTM, 8 flows, NLL, backward) replayed as a CUDA graph; "
                          "gpu_launches counts the kernels inside the graph x steps"},
        "eager": {"ms_per_step": ms_eager, "e2e_ms_per_step": ms_e2e_eager,
                  "value": world * valid_frames / (ms_eager * 1e-3), "e2e_value": world * valid_frames / (ms_e2e_eager * 1e-3),
                  "host_enqueue_ms_per_step": host_ms.get("train_eager"),
                  "note": "RADMMMFlow.forward + flow_nll + backward called eagerly (the Lightning-loop path)"},
        "host_enqueue_ms_per_step": host_ms.get("train"),
        "roofline": roof,
        "step_tensor_roofline": {"achieved_tflops": step_tflops, "peak_tflops": pk["bf16_sustained"],
                                 "frac": step_tflops / pk["bf16_sustained"],
                                 "note": "valid frames x 656.4 MFLOP / step time vs sustained bf16 peak"},
        "contraction_kernels_one_step": kernels, "contraction_ms_one_step": gemm_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(line))
    _finish(world)


def _finish(world):
    """Multi-rank teardown.  destroy_process_group() can block for minutes when CUDA graphs that captured NCCL
    collectives are still alive, so every rank synchronises, meets at a last barrier and leaves without running the
    NCCL / graph destructors."""
    if world <= 1:
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
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
    ap.add_argument("--ref-batch", type=int, default=2)
    ap.add_argument("--ref-frames", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="skip the CUDA-graph capture; time the eager module path only")
    args = ap.parse_args()
    # safety net: a wedged collective / rendezvous must not hold a GPU box for its whole lease
    watchdog = threading.Timer(float(os.environ.get("RADMMM_BENCH_WATCHDOG_S", "1200")), lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
