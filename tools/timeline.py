#!/usr/bin/env python
"""Kernel timeline of one train step (torch.profiler / CUPTI; no nsys in the image): start, duration, stream, name of
every kernel, plus per-stream busy time and the idle gaps on the main stream.  Diagnostic only (never a bench value)."""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=800)
    ap.add_argument("--out", default="gpurun_out/timeline")
    ap.add_argument("--infer", action="store_true")
    ap.add_argument("--graph", action="store_true", help="profile one replay of the CUDA-graphed step instead")
    args = ap.parse_args()
    from radmmm_b200 import decoders, loss as L, synthetic as syn
    from radmmm_b200.common import SequenceLength
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:      # torchrun: the step's bucketed NCCL all-reduce is part of the timeline; rank 0 reports
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    dec = decoders.RADMMMFlow(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2,
                              n_mel_channels=80, n_flows=8)
    dec.load_state_dict(syn.synthetic_state_dict())
    dec = dec.to(dev).set_precision(args.precision).train()
    bt = {k: v.to(dev) for k, v in syn.synthetic_batch(args.batch, args.frames, tag=f"bench.rank{rank}").items()}
    reducer = None
    if world > 1:
        from radmmm_b200.ddp import BucketedGradReducer
        reducer = BucketedGradReducer(dec).install()

    def step():
        for p in dec.parameters():
            p.grad = None
        out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], args.frames), f0=bt["f0"],
                  energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        loss, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2)
        loss.backward()
        if reducer is not None:
            reducer.finish()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if args.graph:
        from radmmm_b200.graphs import GraphedTrainStep
        gstep = GraphedTrainStep(dec, bt, reducer=reducer)
        for _ in range(3):
            gstep(bt)
        torch.cuda.synchronize()
        step = lambda: gstep(bt)      # noqa: E731
    if world > 1 and rank != 0:             # the other ranks just take part in the collectives of the profiled step
        import torch.distributed as dist
        step()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    path = args.out + ".trace.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ker = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ker.sort(key=lambda e: e["ts"])
    t0 = ker[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ker)
    with open(args.out + ".csv", "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in ker:
            f.write(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{e['args'].get('stream')},\"{e['name'][:120]}\"\n")
    busy = collections.defaultdict(float)
    for e in ker:
        busy[e["args"].get("stream")] += e["dur"]
    print(f"span {t1 - t0:.0f} us, {len(ker)} device activities")
    for s, b in sorted(busy.items(), key=lambda kv: -kv[1]):
        print(f"  stream {s}: busy {b:.0f} us")
    # union busy time across streams and the largest idle gaps
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ker)
    cur_s, cur_e = iv[0]
    union, gaps = 0.0, []
    for s, e in iv[1:]:
        if s > cur_e:
            union += cur_e - cur_s
            gaps.append((s - cur_e, cur_e - t0))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    union += cur_e - cur_s
    print(f"  any-stream busy {union:.0f} us, idle {t1 - t0 - union:.0f} us in {len(gaps)} gaps")
    for g, at in sorted(gaps, reverse=True)[:10]:
        print(f"    gap {g:.1f} us at {at:.0f} us")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ker:
        agg[e["name"][:90]][0] += 1
        agg[e["name"][:90]][1] += e["dur"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v[1]:9.1f} us {v[0]:5d}  {k}")
    # collectives: when each NCCL kernel starts / ends relative to the step, and what compute is still running by then
    nccl = [e for e in ker if "nccl" in e["name"].lower()]
    if nccl:
        comp_end = max(e["ts"] + e["dur"] for e in ker if "nccl" not in e["name"].lower())
        print(f"  NCCL kernels: {len(nccl)}; last compute kernel ends at {comp_end - t0:.0f} us, step ends at {t1 - t0:.0f} us "
              f"(exposed tail {t1 - comp_end:.0f} us)")
        for e in nccl:
            print(f"    nccl start {e['ts'] - t0:8.0f} us  dur {e['dur']:7.0f} us  end {e['ts'] + e['dur'] - t0:8.0f} us  stream {e['args'].get('stream')}  {e['name'][:60]}")
    os.remove(path)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
