#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 as markdown tables (for profiles/).

config 5  synthetic throughput sweep: train-step mel-frames/s (CUDA-graph replay, bf16) and inference frames/s over
          T in {128 .. 2048} x B in {1 .. 64} on one B200 -- each point is one `bench.py --quick` run, i.e. the same code and
          timing rules as the headline number; the reference CPU arm at a few small points for scale.
config 4  inference sampling path (16 kHz config = the same decoder dims, configs/RADMMM_16khz_model_config.yaml): `infer()`
          at B in {1, 8}, 100 text tokens with durations 2..10 (~600 frames), sigma sweep; reference `RADMMMFlow.infer` on
          the host cores next to it.

Usage: python tools/sweep_configs.py [--config5] [--config4] [--points small|full] > profiles/r2_sweep.md
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bench_point(batch, frames, extra=(), timeout=900):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--batch", str(batch), "--frames", str(frames),
           "--steps", "10", "--warmup", "3", *extra]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:                                    # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"[:200]}


def config5(points):
    if points == "full":
        Ts, Bs = [128, 256, 512, 1024, 2048], [1, 4, 16, 64]
    else:
        Ts, Bs = [128, 512, 2048], [1, 8, 64]
    print("## Config 5 -- synthetic throughput sweep, 1 x B200, bf16, CUDA-graph replay (bench.py --quick per point)\n")
    print("| T (frames) | B | valid frames | train ms/step | train frames/s | e2e frames/s | step tensor-roofline frac | infer ms | infer frames/s |")
    print("|---|---|---|---|---|---|---|---|---|")
    for T in Ts:
        for B in Bs:
            d = bench_point(B, T)
            if "error" in d:
                print(f"| {T} | {B} | - | error: {d['error']} | | | | | |")
                continue
            vf = d["config"].get("valid_frames_per_gpu", "")
            print(f"| {T} | {B} | {vf} | {d['ms_per_step']:.3f} | {d['value']:.0f} | {d['e2e']['value']:.0f} | "
                  f"{d.get('step_tensor_roofline', {}).get('frac', float('nan')):.3f} | {d['infer']['ms_per_call']:.3f} | {d['infer']['value']:.0f} |")
            sys.stdout.flush()
    print("\nReference CPU arm (`bench.py --impl reference`, unmodified reference on the host cores, fp32) at small points:\n")
    print("| T | B | ms/step | frames/s |\n|---|---|---|---|")
    for T, B in ((128, 1), (512, 8)):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", str(B), "--frames",
                                  str(T), "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"| {T} | {B} | {d['ms_per_step']:.1f} | {d['value']:.0f} |")
        except Exception as exc:                                # noqa: BLE001
            print(f"| {T} | {B} | error {type(exc).__name__} | |")
        sys.stdout.flush()


def _durations(syn, batch, n_tok):
    """Token durations 2..10 frames (config 4: ~600 frames from 100 tokens); every utterance's total is made even so the
    grouped (n_group_size = 2) frame count is exact."""
    dur = (2 + (syn.hash_uniform(f"c4.dur{batch}", (batch, n_tok), 0, 1) * 9).floor().clamp(max=8)).long()
    dur[:, 0] += dur.sum(1) % 2
    return dur


def config4():
    import torch
    from radmmm_b200 import decoders, synthetic as syn
    from radmmm_b200.graphs import GraphedInfer
    import bench as B

    dev = torch.device("cuda:0")
    dec = decoders.RADMMMFlow(**B.MODEL_ARGS)
    sd = syn.synthetic_state_dict()
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(dev).eval()
    dec.set_precision("bf16")
    print("\n## Config 4 -- inference sampling path, sigma sweep, 1 x B200 (RADMMMFlow.infer: length regulation + context LSTM + 8 inverse flows)\n")
    print("| B | text tokens | frames | sigma | precision | eager ms | graph ms | frames/s (graph) |")
    print("|---|---|---|---|---|---|---|---|")
    n_tok = 100
    results = {}
    for batch in (1, 8):
        dur = _durations(syn, batch, n_tok)
        out_lens = dur.sum(1)
        T = int(out_lens.max())
        ex = {"spk_vec": syn.hash_uniform("c4.spk", (batch, 16)).to(dev), "txt_enc": syn.hash_uniform("c4.txt", (batch, 520, n_tok)).to(dev),
              "dur": dur.to(dev), "f0": syn.hash_uniform("c4.f0", (batch, T), 0, 1).to(dev),
              "energy_avg": syn.hash_uniform("c4.en", (batch, T), 0, 1).to(dev), "out_lens": out_lens.to(dev)}
        valid = int(out_lens.sum())
        for prec in ("bf16", "bf16x3"):
            dec.set_precision(prec)
            for sigma in (0.0, 0.333, 0.667, 0.8, 1.0):
                def eager():
                    with torch.no_grad():
                        return dec.infer(ex["spk_vec"], ex["txt_enc"], sigma, dur=ex["dur"], f0=ex["f0"], energy_avg=ex["energy_avg"],
                                         out_lens=ex["out_lens"])["mel"]
                for _ in range(3):
                    eager()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    eager()
                e1.record()
                e1.synchronize()
                ms_e = e0.elapsed_time(e1) / 10
                try:
                    g = GraphedInfer(dec, ex, sigma=sigma)
                    for _ in range(3):
                        g(ex)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(20):
                        g(ex)
                    e1.record()
                    e1.synchronize()
                    ms_g = e0.elapsed_time(e1) / 20
                    del g
                except Exception as exc:                        # noqa: BLE001
                    ms_g = float("nan")
                    print(f"<!-- graph capture failed: {type(exc).__name__}: {exc} -->")
                results[(batch, prec, sigma)] = ms_g
                print(f"| {batch} | {n_tok} | {valid} | {sigma} | {prec} | {ms_e:.3f} | {ms_g:.3f} | {valid / (ms_g * 1e-3):.0f} |")
                sys.stdout.flush()
    # the reference's own infer on the host cores (B = 1 and 8, sigma 0.8)
    try:
        from oracle import ref_import
        root, mods = ref_import.import_reference()
        rdec = mods["decoders"].RADMMMFlow(**B.MODEL_ARGS, n_conv_layers_per_step=4, n_early_size=2, n_early_every=2,
                                           affine_model="wavenet", scaling_fn="tanh", affine_activation="softplus",
                                           use_partial_padding=True)
        rdec.load_state_dict(sd, strict=True)
        rdec.eval()
        torch.set_num_threads(os.cpu_count() or 1)
        # the reference draws its latent with torch.cuda.FloatTensor (decoders.py:221): on the host cores that one constructor
        # is aliased to the CPU type for the duration of the call (a shim in this tool; the reference file is untouched)
        cuda_ft = torch.cuda.FloatTensor
        print("\nReference `decoders.RADMMMFlow.infer` (unmodified, staged copy) on the host cores, fp32, sigma 0.8:\n")
        print("| B | frames | ms/call | frames/s | GPU graph speed-up (bf16) |\n|---|---|---|---|---|")
        for batch in (1, 8):
            dur = _durations(syn, batch, n_tok)
            out_lens = dur.sum(1)
            T = int(out_lens.max())
            spk, txt = syn.hash_uniform("c4.spk", (batch, 16)), syn.hash_uniform("c4.txt", (batch, 520, n_tok))
            f0, en = syn.hash_uniform("c4.f0", (batch, T), 0, 1), syn.hash_uniform("c4.en", (batch, T), 0, 1)
            acc = syn.hash_uniform("c4.acc", (batch, 8))

            def ref_call():
                with torch.no_grad():
                    return rdec.infer(spk, txt, 0.8, dur=dur, f0=f0[:, :int(out_lens.max())], energy_avg=en[:, :int(out_lens.max())],
                                      out_lens=out_lens, accent_vecs=acc)
            torch.cuda.FloatTensor = torch.FloatTensor
            try:
                ref_call()
                t0 = time.perf_counter()
                for _ in range(2):
                    ref_call()
                ms = (time.perf_counter() - t0) / 2 * 1e3
            finally:
                torch.cuda.FloatTensor = cuda_ft
            valid = int(out_lens.sum())
            print(f"| {batch} | {valid} | {ms:.1f} | {valid / (ms * 1e-3):.0f} | {ms / results[(batch, 'bf16', 0.8)]:.0f}x |")
            sys.stdout.flush()
        # and the unmodified reference on the SAME GPU (stock PyTorch kernels, fp32 and bf16 autocast)
        print("\nReference `decoders.RADMMMFlow.infer` (unmodified) on the same B200 with stock PyTorch kernels, sigma 0.8:\n")
        print("| B | frames | autocast | ms/call | frames/s | this package (graph, bf16) speed-up |\n|---|---|---|---|---|---|")
        gdec = rdec.to(dev)
        for batch in (1, 8):
            dur = _durations(syn, batch, n_tok)
            out_lens = dur.sum(1)
            T = int(out_lens.max())
            spk, txt = syn.hash_uniform("c4.spk", (batch, 16)).to(dev), syn.hash_uniform("c4.txt", (batch, 520, n_tok)).to(dev)
            f0, en = syn.hash_uniform("c4.f0", (batch, T), 0, 1).to(dev), syn.hash_uniform("c4.en", (batch, T), 0, 1).to(dev)
            acc, durd, old = syn.hash_uniform("c4.acc", (batch, 8)).to(dev), dur.to(dev), out_lens.to(dev)
            for amp in (False, True):
                def gcall():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                        return gdec.infer(spk, txt, 0.8, dur=durd, f0=f0, energy_avg=en, out_lens=old, accent_vecs=acc)
                try:
                    for _ in range(3):
                        gcall()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        gcall()
                    e1.record()
                    e1.synchronize()
                    ms = e0.elapsed_time(e1) / 10
                    valid = int(out_lens.sum())
                    print(f"| {batch} | {valid} | {'bf16' if amp else 'off (fp32)'} | {ms:.2f} | {valid / (ms * 1e-3):.0f} | {ms / results[(batch, 'bf16', 0.8)]:.1f}x |")
                except Exception as exc:                        # noqa: BLE001
                    print(f"| {batch} | - | {'bf16' if amp else 'off'} | error {type(exc).__name__}: {str(exc)[:80]} | | |")
                sys.stdout.flush()
    except Exception as exc:                                    # noqa: BLE001
        print(f"\nreference infer unavailable: {type(exc).__name__}: {exc}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config5", action="store_true")
    ap.add_argument("--config4", action="store_true")
    ap.add_argument("--points", default="full", choices=["small", "full"])
    a = ap.parse_args()
    print("# Round 2 -- BASELINE.json configs 4 and 5 on one B200 (tools/sweep_configs.py)\n")
    if a.config4 or not a.config5:
        config4()
    if a.config5 or not a.config4:
        config5(a.points)
