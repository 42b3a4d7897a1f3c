"""GraphedTrainStep (CUDA-graph replay of the whole decoder train step) must reproduce the eager module path bit for
bit on the same inputs, follow new inputs / new lengths written into its static buffers, and see parameter updates."""
import pytest
import torch

from radmmm_b200 import synthetic as syn

from .gpu_util import DEV, close

pytestmark = pytest.mark.gpu


def _decoder(precision):
    from radmmm_b200 import decoders
    dec = decoders.RADMMMFlow(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=2)
    dec.load_state_dict(syn.synthetic_state_dict(n_flows=2))
    return dec.to(DEV).set_precision(precision).train()


def _eager(dec, bt, frames):
    from radmmm_b200 import loss as L
    from radmmm_b200.common import SequenceLength
    for p in dec.parameters():
        p.grad = None
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    loss, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2,
                         n_elements=L.n_elements_like_reference(bt["out_lens"], 2))
    loss.backward()
    return loss.detach().clone(), {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_matches_eager(precision):
    from radmmm_b200.graphs import GraphedTrainStep
    batch, frames = 3, 96
    dec = _decoder(precision)
    a = {k: v.to(DEV) for k, v in syn.synthetic_batch(batch, frames, tag="graph.a").items()}
    b = {k: v.to(DEV) for k, v in syn.synthetic_batch(batch, frames, tag="graph.b").items()}
    b["out_lens"] = torch.tensor([frames, 70, 50], device=DEV)
    _eager(dec, a, frames)                          # data-dependent init of flow 0 happens here, once
    ref_a = _eager(dec, a, frames)
    ref_b = _eager(dec, b, frames)
    step = GraphedTrainStep(dec, a)
    for ref, bt in ((ref_a, a), (ref_b, b), (ref_a, a)):
        loss = step(bt)
        torch.cuda.synchronize()
        # atomics in the split-K weight-gradient reductions make the last bits order-dependent
        close(loss, ref[0], 1e-6, rtol=1e-6, what="graph loss")
        bad = []
        for n, p in dec.named_parameters():
            if n in ref[1]:
                g = ref[1][n]
                e = (p.grad - g).abs().max().item()
                if e > 1e-4 + 2e-4 * g.abs().max().item():
                    bad.append(f"{n}: err {e:.3e} (max |g| {g.abs().max().item():.3e})")
        assert not bad, "graph gradients differ from eager:\n" + "\n".join(bad)
    # a parameter update between replays must be seen (weight preparation is inside the graph)
    with torch.no_grad():
        for p in dec.parameters():
            p.add_(0.01 * torch.sign(p))
    loss_new = step(a).clone()
    torch.cuda.synchronize()
    grads_graph = {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}
    ref_new = _eager(dec, a, frames)
    assert abs(float(loss_new) - float(ref_a[0])) > 1e-4
    close(loss_new, ref_new[0], 1e-6, rtol=1e-6, what="graph loss after update")
    for n, g in ref_new[1].items():
        close(grads_graph[n], g, 1e-4, rtol=2e-4, what="graph grad after update " + n)


def test_graphed_infer_matches_eager():
    """GraphedInfer replays RADMMMFlow.infer; with an injected latent sample it must reproduce the eager call exactly,
    and follow new durations / lengths written into its static buffers."""
    from radmmm_b200.graphs import GraphedInfer
    dec = _decoder("bf16x3").eval()
    B, T2, T = 2, 8, 48

    def example(tag, lens):
        dur = torch.zeros(B, T2, dtype=torch.long)
        for b, n in enumerate(lens):
            dur[b] = n // T2
            dur[b, 0] += n - (n // T2) * T2
        ex = {"spk_vec": syn.hash_uniform(tag + ".spk", (B, 16), -1, 1), "txt_enc": syn.hash_uniform(tag + ".txt", (B, 520, T2), -1, 1),
              "dur": dur, "f0": syn.hash_uniform(tag + ".f0", (B, T), 4.4, 6.4), "energy_avg": syn.hash_uniform(tag + ".en", (B, T), 0.5, 1.0),
              "out_lens": torch.tensor(lens), "residual": syn.hash_uniform(tag + ".res", (B, 160, T // 2), -1, 1) * 0.7}
        return {k: v.to(DEV) for k, v in ex.items()}

    def eager(ex):
        with torch.no_grad():
            return dec.infer(ex["spk_vec"], ex["txt_enc"], 0.7, dur=ex["dur"], f0=ex["f0"], energy_avg=ex["energy_avg"],
                             out_lens=ex["out_lens"], residual=ex["residual"], max_frames=T)["mel"].clone()

    a, b = example("ginf.a", [48, 40]), example("ginf.b", [36, 48])
    ref_a, ref_b = eager(a), eager(b)
    g = GraphedInfer(dec, a, sigma=0.7)
    for ref, ex in ((ref_a, a), (ref_b, b), (ref_a, a)):
        mel = g(ex)
        torch.cuda.synchronize()
        assert mel.shape == ref.shape
        for i, n in enumerate(ex["out_lens"].tolist()):
            close(mel[i, :, :n // 2 * 2], ref[i, :, :n // 2 * 2], 1e-6, what="graphed infer mel")


@pytest.mark.gpu
def test_reducer_collective_waits_for_every_finalising_stream():
    """Regression (found by bench.py's allreduce_check at N=2): parameters of one bucket can be finalised on different
    streams (AccumulateGrad nodes keep the stream of their first use); the bucket's collective must wait for all of them, not
    only for the stream the last hook runs on.  Here the weight's gradient lands in the bucket on a slow stream s1, the bias'
    on s2; whatever s2 does after the bucket is launched must see the weight's gradient."""
    from radmmm_b200.ddp import BucketedGradReducer
    lin = torch.nn.Linear(512, 512).to("cuda")
    red = BucketedGradReducer(lin, force_single=True)
    b = red.buckets["rest"]
    hooks = {n: list(p._post_accumulate_grad_hooks.values())[0] for n, p in lin.named_parameters()}
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        torch.cuda._sleep(40_000_000)                      # ~20 ms: s1 is far behind
        lin.weight.grad = torch.full_like(lin.weight, 3.0)
        hooks["weight"](lin.weight)                        # copies into the bucket ON s1
    with torch.cuda.stream(s2):
        lin.bias.grad = torch.full_like(lin.bias, 5.0)
        hooks["bias"](lin.bias)                            # last parameter: launches the bucket from s2
        seen = b["flat"].clone()                           # stands for the collective: enqueued on s2 right after the launch
    torch.cuda.synchronize()
    assert b["launched"] and not b["streams"]
    assert float(seen[:lin.weight.numel()].min()) == 3.0, "the bucket was read before the weight gradient had landed"
    assert float(seen[b["offsets"][1]:b["offsets"][1] + lin.bias.numel()].min()) == 5.0
    red.finish()


def test_graphed_step_prefetch_pipeline():
    """GraphedTrainStep.prefetch + step(): batches staged from PINNED HOST memory on the copy stream while the previous replay
    runs; every replay must see exactly the batch prefetched for it (losses equal the direct path, in order), and calling
    step() with nothing staged is an error."""
    from radmmm_b200.graphs import GraphedTrainStep
    batch, frames = 3, 96
    dec = _decoder("fp32")
    host = [{k: v.pin_memory() for k, v in syn.synthetic_batch(batch, frames, tag=f"prefetch.{i}").items()} for i in range(3)]
    _eager(dec, {k: v.to(DEV) for k, v in host[0].items()}, frames)
    step = GraphedTrainStep(dec, host[0])
    direct = []
    for h in host:
        direct.append(float(step(h).cpu()))
    assert len({round(x, 6) for x in direct}) == 3
    with pytest.raises(RuntimeError):
        step()
    step.prefetch(host[0])
    piped = []
    for i in range(6):                                 # replay i consumes batch i % 3 while batch (i + 1) % 3 is in flight
        loss = step()
        step.prefetch(host[(i + 1) % 3])
        piped.append(float(loss.cpu()))
    for i, v in enumerate(piped):
        assert abs(v - direct[i % 3]) <= 1e-6 * max(1.0, abs(direct[i % 3])), (piped, direct)
