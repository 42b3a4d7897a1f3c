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
    dec.invalidate_weight_cache()
    for p in dec.parameters():
        p.grad = None
    out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], frames), f0=bt["f0"],
              energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
    loss, _ = L.flow_nll(out["z_mel"], out["log_det_W_list"], out["log_s_list"], bt["out_lens"] // 2)
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
