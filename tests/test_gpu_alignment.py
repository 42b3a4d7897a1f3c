"""GPU parity of the hard-alignment path: batched monotonic alignment search and the attention CTC loss,
against reference-made fixtures (tests/golden/alignment.npz), the oracle restatement on larger ragged batches and -- when
the staged reference and numba are importable -- the reference's own numba kernel at the benchmark's shape."""
import numpy as np
import pytest
import torch

from oracle import alignment as oa
from radmmm_b200 import synthetic as syn
from tests.gpu_util import DEV, close, gold

pytestmark = pytest.mark.gpu


def _maps(tag, B, T1, T2, in_lens, out_lens, sharp):
    t1 = torch.arange(T1, dtype=torch.float32)[None, :, None]
    t2 = torch.arange(T2, dtype=torch.float32)[None, None, :]
    il, ol = torch.tensor(in_lens), torch.tensor(out_lens)
    centre = t1 * (il[:, None, None].float() / ol[:, None, None].float().clamp(min=1))
    logits = -torch.tensor(sharp)[:, None, None] * (t2 - centre) ** 2 + 2.0 * syn.hash_uniform(tag, (B, T1, T2))
    attn = torch.softmax(logits.masked_fill(t2 >= il[:, None, None], -float("inf")), dim=2)
    return torch.nan_to_num(attn).unsqueeze(1).contiguous(), logits.unsqueeze(1).contiguous(), il, ol


def test_mas_vs_reference_fixture():
    from radmmm_b200.alignment import binarize_attention, mas_width1
    gd = gold("alignment.npz")
    hard = binarize_attention(gd["attn"].to(DEV), gd["in_lens"].to(DEV), gd["out_lens"].to(DEV))
    assert torch.equal(hard.cpu(), gd["hard"])
    # lengths as host lists, single-map entry point with numpy in / numpy out
    hard2 = binarize_attention(gd["attn"].to(DEV), gd["in_lens"].tolist(), gd["out_lens"].tolist())
    assert torch.equal(hard2.cpu(), gd["hard"])
    one = mas_width1(gd["attn"][1, 0, :57, :23].numpy())
    assert isinstance(one, np.ndarray) and np.array_equal(one, gd["hard"][1, 0, :57, :23].numpy())


@pytest.mark.parametrize("B,T1,T2,in_lens,out_lens", [
    (6, 300, 90, [90, 1, 45, 77, 13, 60], [300, 10, 299, 150, 8, 1]),          # 1 text position, 1 frame, frames < text
    (3, 40, 1500, [1500, 1100, 33], [40, 39, 40]),                            # > 1024 text positions: 2 columns per thread
    (2, 64, 2500, [2500, 2049], [64, 50]),                                    # 4 columns per thread
    (2, 2048, 1024, [1024, 700], [2048, 1500]),                               # back pointers in the global workspace
    (8, 800, 120, [120, 97, 88, 110, 64, 101, 119, 75], [800, 611, 540, 777, 402, 650, 790, 480]),      # benchmark shape
])
def test_mas_vs_oracle(B, T1, T2, in_lens, out_lens):
    from radmmm_b200.alignment import binarize_attention
    sharp = [0.02, 0.3, 4.0, 0.1, 1.0, 0.05, 0.5, 0.2][:B]
    attn, _, il, ol = _maps(f"mas.{B}.{T1}.{T2}", B, T1, T2, in_lens, out_lens, sharp)
    hard = binarize_attention(attn.to(DEV), il.to(DEV), ol.to(DEV)).cpu()
    ref = oa.binarize_attention(attn, il, ol)
    assert torch.equal(hard, ref), f"{(hard != ref).sum().item()} cells differ"
    # structure: inside the valid box one text position per frame (two in frame 0 when the path cannot reach 0), monotone
    for b in range(B):
        h = hard[b, 0, :ol[b], :il[b]]
        assert hard[b].sum() == h.sum()
        pos = h[1:].argmax(dim=1)                  # frame 0 may carry the reference's extra opt[0, 0] = 1
        assert (h.sum(1)[1:] == 1).all() and (pos[1:] >= pos[:-1]).all() and (pos[1:] - pos[:-1] <= 1).all()
        assert h[-1, il[b] - 1] == 1 and h[0, 0] == 1
    # the log-probability entry point: same map from np.log of the probabilities
    with np.errstate(divide="ignore"):
        logp = torch.from_numpy(np.log(attn.numpy()))
    hard_log = binarize_attention(logp.to(DEV), il.to(DEV), ol.to(DEV), is_log=True).cpu()
    assert torch.equal(hard_log, ref)


def test_mas_vs_staged_reference_numba():
    """The reference's own numba kernel (staged copy, oracle/_ref) on the benchmark-shaped batch."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference not staged")
    pytest.importorskip("numba")
    ref_import.import_reference()
    from alignment import mas_width1 as ref_mas
    from radmmm_b200.alignment import binarize_attention
    B, T1, T2 = 8, 800, 120
    in_lens, out_lens = [120, 97, 88, 110, 64, 101, 119, 75], [800, 611, 540, 777, 402, 650, 790, 480]
    attn, _, il, ol = _maps("mas.numba", B, T1, T2, in_lens, out_lens, [0.02, 0.3, 4.0, 0.1, 1.0, 0.05, 0.5, 0.2])
    hard = binarize_attention(attn.to(DEV), il.to(DEV), ol.to(DEV)).cpu()
    a = attn.numpy()
    for b in range(B):
        ref = ref_mas(a[b, 0, :out_lens[b], :in_lens[b]])
        assert np.array_equal(hard[b, 0, :out_lens[b], :in_lens[b]].numpy(), ref), f"utterance {b}"


def test_attention_ctc_vs_reference_fixture():
    from radmmm_b200.loss import AttentionCTCLoss
    gd = gold("alignment.npz")
    lp = gd["logprob"].to(DEV).requires_grad_(True)
    cost, each = AttentionCTCLoss()(lp, gd["in_lens"].to(DEV), gd["out_lens"].to(DEV), return_all=True)
    (cost * 3.0).backward()
    close(cost.detach(), gd["ctc_cost"], 2e-5, what="ctc cost vs reference")
    close(torch.stack(each), gd["ctc_each"], 5e-5, what="per-utterance ctc vs reference")
    assert float(each[3]) == 0.0                                   # infeasible (5 frames, 7 keys): zero_infinity
    close(lp.grad / 3.0, gd["ctc_grad"], 2e-6, what="ctc gradient vs reference")
    assert float(lp.grad[3].abs().max()) == 0.0


@pytest.mark.parametrize("B,T1,T2,in_lens,out_lens", [
    (5, 200, 60, [60, 1, 33, 47, 20], [200, 5, 150, 47, 199]),                # T == K (only the no-blank path is feasible)
    (8, 800, 120, [120, 97, 88, 110, 64, 101, 119, 75], [800, 611, 540, 777, 402, 650, 790, 480]),
    (2, 1200, 1000, [1000, 640], [1200, 1100]),
])
def test_attention_ctc_vs_oracle(B, T1, T2, in_lens, out_lens):
    from radmmm_b200.loss import AttentionCTCLoss
    _, logits, il, ol = _maps(f"ctc.{B}.{T1}.{T2}", B, T1, T2, in_lens, out_lens, [0.02, 0.3, 1.0, 0.1, 1.0, 0.05, 0.5, 0.2][:B])
    lp = logits.to(DEV).requires_grad_(True)
    cost, each = AttentionCTCLoss(blank_logprob=-1)(lp, il, ol, return_all=True)
    cost.backward()
    lr = logits.clone().double().requires_grad_(True)
    cost_r, each_r = oa.attention_ctc_loss(lr, il, ol)
    cost_r.backward()
    close(cost.detach(), cost_r.detach(), 1e-6, rtol=2e-5, what="ctc cost")
    close(torch.stack(each), torch.stack([e.detach() for e in each_r]), 1e-6, rtol=5e-5, what="per-utterance ctc")
    close(lp.grad, lr.grad, 5e-6, what="ctc gradient")      # fp32 alpha-beta over up to 1200 frames vs the fp64 oracle
    # nothing outside the valid boxes
    for b in range(B):
        assert float(lp.grad[b, 0, ol[b]:].abs().sum()) == 0.0 and float(lp.grad[b, 0, :, il[b]:].abs().sum()) == 0.0
