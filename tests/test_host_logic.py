"""Host-side logic of the drop-in modules that needs no GPU: state-dict compatibility with the reference, squeeze /
fold index maps, length regulation, synthetic-data determinism."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import flow as of
from radmmm_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SPECS = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))


@pytest.mark.parametrize("tag", ["radmmm", "radtts_accent", "radmmm_spline2"])
def test_state_dict_matches_reference(tag):
    from radmmm_b200 import decoders
    spec = SPECS[tag]
    dec = decoders.RADMMMFlow(**spec["init_args"])
    mine = [[k, list(v.shape), str(v.dtype)] for k, v in dec.state_dict().items()]
    assert mine == spec["state"]                       # same keys, order, shapes and dtypes as the reference
    assert [n for n, _ in dec.named_parameters()] == spec["params"]
    assert dec.decoder_cond_dims == spec["decoder_cond_dims"]
    assert dec.exit_steps == spec["exit_steps"]


def test_loads_reference_layout_weights():
    from radmmm_b200 import decoders
    dec = decoders.RADMMMFlow(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=2)
    sd = syn.synthetic_state_dict(n_flows=2)
    dec.load_state_dict(sd, strict=True)
    assert torch.equal(dec.flows[1].coupling_tfn.affine_param_predictor.in_layers[3].conv.weight_v,
                       sd["flows.1.coupling_tfn.affine_param_predictor.in_layers.3.conv.weight_v"])
    assert dec.n_group_size == 2 and not dec.is_attribute_unconditional()
    # the reference zero-initialises `end` so couplings start as the identity (common.py:799-802)
    fresh = decoders.RADMMMFlow(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=1)
    assert float(fresh.flows[0].coupling_tfn.affine_param_predictor.end.weight.abs().sum()) == 0.0
    # freeze_whitening_layer (decoders.py:142-143)
    frozen = decoders.RADMMMFlow(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=1, freeze_whitening_layer=True)
    assert not any(p.requires_grad for p in frozen.flows[0].invtbl_conv.parameters())


def test_lus_weight_assembly_matches_oracle():
    from radmmm_b200 import common
    sd = syn.synthetic_state_dict(n_flows=2, n_mel_channels=6, n_group_size=2, tag="inv12")
    m = common.Invertible1x1ConvLUS(12)
    m.load_state_dict({k[len("flows.1.invtbl_conv."):]: v for k, v in sd.items() if k.startswith("flows.1.invtbl_conv.")})
    assert torch.allclose(m._weight(), of.lus_weight(sd, "flows.1.invtbl_conv."), atol=1e-7)
    assert torch.allclose(m.log_det(), torch.linalg.slogdet(m._weight().double())[1].float(), atol=1e-5)
    w = common.DataInitializedInvertible1x1Conv(12)
    w.load_state_dict({k[len("flows.0.invtbl_conv."):]: v for k, v in sd.items() if k.startswith("flows.0.invtbl_conv.")})
    assert torch.allclose(w._weight(), of.whiten_weight(sd, "flows.0.invtbl_conv."))
    # fresh LUS init is an orthonormal matrix with det +1 (common.py:511-515)
    fresh = common.Invertible1x1ConvLUS(16)
    W = fresh._weight()
    assert torch.allclose(W @ W.t(), torch.eye(16), atol=1e-5)
    assert abs(float(fresh.log_det())) < 1e-4


def test_squeeze_fold_roundtrip_and_order():
    from radmmm_b200.models.radmmm import squeeze_time, unsqueeze_time
    x = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 9)
    y = squeeze_time(x, 2)
    assert y.shape == (2, 6, 4)
    ref = torch.nn.Unfold((2, 1), stride=2)(x.unsqueeze(-1))          # what the reference does (decoders.py:119-122)
    assert torch.equal(y, ref)
    assert torch.equal(unsqueeze_time(y, 2), x[:, :, :8])
    assert torch.equal(y, of.squeeze_time(x, 2))


def test_length_regulator_matches_oracle():
    from radmmm_b200.common import LengthRegulator
    x = syn.hash_uniform("lr.x", (3, 7, 5))
    dur = torch.tensor([[1, 0, 3, 2, 0, 1, 4], [2, 2, 2, 2, 2, 0, 0], [0, 0, 0, 5, 0, 0, 0]])
    out = LengthRegulator()(x, dur)
    assert torch.equal(out, of.length_regulate(x, dur))


def test_synthetic_data_is_deterministic():
    a = syn.hash_uniform("abc", (5, 7), -2, 3)
    b = syn.hash_uniform("abc", (5, 7), -2, 3)
    assert torch.equal(a, b)
    assert float(a.min()) >= -2 and float(a.max()) < 3
    assert abs(float(syn.hash_uniform("big", (200000,)).mean())) < 0.01
    assert float(a.flatten()[0]) == pytest.approx(float(syn.hash_uniform("abc", (35,), -2, 3)[0]))
    bt = syn.synthetic_batch(4, 64)
    assert int(bt["out_lens"][0]) == 64 and bool((bt["out_lens"][1:] >= 32).all())
    assert float(bt["mel"][1, :, int(bt["out_lens"][1]):].abs().sum()) == 0.0
    p = syn.hash_permutation("p", 16)
    assert sorted(p.tolist()) == list(range(16))


def test_graph_helpers_refuse_cpu_modules():
    """GraphedTrainStep / GraphedInfer are CUDA-graph wrappers: on a CPU module they must fail loudly (no fallback)."""
    import pytest
    import torch
    from radmmm_b200 import decoders, graphs
    dec = decoders.RADMMMFlow(n_accent_dim=8, n_text_dim=520, n_group_size=2, n_flows=1)
    ex = {"mel": torch.zeros(1, 80, 8), "spk_vecs": torch.zeros(1, 16), "context": torch.zeros(1, 520, 8),
          "out_lens": torch.tensor([8]), "f0": torch.zeros(1, 8), "energy_avg": torch.zeros(1, 8), "accent_vecs": torch.zeros(1, 8)}
    with pytest.raises(RuntimeError, match="CUDA"):
        graphs.GraphedTrainStep(dec, ex)
    with pytest.raises(RuntimeError, match="CUDA"):
        graphs.GraphedInfer(dec, {"spk_vec": ex["spk_vecs"], "txt_enc": torch.zeros(1, 520, 2), "dur": torch.ones(1, 2).long(),
                                  "f0": ex["f0"], "energy_avg": ex["energy_avg"], "out_lens": ex["out_lens"]})


def test_length_regulator_total_argument():
    """LengthRegulator with a caller-supplied padded length (no device sync) equals the synced default, zero padded."""
    import torch
    from radmmm_b200.common import LengthRegulator
    x = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    dur = torch.tensor([[2, 0, 3], [1, 1, 1]])
    lr = LengthRegulator()
    a = lr(x, dur)
    b = lr(x, dur, total=7)
    assert a.shape == (2, 5, 4) and b.shape == (2, 7, 4)
    assert torch.equal(b[:, :5], a) and float(b[:, 5:].abs().sum()) == 0.0
    assert torch.equal(a[0, :, 0], torch.tensor([0., 0., 8., 8., 8.]))


def test_pad_batch_and_step_pool_bucketing():
    """pad_batch zero-pads the time axes only; GraphedTrainStepPool picks the smallest bucket, builds one step per bucket
    (here a stub instead of a CUDA graph) and feeds it the padded batch."""
    import pytest
    import torch
    from radmmm_b200.graphs import GraphedTrainStepPool, pad_batch
    bt = {"mel": torch.ones(2, 80, 10), "context": torch.ones(2, 520, 10), "f0": torch.ones(2, 10), "energy_avg": torch.ones(2, 10),
          "spk_vecs": torch.ones(2, 16), "out_lens": torch.tensor([10, 7]), "accent_vecs": torch.ones(2, 8)}
    p = pad_batch(bt, 16)
    assert p["mel"].shape == (2, 80, 16) and p["context"].shape == (2, 520, 16) and p["f0"].shape == (2, 16)
    assert float(p["mel"][:, :, 10:].abs().sum()) == 0.0 and torch.equal(p["mel"][:, :, :10], bt["mel"])
    assert p["spk_vecs"] is bt["spk_vecs"] and torch.equal(p["out_lens"], bt["out_lens"])
    with pytest.raises(ValueError):
        pad_batch(bt, 8)
    built = []

    class Stub:
        def __init__(self, dec, ex):
            built.append(ex["mel"].shape[2])
            self.frames = ex["mel"].shape[2]

        def __call__(self, b):
            assert b["mel"].shape[2] == self.frames
            return self.frames

    pool = GraphedTrainStepPool(object(), [32, 16, 64], step_factory=Stub)
    assert pool(bt) == 16 and pool(bt) == 16 and built == [16]
    bt2 = dict(bt, mel=torch.ones(2, 80, 20), context=torch.ones(2, 520, 20), f0=torch.ones(2, 20), energy_avg=torch.ones(2, 20))
    assert pool(bt2) == 32 and built == [16, 32]
    with pytest.raises(ValueError):
        pool.bucket_for(65)
