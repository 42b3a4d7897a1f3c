"""world_size-2 gloo test (CPU) of the bucketed gradient reducer used for multi-GPU training."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.flows = torch.nn.ModuleList([torch.nn.Linear(4, 4) for _ in range(3)])
        self.context_lstm = torch.nn.Linear(4, 4)
        self.frozen = torch.nn.Linear(4, 4)
        for p in self.frozen.parameters():
            p.requires_grad = False

    def forward(self, x):
        x = self.context_lstm(x)
        for f in self.flows:
            x = torch.tanh(f(x))
        return self.frozen(x).sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from radmmm_b200.ddp import BucketedGradReducer, default_bucket_key
    torch.manual_seed(0)
    model = _Toy()
    red = BucketedGradReducer(model)
    assert sorted(red.buckets) == ["flow0", "flow1", "flow2", "rest"]
    assert default_bucket_key("flows.3.coupling_tfn.affine_param_predictor.start.bias") == "flow3"
    for step in range(2):
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(100 + rank + 10 * step)
        x = torch.randn(5, 4)
        model(x).backward()
        red.finish()
    grads = torch.cat([p.grad.flatten() for p in model.parameters() if p.requires_grad])
    # every gradient is a view of its bucket (no extra copies survive)
    for b in red.buckets.values():
        for p, v in zip(b["params"], b["views"]):
            assert p.grad.data_ptr() == v.data_ptr()
    out[rank] = grads.clone()
    dist.destroy_process_group()


def _single(seed_list):
    torch.manual_seed(0)
    model = _Toy()
    total = None
    for s in seed_list:
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(s)
        model(torch.randn(5, 4)).backward()
        g = torch.cat([p.grad.flatten() for p in model.parameters() if p.requires_grad])
        total = g if total is None else total + g
    return total / len(seed_list)


def test_bucketed_reducer_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ref = _single([110, 111])          # step 1 seeds of rank 0 and rank 1
    assert torch.allclose(out[0], out[1])
    assert torch.allclose(out[0], ref, atol=1e-6)


def test_bucket_key_keeps_late_gradients_out_of_the_flow_buckets():
    """The 1x1-conv parameters receive their gradients at the end of backward (their matrix is assembled ahead of the flow
    chain): they must not delay a flow bucket's all-reduce."""
    from radmmm_b200.ddp import default_bucket_key
    assert default_bucket_key("flows.3.invtbl_conv.upper_diag") == "rest"
    assert default_bucket_key("flows.0.invtbl_conv.upper") == "rest"
    assert default_bucket_key("flows.7.coupling_tfn.affine_param_predictor.end.weight") == "flow7"
    assert default_bucket_key("context_lstm.weight_ih_l0") == "rest"


# ---------------------------------------------------------------------------------------------- reducer corner cases
class _ToyUnused(_Toy):
    """flows[2] is skipped on odd steps: its bucket receives no gradient at all in that step."""

    def forward(self, x, skip_last=False):
        x = self.context_lstm(x)
        for f in (self.flows[:2] if skip_last else self.flows):
            x = torch.tanh(f(x))
        return self.frozen(x).sum()


def _worker_corner(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from radmmm_b200.ddp import BucketedGradReducer
    torch.manual_seed(0)
    model = _ToyUnused()
    red = BucketedGradReducer(model)
    # (1) a bucket whose parameters get no gradient this step is still reduced (zeros) -- no rank may skip a collective
    for p in model.parameters():
        p.grad = None
    torch.manual_seed(200 + rank)
    model(torch.randn(5, 4), skip_last=True).backward()
    red.finish()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.flows[2].parameters())
    g_first = model.flows[0].weight.grad.clone()
    # (2) zero_grad(set_to_none=False) then backward: the sink declines (the slot holds the live gradient) and autograd's
    #     accumulate gives zero + new, never 2 x new
    p0 = model.flows[0].weight
    assert red.fresh_view(p0) is None                      # .grad is set -> no aliasing view is handed out
    for p in model.parameters():
        if p.grad is not None:
            p.grad.zero_()
    torch.manual_seed(200 + rank)
    model(torch.randn(5, 4), skip_last=True).backward()
    red.finish()
    assert torch.allclose(model.flows[0].weight.grad, g_first, atol=1e-7)
    p0.grad = None
    assert red.fresh_view(p0) is not None and red.fresh_view(p0).data_ptr() == red.grad_view(p0).data_ptr()
    out[rank] = g_first
    dist.destroy_process_group()


def test_reducer_unused_bucket_and_accumulate():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_corner, args=(world, port, out), nprocs=world, join=True)
    assert torch.allclose(out[0], out[1])


# ---------------------------------------------------------------------------------------------- the other collectives
def _worker_sync(rank, world, port, out):
    """MaskedBatchNorm1d.distributed_sync (maskedbatchnorm1d.py:88-95) and the whitening-init broadcast
    (common.py:584-586) over a real process group."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from radmmm_b200 import synthetic as syn
    from radmmm_b200.common import DataInitializedInvertible1x1Conv
    from radmmm_b200.splines import MaskedBatchNorm1d
    rows, ch = 40, 6
    x_all = syn.hash_uniform("syncbn.x", (world * rows, ch), -2, 2)
    m_all = (syn.hash_uniform("syncbn.m", (world * rows, 1), 0, 1) < 0.7).float()
    bn = MaskedBatchNorm1d(ch)
    bn.weight.data.copy_(syn.hash_uniform("syncbn.w", (ch,), 0.5, 1.5))
    bn.bias.data.copy_(syn.hash_uniform("syncbn.b", (ch,), -0.5, 0.5))
    bn.distributed_sync = True
    bn.train()
    x = x_all[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
    y = bn.forward_rows(x, m_all[rank * rows:(rank + 1) * rows])
    (y * m_all[rank * rows:(rank + 1) * rows]).sum().backward()
    out[("bn_y", rank)] = y.detach().clone()
    out[("bn_rm", rank)] = bn.running_mean.clone()
    out[("bn_dx", rank)] = x.grad.clone()
    # whitening init: every rank sees different data, all end up with rank 0's matrix and mean
    conv = DataInitializedInvertible1x1Conv(8)
    conv.train()
    data = syn.hash_uniform(f"winit.{rank}", (2, 8, 30), -2, 2)
    conv.initialize(data, torch.tensor([30, 17]))
    out[("w_diag", rank)] = conv.upper_diag.detach().clone()
    out[("w_mean", rank)] = conv.input_mean.clone()
    dist.destroy_process_group()


def test_masked_bn_sync_and_whitening_broadcast():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import flow as of
    from radmmm_b200 import synthetic as syn
    from radmmm_b200.splines import MaskedBatchNorm1d
    world, rows, ch = 2, 40, 6
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_sync, args=(world, port, out), nprocs=world, join=True)
    # single-process statistics over the concatenated rows are what the synced ranks must have used
    x_all = syn.hash_uniform("syncbn.x", (world * rows, ch), -2, 2).requires_grad_(True)
    m_all = (syn.hash_uniform("syncbn.m", (world * rows, 1), 0, 1) < 0.7).float()
    bn = MaskedBatchNorm1d(ch)
    bn.weight.data.copy_(syn.hash_uniform("syncbn.w", (ch,), 0.5, 1.5))
    bn.bias.data.copy_(syn.hash_uniform("syncbn.b", (ch,), -0.5, 0.5))
    bn.train()
    y = bn.forward_rows(x_all, m_all)
    (y * m_all).sum().backward()
    for r in range(world):
        assert torch.allclose(out[("bn_y", r)], y[r * rows:(r + 1) * rows].detach(), atol=1e-5)
        assert torch.allclose(out[("bn_rm", r)], bn.running_mean, atol=1e-6)
        assert torch.allclose(out[("bn_dx", r)], x_all.grad[r * rows:(r + 1) * rows], atol=1e-5)
    # whitening init: rank 0's statistics everywhere, equal to the oracle's init on rank 0's data
    mean_ref, _, diag_ref = of.whitening_init(syn.hash_uniform("winit.0", (2, 8, 30), -2, 2), torch.tensor([30, 17]))
    for r in range(world):
        assert torch.allclose(out[("w_diag", r)], diag_ref, atol=1e-4)
        assert torch.allclose(out[("w_mean", r)].flatten(), mean_ref.flatten(), atol=1e-5)
