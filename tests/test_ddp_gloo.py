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
