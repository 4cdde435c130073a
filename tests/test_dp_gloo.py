"""world_size-2 gloo test (CPU) of the data-parallel gradient bucketing host logic."""
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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vae_play_b200.parallel import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    params = list(net.parameters())
    gb = GradBuckets(params, world, bucket_mb=0.0001)   # tiny buckets: several of them, exercised in reverse order
    assert len(gb.buckets) >= 3
    res = []
    for it in range(2):
        torch.manual_seed(100 + rank + 10 * it)
        x = torch.randn(4, 7)
        net.zero_grad(set_to_none=True)
        # equivalent single-process loss: (1/W) sum_r mean_r + sum_r sum_r  (mean-type + sum-type terms)
        y = net(x)
        loss = y.pow(2).mean() / world + y.sum()
        loss.backward()
        gb.allreduce()
        res.append([p.grad.clone() for p in params])
        for p in params:    # grads live inside the buckets
            assert p.grad.data_ptr() == gb.slot[id(p)][1].data_ptr()
    # reference: both ranks' batches in one process
    outs = []
    for it in range(2):
        net.zero_grad(set_to_none=True)
        tot = 0
        for r in range(world):
            torch.manual_seed(100 + r + 10 * it)
            x = torch.randn(4, 7)
            y = net(x)
            tot = tot + y.pow(2).mean() / world + y.sum()
        gs = torch.autograd.grad(tot, params)
        outs.append(gs)
    ok = all(torch.allclose(a, b, atol=1e-5) for ra, rb in zip(res, outs) for a, b in zip(ra, rb))
    # every slot starts 16-byte aligned (vector reductions of the weight-gradient kernels), whatever the parameter sizes
    ok = ok and all(gb.slot[id(p)][2] % 4 == 0 for p in params)
    # the pipelined exchange of bench.py: all buckets queued at once, consumed one by one (update of bucket i while
    # bucket i+1 is still being reduced); result identical to allreduce()
    gb.overlap = False          # graph-replay mode of bench.py: the hooks only place gradients, the caller issues the collectives
    torch.manual_seed(500 + rank)
    x = torch.randn(4, 7)
    net.zero_grad(set_to_none=True)
    y = net(x)
    (y.pow(2).mean() / world + y.sum()).backward()
    local = [p.grad.clone() for p in params]
    early = gb.buckets_within(params[2:])               # buckets made only of the later layers' parameters
    ok = ok and len(early) >= 1 and len(early) < len(gb.buckets)
    gb.allreduce_subset(early)
    gb.allreduce_subset(range(len(gb.buckets)))         # the rest; already-issued buckets are not issued twice
    for bi in range(len(gb.buckets)):
        gb.wait_bucket(bi)
    summed = [g.clone() for g in local]
    for g in summed:
        dist.all_reduce(g)
    ok = ok and all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(params, summed))
    ok = ok and all(b["handle"] is None and b["ready"] == 0 for b in gb.buckets)
    # several backward calls per step (reference train.py:68-73): all but the last inside no_sync(); with overlap on, a
    # second synchronising backward must raise instead of reducing a bucket twice
    gb.overlap = True
    torch.manual_seed(700 + rank)
    x = torch.randn(4, 7)
    net.zero_grad(set_to_none=True)
    y = net(x)
    l1, l2 = y.pow(2).mean() / world, y.sum()
    with gb.no_sync():
        l1.backward(retain_graph=True)
    l2.backward()
    gb.allreduce()
    got = [p.grad.clone() for p in params]
    net.zero_grad(set_to_none=True)
    with gb.no_sync():
        (net(x).pow(2).mean() / world + net(x).sum()).backward()
    want = [p.grad.clone() for p in params]
    for g in want:
        dist.all_reduce(g)
    ok = ok and all(torch.allclose(a, b, atol=1e-5) for a, b in zip(got, want))
    net.zero_grad(set_to_none=True)
    y = net(x)
    y.sum().backward(retain_graph=True)
    raised = False
    try:
        y.pow(2).mean().backward()
    except RuntimeError:
        raised = True
    ok = ok and raised
    for b in gb.buckets:            # drain what the first backward launched
        if b["handle"] is not None:
            b["handle"].wait()
            b["handle"] = None
        b["ready"] = 0
    q.put((rank, ok))
    gb.remove()
    dist.destroy_process_group()


def test_grad_buckets_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
