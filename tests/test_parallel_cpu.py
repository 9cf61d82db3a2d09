"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat gradient buckets filled in
tape order, all-reduce launched per bucket, SUM semantics, untouched parameters reduced as zeros,
parameter broadcast from rank 0."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-segmentation-networks_b200")


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from esn import parallel, train as T
    torch.manual_seed(rank)          # different initial weights per rank: broadcast must fix that
    m = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 4, 1), nn.Conv2d(4, 4, 1))
    parallel.data_parallel(m, bucket_bytes=256)
    w0 = m[0].weight.detach().clone()
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert torch.equal(gathered[0], gathered[1])
    buckets = m.__dict__["_esn_buckets"]
    assert len(buckets.buckets) >= 2
    tape = T.Tape(buckets)
    params = [p for p in m.parameters()]
    used = params[:-2]               # the last conv produces no gradient this step
    for p in reversed(used):         # tape order = reverse registration order
        tape.add_param_grad(p, torch.full_like(p, float(rank + 1)))
    grads = tape.backward()
    for p in used:
        assert torch.allclose(grads[p], torch.full_like(p, 3.0)), "SUM over ranks 1 + 2"
    for i, b in enumerate(buckets.buckets):
        for p in b:
            if p not in grads:
                assert float(buckets.views[i][p].abs().sum()) == 0.0
    out[rank] = 1
    dist.destroy_process_group()


def test_grad_buckets_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
