#!/usr/bin/env python
"""Data-parallel correctness on N GPUs (run under torchrun): with IDENTICAL data on every rank the
all-reduced gradients must equal the single-process gradients (each rank's loss is divided by the
global sum of weights, gradients are summed), and different data must give identical results on all
ranks.  Prints one line per check; exit code 0 = pass."""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from esn import parallel  # noqa: E402
from oracle import fixture  # noqa: E402
from utils.losses.loss import CrossEntropyLoss2d  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wt = torch.tensor(fixture.CLASS_WEIGHTS)


def run(dp, seed, distributed_loss, dtype, fused=False):
    m = build_model("DABNet", 19)
    m.load_state_dict(bench.fixture_state_dict("DABNet"))
    m = m.cuda().train()
    if dp:
        parallel.data_parallel(m)
    crit = CrossEntropyLoss2d(weight=wt, ignore_label=255, distributed=distributed_loss).cuda()
    x = fixture.make_input(2, 128, 256, seed=seed).cuda()
    y = fixture.make_labels(2, 128, 256, 19, seed=seed).cuda()
    call = (lambda: m.fused_loss(x, y, crit)) if fused else (lambda: crit(m(x), y))     # fused: esn_bilinear_ce close
    if dtype == torch.bfloat16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = call()
    else:
        loss = call()
    loss.backward()
    return loss.detach(), torch.cat([p.grad.flatten().float() for p in m.parameters()])


ok = True
for dtype in (torch.float32, torch.bfloat16):
    l_ref, g_ref = run(False, 7, False, dtype)           # single process, local loss
    l_dp, g_dp = run(True, 7, True, dtype)               # same data on every rank
    rel = ((g_dp - g_ref).norm() / g_ref.norm()).item()
    same = abs(l_dp.item() - l_ref.item()) / abs(l_ref.item())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    if rank == 0:
        print("%s identical-data: loss rel diff %.2e, all-reduced grad vs single-process rel-L2 %.2e (tol %.0e)" % (dtype, same, rel, tol))
    ok &= rel < tol and same < 1e-5
    l_f, g_f = run(True, 7, True, dtype, fused=True)     # the fused head + loss (DABNet.fused_loss), same data on every rank
    rel_f = ((g_f - g_ref).norm() / g_ref.norm()).item()
    same_f = abs(l_f.item() - l_ref.item()) / abs(l_ref.item())
    if rank == 0:
        print("%s identical-data, fused close: loss rel diff %.2e, all-reduced grad vs single-process rel-L2 %.2e" % (dtype, same_f, rel_f))
    ok &= rel_f < max(tol, 2e-4) and same_f < 1e-5
    l2, g2 = run(True, 100 + rank, True, dtype)          # different data per rank
    gs = [torch.zeros_like(g2) for _ in range(world)]
    dist.all_gather(gs, g2)
    ls = [torch.zeros_like(l2) for _ in range(world)]
    dist.all_gather(ls, l2)
    d = max(((g - gs[0]).abs().max() / gs[0].abs().max()).item() for g in gs)
    if rank == 0:
        print("%s sharded-data: global loss %.6f on every rank (max diff %.1e), gradient max diff across ranks %.1e" %
              (dtype, ls[0].item(), max(abs(l.item() - ls[0].item()) for l in ls), d))
    ok &= d == 0.0
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
