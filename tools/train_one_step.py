#!/usr/bin/env python
"""One eager training iteration of a bench workload inside a cudaProfilerStart/Stop range -- the command profiled by
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ...
    python tools/train_one_step.py [DABNet|ERFNet|FastSCNN|ESPNet_v2] [batch] [H] [W]
"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from oracle import fixture  # noqa: E402
from utils.losses.loss import CrossEntropyLoss2d  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "DABNet"
batch, H, W = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 8), (3, 512), (4, 1024)))
m = build_model(name, 19)
m.load_state_dict(bench.fixture_state_dict(name))
m = m.cuda().train()
crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
from esn.optim import Adam  # noqa: E402  (train.py:212-215's torch.optim.Adam as one launch; what bench.py steps)
opt = Adam(m.parameters(), lr=5e-4, weight_decay=1e-4)
x = fixture.make_input(batch, H, W).cuda()
y = fixture.make_labels(batch, H, W, 19).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):     # the iteration esn.graph.GraphedTrainStep captures
        loss = m.fused_loss(x, y, crit) if hasattr(m, "fused_loss") else crit(m(x), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", name, float(loss))
