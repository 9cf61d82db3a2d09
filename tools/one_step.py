#!/usr/bin/env python
"""One eager (no CUDA graph) inference step of a bench workload -- the command profiled by ncu.
    python tools/one_step.py [ERFNet|DABNet] [batch] [H] [W] [steps]
"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from oracle import fixture  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ERFNet"
batch, H, W = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 16), (3, 1024), (4, 2048)))
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
m = build_model(name, 19)
m.load_state_dict(bench.fixture_state_dict(name))
m = m.cuda().eval()
x = fixture.make_input(batch, H, W).cuda()
for _ in range(steps):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        mask = m.predict_mask(x)
torch.cuda.synchronize()
print("ok", name, tuple(mask.shape), int(mask.sum()))
