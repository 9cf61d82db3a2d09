#!/usr/bin/env python
"""Diagnosis: parameter gradients of one training step with the fused BatchNorm forward on vs off (same weights, same batch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from esn import train as T  # noqa: E402
from oracle import fixture  # noqa: E402
from utils.losses.loss import CrossEntropyLoss2d  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "FastSCNN"
res = {}
for mode in ("bwd", "fwd", "fwd,bwd"):
    T._FUSED_DIR = mode
    torch.manual_seed(0)
    m = build_model(name, 19)
    m.load_state_dict(bench.fixture_state_dict(name))
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
    x = fixture.make_input(2, 64, 128).cuda()
    lab = fixture.make_labels(2, 64, 128, 19).cuda()
    crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = crit(m(x), lab)
    loss.backward()
    torch.cuda.synchronize()
    res[mode] = (loss.item(), {k: p.grad.detach().float().clone() for k, p in m.named_parameters()})
print("loss", {k: v[0] for k, v in res.items()})
ref = res["bwd"][1]
for k in ref:
    a, b = res["fwd"][1][k], ref[k]
    r = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    print("%-60s %-18s |g| %.3e  fwd-fused vs unfused-fwd rel %.3e" % (k, tuple(b.shape), b.norm().item(), r))
