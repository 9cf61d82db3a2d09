#!/usr/bin/env python
"""Every C-ABI launch of one eager inference step with its CUDA-event time and algorithmic bytes:
    python tools/layer_times.py DABNet 16 1024 2048 [kernel-substring]"""
import json
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from builders.model_builder import build_model  # noqa: E402
from esn import ops  # noqa: E402
from oracle import fixture  # noqa: E402

name, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
flt = sys.argv[5] if len(sys.argv) > 5 else ""
spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
proto = {k: torch.empty(shape, dtype=getattr(torch, dt.split(".")[1])) for k, shape, dt in spec[name]["keys"]}
m = build_model(name, 19)
m.load_state_dict(fixture.randomize_state_dict(proto, 1234))
m = m.cuda().eval()
x = fixture.make_input(n, h, w).cuda()
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        ops.PROFILE = []
        m.predict_mask(x)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
tot = 0.0
for i, r in enumerate(prof):
    ms = r["ev"][0].elapsed_time(r["ev"][1])
    tot += ms
    if flt in r["kernel"]:
        print("%3d %-28s %-32s %8.4f ms  %8.1f MB  %7.1f GB/s" % (i, r["kernel"], r["tag"], ms, r["bytes"] / 1e6, r["bytes"] / ms / 1e6))
print("total %.3f ms in %d launches" % (tot, len(prof)))
