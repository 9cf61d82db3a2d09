#!/usr/bin/env python
"""A/B of esn_conv2d_umma plan variants in ONE process (the library reads these environment variables per call):
    python tools/conv_ab.py "ESN_UMMA_NS=2" "ESN_UMMA_NOHROWS=1" ...
Each shape is timed with the default plan and with every listed VAR=VALUE set (30 launches after 3 warm-ups)."""
import os
import sys
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import ops  # noqa: E402
from esn._lib import ACT_RELU  # noqa: E402

SHAPES = [  # cin, cout, kh, kw, dil, n, h, w, residual
    (128, 64, 3, 3, 1, 16, 128, 256, 0), (128, 64, 3, 3, 4, 16, 128, 256, 0),
    (32, 64, 1, 1, 1, 16, 256, 512, 1), (64, 128, 1, 1, 1, 16, 128, 256, 1),
    (64, 64, 1, 3, 1, 16, 256, 512, 1), (64, 64, 3, 1, 1, 16, 256, 512, 1), (16, 16, 3, 1, 1, 16, 512, 1024, 1),
    (128, 128, 1, 3, 2, 16, 128, 256, 1), (128, 128, 3, 1, 2, 16, 128, 256, 1),
]


def time_one(cin, cout, kh, kw, dil, n, h, w, use_res, iters=30):
    m = nn.Conv2d(cin, cout, (kh, kw), padding=(dil * (kh // 2), dil * (kw // 2)),
                  dilation=(dil if kh > 1 else 1, dil if kw > 1 else 1)).cuda()
    prep = ops.ConvPrep(m, act=ACT_RELU)
    x = ops.new_act(n, cin, h, w, torch.bfloat16, "cuda")
    x.copy_(torch.randn(n, cin, h, w, device="cuda"))
    y = ops.new_act(n, cout, h, w, torch.bfloat16, "cuda")
    res = None
    if use_res:
        res = ops.new_act(n, cout, h, w, torch.bfloat16, "cuda")
        res.copy_(torch.randn(n, cout, h, w, device="cuda"))
    for _ in range(3):
        ops.conv2d(x, prep, out=y, residual=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv2d(x, prep, out=y, residual=res)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, y.float().abs().sum().item()


variants = [None] + sys.argv[1:]
for sh in SHAPES:
    out = []
    for v in variants:
        if v:
            k, val = v.split("=")
            os.environ[k] = val
        ms, chk = time_one(*sh)
        if v:
            del os.environ[v.split("=")[0]]
        out.append("%s %.4f ms (sum %.6g)" % (v or "default", ms, chk))
    print("c%d-%d %dx%d d%d %dx%dx%d res=%d: " % sh + " | ".join(out), flush=True)
