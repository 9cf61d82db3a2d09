#!/usr/bin/env python
"""Run one tcgen05 conv configuration a few times (for ncu / timing).
    python tools/prof_conv.py CIN COUT KH KW DIL N H W [residual] [iters]
"""
import os
import sys
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import ops  # noqa: E402
from esn._lib import ACT_RELU  # noqa: E402

cin, cout, kh, kw, dil, n, h, w = [int(v) for v in sys.argv[1:9]]
use_res = len(sys.argv) > 9 and sys.argv[9] == "1"
iters = int(sys.argv[10]) if len(sys.argv) > 10 else 5
m = nn.Conv2d(cin, cout, (kh, kw), padding=(dil * (kh // 2), dil * (kw // 2)), dilation=(dil if kh > 1 else 1, dil if kw > 1 else 1)).cuda()
prep = ops.ConvPrep(m, act=ACT_RELU)
x = ops.new_act(n, cin, h, w, torch.bfloat16, "cuda")
x.copy_(torch.randn(n, cin, h, w, device="cuda"))
res = x if (use_res and cin == cout) else None
y = ops.new_act(n, cout, h, w, torch.bfloat16, "cuda")
for _ in range(2):
    ops.conv2d(x, prep, out=y, residual=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv2d(x, prep, out=y, residual=res)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
nbytes = (x.numel() + y.numel() + (x.numel() if res is not None else 0)) * 2
print("conv c%d-%d %dx%d d%d  %dx%dx%d res=%d: %.4f ms  %.1f GB/s alg  %.1f TFLOP/s" %
      (cin, cout, kh, kw, dil, n, h, w, int(res is not None), ms, nbytes / ms / 1e6,
       2.0 * n * h * w * cin * cout * kh * kw / ms / 1e9))
