#!/usr/bin/env python
"""Run the fused pair kernel once at a bench shape (with csrc built with EXTRA=-DESN_PAIR_TIMING it prints the
per-role wait breakdown):  python tools/pair_timing.py C H W [residual 0|1] [batch]"""
import sys, torch
sys.path[:0]=['/root/repo','/root/repo/efficient-segmentation-networks_b200']
import torch.nn as nn
from esn import ops
from esn._lib import ACT_RELU
C_, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
with_res = int(sys.argv[4]) if len(sys.argv) > 4 else 1
N = int(sys.argv[5]) if len(sys.argv) > 5 else 16
m1 = nn.Conv2d(C_, C_, (3, 1), padding=(1, 0)).cuda(); m2 = nn.Conv2d(C_, C_, (1, 3), padding=(0, 1)).cuda()
p1 = ops.ConvPrep(m1, act=ACT_RELU); p2 = ops.ConvPrep(m2, None, None, ACT_RELU)
x = ops.new_act(N, C_, H, W, torch.bfloat16, "cuda"); x.copy_(torch.randn(N, C_, H, W, device="cuda"))
res = None
if with_res:
    res = ops.new_act(N, C_, H, W, torch.bfloat16, "cuda"); res.copy_(torch.randn(N, C_, H, W, device="cuda"))
out = ops.new_act(N, C_, H, W, torch.bfloat16, "cuda")
ref = ops.conv2d(ops.conv2d(x, p1), p2, residual=res)
for _ in range(3):
    ops.conv_pair(x, p1, p2, out=out, residual=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
e0.record()
for _ in range(reps):
    ops.conv_pair(x, p1, p2, out=out, residual=res)
e1.record()
torch.cuda.synchronize()
print("ms per launch %.4f" % (e0.elapsed_time(e1) / reps))
print("ok", C_, H, W, with_res, ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item())
