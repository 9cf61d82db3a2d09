#!/usr/bin/env python
"""Run one weight-gradient configuration a few times (for ncu / timing), through the C ABI.
    python tools/prof_wgrad.py CIN COUT KH KW DIL N H W [groups] [iters]
"""
import ctypes as C
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import ops, _lib as L  # noqa: E402

cin, cout, kh, kw, dil, n, h, w = [int(v) for v in sys.argv[1:9]]
groups = int(sys.argv[9]) if len(sys.argv) > 9 else 1
iters = int(sys.argv[10]) if len(sys.argv) > 10 else 5
x = ops.new_act(n, cin, h, w, torch.bfloat16, "cuda")
x.copy_(torch.randn(n, cin, h, w, device="cuda"))
g = ops.new_act(n, cout, h, w, torch.bfloat16, "cuda")
g.copy_(torch.randn(n, cout, h, w, device="cuda"))
dw = torch.zeros((kh * kw, cin // groups, cout), dtype=torch.float32, device="cuda")
p = L.EsnConv()
p.x, p.y, p.w = ops.tdesc(x), ops.tdesc(g), dw.data_ptr()
p.kh, p.kw, p.stride = kh, kw, 1
p.pad_h, p.pad_w = dil * (kh // 2), dil * (kw // 2)
p.dil_h, p.dil_w = (dil if kh > 1 else 1), (dil if kw > 1 else 1)
p.groups, p.transposed, p.cout_pad = groups, 0, cout
tc = L.lib.esn_wgrad_umma_supported(C.byref(p))
for _ in range(2):
    L.check(L.lib.esn_conv2d_wgrad(C.byref(p), ops.stream()), "wgrad")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    L.check(L.lib.esn_conv2d_wgrad(C.byref(p), ops.stream()), "wgrad")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
nbytes = (x.numel() + g.numel()) * 2
print("wgrad c%d-%d %dx%d d%d g%d  %dx%dx%d tcgen05=%d: %.4f ms  %.1f GB/s alg  %.1f TFLOP/s" %
      (cin, cout, kh, kw, dil, groups, n, h, w, tc, ms, nbytes / ms / 1e6,
       2.0 * n * h * w * (cin // groups) * cout * kh * kw / ms / 1e9))
