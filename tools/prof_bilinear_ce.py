#!/usr/bin/env python
"""The fused training close (esn_bilinear_ce) at a given size next to the four launches it replaces
(esn_head_bilinear -> esn_weighted_ce x 2 -> esn_bilinear_bwd), for timing / ncu.
    python tools/prof_bilinear_ce.py N CLASSES h w SCALE [iters]       e.g. 8 19 64 128 8
"""
import ctypes as C
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import ops, _lib as L  # noqa: E402

n, c, h, w, s = [int(v) for v in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
H, W = s * h, s * w
torch.manual_seed(0)
x = ops.new_act(n, c, h, w, torch.bfloat16, "cuda", c_alloc=32)
x.copy_(torch.randn(n, c, h, w, device="cuda") * 3)
tgt = torch.randint(0, c, (n, H, W), device="cuda")
tgt[torch.rand(n, H, W, device="cuda") < 0.1] = 255
wt = torch.rand(c, device="cuda") + 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def fused():
    return ops.bilinear_ce(x, tgt, wt, 255, H, W)


def separate():
    logits, _ = ops.head_bilinear(x, c, H, W, True, False, torch.float32)
    sums, _ = ops.weighted_ce(logits, tgt, wt, 255, want_grad=False)
    one = torch.ones(1, device="cuda")
    _, g = ops.weighted_ce(logits, tgt, wt, 255, want_grad=True, sums=torch.zeros(2, device="cuda"), gnorm=one, gout=one)
    dlow = ops.new_act(n, c, h, w, torch.bfloat16, "cuda", c_alloc=32)
    a, b = ops.tdesc(g), ops.tdesc(dlow)
    a.layout, a.c_stride = L.ESN_NCHW, 0
    ops._call(L.lib.esn_bilinear_bwd, "esn_bilinear_bwd", (C.byref(a), C.byref(b), C.c_float(1.0)), g.numel() * 4)
    return sums, dlow


def timed(fn):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


tf, ts = timed(fused), timed(separate)
sf, df = fused()
ss, dsep = separate()
rel = ((df.float() - dsep.float()).norm() / dsep.float().norm()).item()
print("bilinear + CE close %dx%dx%dx%d x%d: fused %.4f ms | four launches %.4f ms | loss sums %.6g vs %.6g | d scores rel-L2 %.2e (bf16 rounding of the separate path)"
      % (n, c, h, w, s, tf, ts, sf[0].item(), ss[0].item(), rel))
