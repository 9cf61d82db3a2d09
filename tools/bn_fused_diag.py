#!/usr/bin/env python
"""Diagnosis: run every train-mode BatchNorm layer of a net's forward through BOTH the one-launch kernel and the three-launch
path on the same input and print where they differ (output, saved scale / shift / mean / invstd).
    python tools/bn_fused_diag.py [FastSCNN] [N H W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from esn import train as T  # noqa: E402
from oracle import fixture  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "FastSCNN"
n, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 64, 128)
m = build_model(name, 19)
m.load_state_dict(bench.fixture_state_dict(name))
m = m.cuda().train()
x = fixture.make_input(n, h, w).cuda()
orig = T.BNActT.forward
idx = [0]


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def both(self, tape, xv, out=None):
    if self.bn is None:
        return orig(self, tape, xv, out=out)
    sd = {k: v.clone() for k, v in self.bn.state_dict().items()}
    T.FUSED_BN = False
    t2 = T.Tape()
    y_ref = orig(self, t2, xv, out=None)
    self.bn.load_state_dict(sd)
    T.FUSED_BN = True
    steps0 = len(tape.steps)
    y = orig(self, tape, xv, out=out)
    torch.cuda.synchronize()
    # saved statistics live in the closures: compare through a backward of a fixed gradient instead
    g = torch.randn_like(y_ref.t)
    r = rel(y.t.float(), y_ref.t.float())
    mx = (y.t.float() - y_ref.t.float()).abs().max().item()
    print("BN %3d  x %-22s stride %4d  act %d  fused=%s  out rel %.3e  max|d| %.3e  |y| max %.3e" % (
        idx[0], tuple(xv.t.shape), xv.t.stride(3), self.act, len(tape.steps) > steps0 and T._v8(xv.t), r, mx,
        y_ref.t.float().abs().max().item()), flush=True)
    idx[0] += 1
    return y


T.BNActT.forward = both
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = m(x)
print("done", tuple(out.shape))
