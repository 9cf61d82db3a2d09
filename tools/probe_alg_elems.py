#!/usr/bin/env python
"""Block-fused algorithmic elements per input pixel and conv GMACs of a reference net (build container only).

Re-creates probes 2 and 4 of SURVEY.md's appendix: forward hooks on the block classes of SURVEY 8a/8f count, for the
OUTERMOST units only, the float elements entering and leaving each unit (each tensor once per unit); top-level
F.interpolate calls are counted the same way; conv hooks sum the MACs.  The result feeds bench.py's
ALG_ELEMS_PER_PX / GMAC_512x1024 tables (logits term later swapped for the 1-byte mask there).

    python tools/probe_alg_elems.py ESNet ContextNet
"""
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

from make_golden import import_reference

UNITS = {
    "ESNet": ("DownsamplerBlock", "FCU", "PFCU", "UpsamplerBlock"),
    "ContextNet": ("Custom_Conv", "DepthSepConv", "LinearBottleneck", "FeatureFusionModule", "Classifer"),
    "EDANet": ("DownsamplerBlock", "EDAModule"),
    "LEDNet": ("DownsamplerBlock", "SS_nbt_module_paper", "APNModule"),
    "ERFNet": ("DownsamplerBlock", "non_bottleneck_1d", "UpsamplerBlock"),
    "FastSCNN": ("_ConvBNReLU", "_DSConv", "LinearBottleneck", "PyramidPooling", "FeatureFusionModule", "Classifer"),
}


def probe(build_model, name, h=512, w=1024):
    m = build_model(name, 19).eval()
    state = {"depth": 0, "elems": 0, "macs": 0}

    def numel(t):
        if torch.is_tensor(t):
            return t.numel() if t.is_floating_point() else 0
        if isinstance(t, (tuple, list)):
            return sum(numel(v) for v in t)
        return 0

    def pre(mod, inp):
        state["depth"] += 1

    def post(mod, inp, out):
        state["depth"] -= 1
        if state["depth"] == 0:
            state["elems"] += numel(inp) + numel(out)

    def conv_hook(mod, inp, out):
        kh, kw = mod.kernel_size
        if isinstance(mod, nn.ConvTranspose2d):
            x = inp[0]
            state["macs"] += x.shape[2] * x.shape[3] * mod.in_channels * mod.out_channels // mod.groups * kh * kw
        else:
            state["macs"] += out.shape[2] * out.shape[3] * mod.out_channels * (mod.in_channels // mod.groups) * kh * kw
        if state["depth"] == 0:          # a conv that is its own unit (output_conv heads)
            state["elems"] += numel(inp) + numel(out)

    for mod in m.modules():
        if type(mod).__name__ in UNITS[name]:
            mod.register_forward_pre_hook(pre)
            mod.register_forward_hook(post)
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
            mod.register_forward_hook(conv_hook)
    real = F.interpolate

    def counted(x, *a, **k):
        y = real(x, *a, **k)
        if state["depth"] == 0:
            state["elems"] += x.numel() + y.numel()
        return y
    F.interpolate = counted
    try:
        with torch.no_grad():
            m(torch.randn(1, 3, h, w))
    finally:
        F.interpolate = real
    return state["elems"] / (h * w), state["macs"] / 1e9


if __name__ == "__main__":
    build_model, _ = import_reference()
    torch.set_num_threads(8)
    for name in sys.argv[1:] or ["ESNet", "ContextNet"]:
        e, g = probe(build_model, name)
        print("%-12s block-fused elements/px %.1f   GMAC @512x1024 %.2f" % (name, e, g))
