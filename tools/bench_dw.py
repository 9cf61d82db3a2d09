#!/usr/bin/env python
"""Microbenchmark of the depthwise kernels behind esn_conv2d_direct: register-strip kernel vs the round-1 gather kernel
(ESN_DISABLE_DW_STRIP=1) on the shapes of CGNet / Fast-SCNN / ESPNetv2 / DABNet at 1024x2048 input, bf16.
Algorithmic bytes = |x| + |y|; CUDA events, 20 launches after 5 warm-ups, tensors >> L2 or an L2 flush in between."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

SHAPES = [
    # name, N, C, H, W, k, dil
    ("cgnet F_loc c64 @128x256 b32", 32, 64, 128, 256, (3, 3), (1, 1)),
    ("cgnet F_sur c64 d4 @128x256 b32", 32, 64, 128, 256, (3, 3), (4, 4)),
    ("cgnet F_loc c32 @256x512 b32", 32, 32, 256, 512, (3, 3), (1, 1)),
    ("cgnet F_sur c32 d2 @256x512 b32", 32, 32, 256, 512, (3, 3), (2, 2)),
    ("fastscnn dw c128 @128x256 b16", 16, 128, 128, 256, (3, 3), (1, 1)),
    ("fastscnn dw c384 @64x128 b16", 16, 384, 64, 128, (3, 3), (1, 1)),
    ("espnetv2 dw c128 d2 @64x128 b16", 16, 128, 64, 128, (3, 3), (2, 2)),
    ("dabnet 3x1 c32 @256x512 b16", 16, 32, 256, 512, (3, 1), (1, 1)),
    ("dabnet 1x3 c32 d2 @256x512 b16", 16, 32, 256, 512, (1, 3), (1, 2)),
    ("dabnet 3x1 c64 d16 @128x256 b16", 16, 64, 128, 256, (3, 1), (16, 1)),
    ("dabnet-train 3x1 c32 @128x256 b8", 8, 32, 128, 256, (3, 1), (1, 1)),
]


def main():
    from esn import ops
    from esn._lib import ACT_PRELU
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, n, c, h, w, k, dil in SHAPES:
        pad = ((k[0] // 2) * dil[0], (k[1] // 2) * dil[1])
        m = nn.Conv2d(c, c, k, 1, pad, dil, c, bias=False).cuda()
        prep = ops.ConvPrep(m, torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda"), ACT_PRELU, torch.rand(c, device="cuda"))
        x = ops.new_act(n, c, h, w, torch.bfloat16, "cuda")
        x.normal_()
        y = ops.new_act(n, c, h, w, torch.bfloat16, "cuda")
        nbytes = 2 * x.numel() * 2
        row = {"shape": name, "alg_MB": round(nbytes / 1e6, 1)}
        for label, env in (("strip", None), ("gather", "1")):
            if env:
                if os.environ.get("BENCH_DW_SKIP_GATHER"):
                    row["gather_ms"], row["gather_GBps"] = float("nan"), float("nan")
                    continue
                os.environ["ESN_DISABLE_DW_STRIP"] = env
            for _ in range(5):
                ops.conv2d(x, prep, out=y, force_direct=True)
            ts = []
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv2d(x, prep, out=y, force_direct=True)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            os.environ.pop("ESN_DISABLE_DW_STRIP", None)
            ts.sort()
            ms = ts[len(ts) // 2]
            row[label + "_ms"] = round(ms, 4)
            row[label + "_GBps"] = round(nbytes / ms / 1e6, 1)
        row["speedup"] = round(row["gather_ms"] / row["strip_ms"], 2)
        print(json.dumps(row), flush=True)
        rows.append(row)
    if len(sys.argv) > 1:
        json.dump({"what": __doc__, "peak_GBps": 6542.1, "rows": rows}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
