#!/usr/bin/env python
"""Microbenchmark of esn_stem_conv3x3s2 (bf16 output): the mma.sync kernel (image split hi + lo / plain bf16) against the
CUDA-core kernel (ESN_STEM_FP32=1) on the stems of ERFNet (13 + 3 pool), DABNet / CGNet (32), Fast-SCNN (32, pad 0) at
16 x 1024 x 2048.  Algorithmic bytes = fp32 image read + bf16 output write.  Each variant runs in its own process (the
switches are read once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("erfnet 13+3pool", 13, 1, 1), ("dabnet 32", 32, 0, 1), ("fastscnn 32 pad0", 32, 0, 0), ("enet 13+3pool3x3", 13, 2, 1)]


def child():
    for p in (ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")):
        sys.path.insert(0, p)
    import torch
    from esn import ops
    from esn._lib import ACT_PRELU
    x = (torch.randint(0, 256, (16, 3, 1024, 2048), device="cuda").float() - 80.0).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name, cconv, pool, pad in CASES:
        ctot = cconv + (3 if pool else 0)
        w = (torch.randn(9, 3, cconv, device="cuda") * 0.2).contiguous()
        sc, sh, al = torch.rand(ctot, device="cuda"), torch.randn(ctot, device="cuda"), torch.rand(ctot, device="cuda")
        ho, wo = (1024 + 2 * pad - 3) // 2 + 1, (2048 + 2 * pad - 3) // 2 + 1
        y = ops.new_act(16, ctot, ho, wo, torch.bfloat16, "cuda")
        ts = []
        for i in range(25):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.stem_conv3x3s2(x, w, cconv, pool | (0 if pad else 256), y, sc, sh, al, ACT_PRELU)
            e1.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        nbytes = x.numel() * 4 + y.numel() * 2
        out[name] = {"ms": round(ts[len(ts) // 2], 4), "GBps": round(nbytes / ts[len(ts) // 2] / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        res = {}
        for label, env in (("mma_split", {"ESN_STEM_SPLIT": "1"}), ("mma_bf16", {"ESN_STEM_SPLIT": "0"}), ("cuda_cores", {"ESN_STEM_FP32": "1"})):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
            res[label] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}
            print(label, res[label], flush=True)
        if len(sys.argv) > 1:
            json.dump({"what": __doc__, "peak_GBps": 6542.1, "results": res}, open(sys.argv[1], "w"), indent=1)
