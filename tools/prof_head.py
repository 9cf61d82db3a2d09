#!/usr/bin/env python
"""ERFNet's head (ConvTranspose2d(16, 19, 2, 2) + argmax) at a given input size: the tensor-core mask kernel
(esn_head_convt2x2_mask) next to the CUDA-core kernel it replaces (for timing / ncu).
    python tools/prof_head.py N H W [iters]      (H, W = size of the 16-channel feature map, e.g. 16 512 1024)
"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import ops  # noqa: E402

n, h, w = [int(v) for v in sys.argv[1:4]]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
classes = 19
torch.manual_seed(0)
x = ops.new_act(n, 16, h, w, torch.bfloat16, "cuda").normal_()
wt = torch.randn(16, classes, 2, 2, device="cuda") * 0.3
b = torch.randn(classes, device="cuda")
frags = ops.pack_convt2x2_frags(wt, classes)
packed = torch.zeros((2, 2, 16, 32), dtype=torch.float32, device="cuda")
packed[:, :, :, :classes] = wt.permute(2, 3, 0, 1)
packed = packed.contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


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


nbytes = x.numel() * 2 + n * 4 * h * w
t_new = timed(lambda: ops.head_convt2x2_mask(x, frags, b, classes))
t_old = timed(lambda: ops.head_convt2x2(x, packed, b, classes, False, True, torch.bfloat16))
m_new = ops.head_convt2x2_mask(x, frags, b, classes)
m_old = ops.head_convt2x2(x, packed, b, classes, False, True, torch.bfloat16)[1]
print("head convT2x2+argmax %dx%dx%d: mma %.4f ms (%.1f GB/s alg) | CUDA cores %.4f ms (%.1f GB/s) | masks agree %.6f"
      % (n, h, w, t_new, nbytes / t_new / 1e6, t_old, nbytes / t_old / 1e6, (m_new == m_old).float().mean().item()))
