#!/usr/bin/env python
"""Blind-debug helper for the tcgen05 conv: runs every case of tests/test_umma_gpu.py in
isolation (no epilogue first), prints error structure.  Usage (GPU box):
    python tools/umma_diag.py [case_name ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200"), os.path.join(ROOT, "tests")]
from test_umma_gpu import CASES, run_case  # noqa: E402


def main():
    want = set(sys.argv[1:])
    for case in CASES:
        if want and case[0] not in want:
            continue
        for ep in (False, True):
            try:
                err, y, ref = run_case(case, ep)
            except Exception as e:  # noqa
                print("%-16s epilogue=%d EXC %r" % (case[0], ep, e), flush=True)
                continue
            status = "ok " if (err == err and err < 1.5e-2) else "BAD"
            print("%-16s epilogue=%d %s max-rel-err %.3e" % (case[0], ep, status, err), flush=True)
            if status == "BAD":
                d = (y.float() - ref).abs()
                nanfrac = torch.isnan(y.float()).float().mean().item()
                bad = (d > 2e-2 * ref.abs().max()) | torch.isnan(d)
                print("   nan-frac %.3f  bad-frac %.3f" % (nanfrac, bad.float().mean().item()))
                print("   bad per channel (first 16):", bad.float().mean(dim=(0, 2, 3))[:16].tolist())
                print("   bad per row h (first 16):  ", bad.float().mean(dim=(0, 1, 3))[:16].tolist())
                print("   bad per col w (first 24):  ", bad.float().mean(dim=(0, 1, 2))[:24].tolist())
                print("   y[0,:8,0,0]  ", y[0, :8, 0, 0].float().tolist())
                print("   ref[0,:8,0,0]", ref[0, :8, 0, 0].tolist())
                print("   y[0,:8,1,5]  ", y[0, :8, 1, 5].float().tolist())
                print("   ref[0,:8,1,5]", ref[0, :8, 1, 5].tolist())


if __name__ == "__main__":
    main()
