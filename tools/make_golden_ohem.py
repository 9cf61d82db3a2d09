#!/usr/bin/env python
"""Golden vectors for ProbOhemCrossEntropy2d (build container only): runs the UNMODIFIED reference class
(/root/reference/utils/losses/loss.py:163-216) in fp64 on the logits / labels of tests/golden/loss.npz and stores loss and
d loss / d logits for four (thresh, min_kept, use_weight) settings that take every branch of its forward -> tests/golden/ohem.npz.

One shim beyond the import stubs of make_golden.py: the class negates its masks as `1 - valid_mask` (loss.py:193, 209), which
worked in its pinned torch 1.1 where comparisons returned uint8 tensors; the torch in this image returns bool tensors and
refuses `1 - bool`.  For the duration of the call Tensor.__rsub__ maps `1 - <bool tensor>` to logical_not -- the torch-1.1
meaning of that expression -- and nothing else."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import GOLD, import_reference  # noqa: E402

CASES = [
    # name, thresh, min_kept, use_weight
    ("kth_above_thresh", 0.0001, 300, True),    # fewer than min_kept pixels under thresh: threshold = the 300th smallest probability (5.5e-4)
    ("thresh_wins", 0.7, 50, True),             # enough hard pixels: threshold = thresh
    ("nothing_filtered", 0.7, 5000, False),     # min_kept > number of valid pixels
    ("unweighted", 0.3, 200, False),
]


@contextlib.contextmanager
def torch11_mask_negation():
    orig = torch.Tensor.__rsub__

    def rsub(self, other):
        if self.dtype == torch.bool and other == 1:
            return ~self
        return orig(self, other)
    torch.Tensor.__rsub__ = rsub
    try:
        yield
    finally:
        torch.Tensor.__rsub__ = orig


def main():
    import_reference()
    from utils.losses.loss import ProbOhemCrossEntropy2d
    g = np.load(os.path.join(GOLD, "loss.npz"))
    lab = torch.from_numpy(g["labels"])
    out = {}
    for name, thresh, min_kept, use_weight in CASES:
        logits = torch.from_numpy(g["logits"]).clone().requires_grad_(True)
        with contextlib.redirect_stdout(io.StringIO()):          # the class prints on every call
            crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=thresh, min_kept=min_kept, use_weight=use_weight).double()
            with torch11_mask_negation():
                loss = crit(logits, lab.clone())
        loss.backward()
        out[name + "_loss"] = np.array([loss.item()])
        out[name + "_grad"] = logits.grad.numpy()
        out[name + "_cfg"] = np.array([thresh, min_kept, float(use_weight)])
        print(name, loss.item(), "kept", int((logits.grad.abs().sum(1) > 0).sum()))
    np.savez_compressed(os.path.join(GOLD, "ohem.npz"), **out)


if __name__ == "__main__":
    main()
