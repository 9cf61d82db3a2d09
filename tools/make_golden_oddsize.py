#!/usr/bin/env python
"""tests/golden/oddsize.npz from the UNMODIFIED reference (build container only): ESNet and LEDNet on inputs whose height /
width are odd at one or more levels, i.e. the F.pad path of their DownsamplerBlock (model/ESNet.py:22-29, LEDNet.py:76-96).  Same import stubs and seeded
fixture weights as tools/make_golden.py.

    python tools/make_golden_oddsize.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
from make_golden import GOLD, import_reference  # noqa: E402

SIZES = ((1, 51, 77), (2, 36, 50), (1, 33, 64))     # odd at level 1 / 2 / 3; odd at level 2 and 3 only; odd height only


def main():
    from oracle import fixture
    build_model, _ = import_reference()
    torch.set_num_threads(8)
    torch.manual_seed(1234)
    out = {}
    for name in ("ESNet", "LEDNet"):      # LEDNet: the same padded DownsamplerBlock (LEDNet.py:76-96) + the attention pyramid's
        torch.manual_seed(1234)           # own rounding of its three levels (LEDNet.py:245-264)
        m = build_model(name, 19)
        m.load_state_dict(fixture.randomize_state_dict(m.state_dict(), 1234))
        m.eval()
        with torch.no_grad():
            for n, h, w in SIZES:
                y = m(fixture.make_input(n, h, w, 1234))
                out["%s_%dx%dx%d_logits" % (name, n, h, w)] = y.numpy()
                print(name, (n, h, w), "->", tuple(y.shape))
    np.savez_compressed(os.path.join(GOLD, "oddsize.npz"), **out)


if __name__ == "__main__":
    main()
