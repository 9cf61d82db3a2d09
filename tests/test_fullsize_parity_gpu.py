"""Oracle parity AT THE BENCHMARKED SHAPE (BASELINE.json configs[1]: 1024x2048; configs[2]: 512x1024): the CUDA path through
the C ABI against the CPU oracle on one full-size image -- the shapes at which the fused factorized-pair kernel (row widths
512 / 1024), the row-reuse modes and the 128-channel convs at 128x256 actually run in bench.py.  fp32: logits within 1e-3
relative, argmax >= 99.9 %.  bf16: logits within 5e-2 relative; argmax raw and margin-aware (random-init logits have tiny
top-2 margins), with torch's own bf16 autocast of the same graph as the noise floor."""
import json
import os

import pytest

import parity_util as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,h,w", [("ERFNet", 1024, 2048), ("DABNet", 1024, 2048), ("ERFNet", 512, 1024),
                                      ("DABNet", 512, 1024)])
def test_full_size_oracle_parity(name, h, w, spec):
    r = P.measure(name, spec, 1, h, w)
    print(json.dumps(r))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_fullsize_%s_%dx%d.json" % (name, h, w)), "w") as f:
            json.dump(r, f, indent=1)
    assert r["fp32_rel_l2"] < 1e-3 and r["fp32_max_rel"] < 1e-3, r
    assert r["fp32_argmax_raw"] >= 0.999, r
    assert r["fp32_fused_argmax_equals_argmax_of_logits"]
    assert r["bf16_rel_l2"] < max(5e-2, 1.5 * r["torch_autocast_rel_l2"]), r
    assert r["bf16_argmax_margin_aware"] >= min(0.999, r["torch_autocast_argmax_margin_aware"] - 5e-3), r
    assert r["bf16_argmax_raw"] >= min(0.99, r["torch_autocast_argmax_raw"] - 0.02), r
