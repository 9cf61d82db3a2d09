"""TEST INFRASTRUCTURE ONLY (run as a subprocess by tests/test_abi_emulation_cpu.py, never imported by the package).

Sends every C-ABI call of a bf16 forward to the REAL library (its test build libesn_sm100_testing.so, selected through
ESN_LIB_PATH: the shipped libesn_sm100.so has no dry-run switch) with ESN_DRY_RUN=1: argument validation of every
entry point runs as on the device, and esn_conv2d_umma additionally runs its whole launch planner (tile shapes, reuse
mode, shared-memory / TMEM budget) against nominal B200 limits and returns before launching.  Return codes:
  0  = accepted (planner dry run), -4 = accepted by validation, then the launch failed because this box has no GPU;
  anything else = the library would refuse the call on the device.
Shapes only: the CPU model of the arithmetic is skipped, so full benchmark resolutions are cheap.

    ESN_DRY_RUN=1 ESN_LIB_PATH=.../esn/libesn_sm100_testing.so python tests/planner_dry_run.py NET H W [NET H W ...]   ->  one JSON object on stdout
"""
import collections
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "efficient-segmentation-networks_b200")]

import abi_emulation as A  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from esn import ops  # noqa: E402


def main(argv):
    assert os.environ.get("ESN_DRY_RUN") == "1"
    result = {}
    for i in range(0, len(argv), 3):
        name, h, w = argv[i], int(argv[i + 1]), int(argv[i + 2])
        m = build_model(name, 19).eval()
        counts, refused = collections.Counter(), []

        def probing(fn, entry, arg_refs, alg_bytes=0, flops=0, tag="", allow_unsupported=False):
            rc = fn(*arg_refs, None)
            counts["%s:%d" % (entry, rc)] += 1
            if rc not in (0, -4):
                refused.append([entry, tag, rc, bool(allow_unsupported)])
                if rc == -3 and allow_unsupported:
                    return False
            return True
        with torch.no_grad(), A.emulate_abi(bf16=True):
            ops._call = probing                       # shapes only: no arithmetic model behind the call
            ops.new_act = lambda n, c, hh, ww, dtype, device, c_alloc=None, zero=False: _alloc(n, c, hh, ww, dtype, c_alloc)
            m.predict_mask(torch.zeros(1, 3, h, w))
        result["%s@%dx%d" % (name, h, w)] = {"calls": dict(counts), "refused": refused}
    print(json.dumps(result))


def _alloc(n, c, h, w, dtype, c_alloc):
    ca = c if c_alloc is None else c_alloc
    t = torch.zeros((n, h, w, ca), dtype=dtype).permute(0, 3, 1, 2)
    return t if ca == c else t[:, :c]


if __name__ == "__main__":
    main(sys.argv[1:])
