#!/usr/bin/env python
"""Achieved parity numbers of every net on the GPU (tests/parity_util.measure) -> one JSON file (profiles/r02_parity.json
is a copy of a run on B200).  Usage: python tools/parity_report.py OUT.json [H W]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import parity_util as P  # noqa: E402

NETS = ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet", "ESNet", "ContextNet", "EDANet", "LEDNet"]


def main():
    out = sys.argv[1]
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 1024)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    rows = []
    for name in NETS:
        try:
            r = P.measure(name, spec, 1, h, w)
        except Exception as exc:  # noqa: BLE001
            r = {"net": name, "error": repr(exc)[:300]}
        print(json.dumps(r), flush=True)
        rows.append(r)
    json.dump({"device": torch.cuda.get_device_name(0), "oracle": "oracle/nets.py fp32 on the host CPU",
               "margin_rule": "pixels whose oracle top-2 logit margin exceeds 5e-2 of the top logit", "rows": rows},
              open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
