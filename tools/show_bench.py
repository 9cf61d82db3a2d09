#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (kernel and layer breakdown)."""
import json
import sys
d = json.load(open(sys.argv[1]))
print("value %.1f %s  ms/step %.3f  e2e %.1f (%.3f ms)  launches/step %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gpu_launches_per_step")))
print("clocks", d.get("clocks"))
print("roofline", {k: v for k, v in d["roofline"].items() if k != "definition"})
print("model_roofline", d.get("model_roofline"))
for k, v in d["kernels"].items():
    print("  %-28s" % k, v)
for k, v in d.get("layers", {}).items():
    print("    %-44s" % k, v)
