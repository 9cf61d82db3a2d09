#!/usr/bin/env python
"""Per-kernel device timeline of the GRAPH-REPLAYED step of a bench workload (CUPTI activity records through torch.profiler;
no kernel replay, caches warm, real overlap) -- what the eager per-launch event timings and the ncu launch list (cold caches,
serialised) cannot show: where the replayed step's time goes, kernel by kernel, and how much of it is gaps between kernels.

    python tools/graph_timeline.py [workload] [out.json]         # a diagnostic: the profiler perturbs the step slightly,
                                                                 # numbers printed here are never bench values
"""
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
import bench  # noqa: E402
from builders.model_builder import build_model  # noqa: E402
from oracle import fixture  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "dabnet_train_bf16_b8_512x1024"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/timeline_%s.json" % wl
name, batch, H, W, mode = bench.WORKLOADS[wl]
m = build_model(name, 19)
m.load_state_dict(bench.fixture_state_dict(name))
m = m.cuda()
x = fixture.make_input(batch, H, W).cuda()
if mode == "train":
    from esn.graph import GraphedTrainStep
    from utils.losses.loss import CrossEntropyLoss2d
    m.train()
    crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
    opt = torch.optim.Adam(m.parameters(), lr=5e-4, weight_decay=1e-4, fused=True, capturable=True)
    y = fixture.make_labels(batch, H, W, 19).cuda()
    g = GraphedTrainStep(m, crit, opt, x, y, warmup=3).graph
else:
    m.eval()

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return m.predict_mask(x)
    for _ in range(3):
        step()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        with torch.cuda.graph(g, stream=s):
            step()
    torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 5

REPS = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() >= 0]
ev.sort(key=lambda e: e.time_range.start)
per = defaultdict(lambda: [0, 0.0])
busy = 0.0
gap = 0.0
last_end = None
rows = []
for e in ev:
    d = e.time_range.elapsed_us()
    short = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("at::native::", "")
    short = short.split("(")[0][:110]
    per[short][0] += 1
    per[short][1] += d
    st, en = e.time_range.start, e.time_range.end
    if last_end is not None and st > last_end:
        gap += st - last_end
    if last_end is None or en > last_end:
        busy += en - max(st, last_end if last_end is not None else st)
        last_end = en
    rows.append((st, d, short))
span = (ev[-1].time_range.end - ev[0].time_range.start) if ev else 0
tab = sorted(((k, v[0] / REPS, v[1] / REPS / 1e3) for k, v in per.items()), key=lambda r: -r[2])
print("workload %s: replay %.3f ms unprofiled; profiled span %.3f ms/step, union of kernel time %.3f, gaps %.3f, kernels/step %.0f"
      % (wl, plain_ms, span / REPS / 1e3, busy / REPS / 1e3, gap / REPS / 1e3, len(ev) / REPS))
for k, n, ms in tab[:45]:
    print("%9.4f ms %6.1f x  %s" % (ms, n, k))
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"workload": wl, "replay_ms_unprofiled": plain_ms, "span_ms": span / REPS / 1e3, "busy_ms": busy / REPS / 1e3,
           "gap_ms": gap / REPS / 1e3, "kernels_per_step": len(ev) / REPS,
           "kernels": [{"name": k, "launches": n, "ms": ms} for k, n, ms in tab],
           "first_step_sequence": [{"t_us": r[0] - rows[0][0], "us": r[1], "name": r[2]} for r in rows[:len(rows) // REPS]]},
          open(out, "w"), indent=0)
