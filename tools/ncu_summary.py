#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the handful of counters the roofline argument uses.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]
print("# ncu --set full --clock-control none summary of", rep)
for h, u, v in zip(hdr, units, val):
    if any(h == w or (h.startswith(w) and h[len(w):] in ("", ".per_second")) for w in want):
        print("%-72s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("\n# warp stall samples (all warps), total %d" % tot)
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print("%-28s %6d  %5.1f%%" % (h, v, 100.0 * v / max(tot, 1)))
    print("\n# SASS instructions executed (warp-level), total %d" % sum(int(r[ix["Instructions Executed"]]) for r in data))
    print("# hottest SASS lines by samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
        print("%6d  %-70s exec=%s" % (int(r[ix["# Samples"]]), r[ix["Source"]].strip()[:70], r[ix["Instructions Executed"]]))
