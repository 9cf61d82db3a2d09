#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list per kernel family -> profiles/r01_traffic_<net>.json (read by bench.py's roofline.traffic).
    python tools/launches_to_traffic.py gpurun_out/r01_launches_erfnet.csv erfnet_infer_bf16_b16_1024x2048 profiles/r01_traffic_erfnet.json
"""
import collections
import csv
import json
import sys

src, workload, dst = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}
per = collections.OrderedDict()
for r in rows[1:]:
    a = per.setdefault((r[ix["ID"]], r[ix["Kernel Name"]]), {})
    a[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, k), a in per.items():
    f = fam[k.split("(")[0].split("::")[-1]]
    f[0] += 1
    f[1] += a["gpu__time_duration.sum"]
    f[2] += a["dram__bytes_read.sum"]
    f[3] += a["dram__bytes_write.sum"]
tot = sum(f[1] for f in fam.values())
out = {}
print("%-40s %4s %9s %6s %10s %10s" % ("kernel family", "n", "ms(ncu)", "share", "readMB", "writeMB"))
for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %4d %9.3f %6.3f %10.1f %10.1f" % (k, f[0], f[1] * 1e3, f[1] / tot, f[2] / 1e6, f[3] / 1e6))
    out[k] = {"launches": f[0], "ms": round(f[1] * 1e3, 4), "share": round(f[1] / tot, 4),
              "dram_read_bytes": int(f[2]), "dram_write_bytes": int(f[3])}
json.dump({"workload": workload, "source": "profiles/%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
           "dram__bytes_write.sum --clock-control none, one eager step, our kernels only)" % src.split("/")[-1],
           "families": out}, open(dst, "w"), indent=1)
