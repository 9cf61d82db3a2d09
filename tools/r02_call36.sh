#!/usr/bin/env bash
# Round-2 call 36: the default bench line as the driver runs it (N = 1, all legs) + every other workload on the current tree
set -u
mkdir -p gpurun_out/sweep
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err ) 2> gpurun_out/r02_bench_default_n1.time
tail -3 gpurun_out/r02_bench_default_n1.time
python tools/show_bench.py gpurun_out/r02_bench_default_n1.json 2>/dev/null | head -4
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err; head -c 400 gpurun_out/r02_bench_reference_cpu.json; echo
bash tools/r02_sweep.sh 2>&1 | tail -20
