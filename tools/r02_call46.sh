#!/usr/bin/env bash
# Round-2 call 46: full GPU test-suite, the default bench line as the driver runs it, every other workload (final-tree sweep)
set -u
mkdir -p gpurun_out/sweep
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests_call46.log 2>&1
tail -3 gpurun_out/r02_tests_call46.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err ) 2> gpurun_out/r02_bench_default_n1.time
tail -3 gpurun_out/r02_bench_default_n1.time
python tools/show_bench.py gpurun_out/r02_bench_default_n1.json 2>/dev/null | head -3 | cut -c1-300
bash tools/r02_sweep.sh 2>&1 | tail -20
