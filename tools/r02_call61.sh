#!/usr/bin/env bash
# Round-2 call 61: final-tree sweep -- full GPU test-suite, the default bench line as the driver runs it (and the reference arms),
# every other workload
set -u
mkdir -p gpurun_out/sweep
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02_tests_call61.log 2>&1
tail -3 gpurun_out/r02_tests_call61.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err ) 2> gpurun_out/r02_bench_default_n1.time
tail -3 gpurun_out/r02_bench_default_n1.time
python tools/show_bench.py gpurun_out/r02_bench_default_n1.json 2>/dev/null | head -3 | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err; cut -c1-300 gpurun_out/r02_bench_reference_cpu.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
bash tools/r02_sweep.sh 2>&1 | tail -20
