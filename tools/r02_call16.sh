#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "stem" 2>&1 | tail -n 4 > gpurun_out/r02_tests_call16.log
timeout 300 python tools/bench_stem.py gpurun_out/r02_bench_stem.json > gpurun_out/r02_bench_stem.log 2>&1
cat gpurun_out/r02_tests_call16.log; cat gpurun_out/r02_bench_stem.log
