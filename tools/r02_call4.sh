#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -n 15 > gpurun_out/r02_tests_call4.log
python tools/bench_dw.py gpurun_out/r02_bench_dw.json > gpurun_out/r02_bench_dw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dw_strip_kernel -c 1 -o gpurun_out/r02_dw_strip python tools/bench_dw.py > gpurun_out/ncu_dw.log 2>&1
tail -n 8 gpurun_out/r02_tests_call4.log; cat gpurun_out/r02_bench_dw.log
