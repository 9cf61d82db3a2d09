#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -n 5 > gpurun_out/r02_tests_call5.log
ESN_DW_OCC=3 python tools/bench_dw.py gpurun_out/r02_bench_dw.json > gpurun_out/r02_bench_dw.log 2>&1
BENCH_DW_SKIP_GATHER=1 ESN_DW_OCC=2 python tools/bench_dw.py > gpurun_out/r02_bench_dw_occ2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dw_strip_kernel -c 1 -o gpurun_out/r02_dw_strip python tools/bench_dw.py > gpurun_out/ncu_dw.log 2>&1
tail -n 3 gpurun_out/r02_tests_call5.log; cat gpurun_out/r02_bench_dw.log; echo OCC2; cat gpurun_out/r02_bench_dw_occ2.log
