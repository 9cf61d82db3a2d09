#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "stem" 2>&1 | tail -n 12 > gpurun_out/r02_tests_call15.log
timeout 300 python tools/bench_stem.py gpurun_out/r02_bench_stem.json > gpurun_out/r02_bench_stem.log 2>&1
timeout 600 python -m pytest tests/test_models_gpu.py tests/test_fullsize_parity_gpu.py tests/test_zz_widening_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 8 >> gpurun_out/r02_tests_call15.log
cat gpurun_out/r02_tests_call15.log; cat gpurun_out/r02_bench_stem.log
