#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/bn_fused_diag2.py FastSCNN 2>&1 | grep "^loss"
timeout 1200 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/r02_tests_call24.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call24.log | head -20
BN_ONLY_FUSED=1 timeout 200 python tools/bench_bn.py gpurun_out/bench_bn_final.json 2>&1 | grep -v Warn
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err
python tools/show_bench.py gpurun_out/c24_bench.json 2>/dev/null | head -2
