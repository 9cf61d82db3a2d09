#!/usr/bin/env bash
# Round-2 call 29: all training tests with side-stream weight gradients, batched-load bilinear backward; timeline + bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/r02_tests_call29.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call29.log | head -20
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c29.json 2>&1 | grep -v Warn | head -34
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
python tools/show_bench.py gpurun_out/c29_bench.json 2>/dev/null | head -2; tail -3 gpurun_out/c29_bench.err
