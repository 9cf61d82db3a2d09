#!/usr/bin/env bash
# Round-2 call 22: training tests on the one-launch BatchNorm (v2) + concat-gradient fixes, timeline, default bench line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r02_tests_call22.log 2>&1
tail -5 gpurun_out/r02_tests_call22.log
BN_ONLY_FUSED=1 timeout 200 python tools/bench_bn.py gpurun_out/bench_bn_final.json 2>&1 | grep -v Warn
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c22.json 2>&1 | grep -v Warn | head -24
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
python tools/show_bench.py gpurun_out/c22_bench.json 2>/dev/null | head -5 || head -c 600 gpurun_out/c22_bench.json
