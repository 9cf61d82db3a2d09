#!/usr/bin/env bash
# Round-2 call 35: bilinear backward with shared vertical weights, streaming bare-ReLU backward; tests, timelines, benches
set -u
mkdir -p gpurun_out/sweep2
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/r02_tests_call35.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call35.log | head -20
for wl in erfnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048 dabnet_train_bf16_b8_512x1024; do
  timeout 300 python tools/graph_timeline.py $wl gpurun_out/timeline_$wl.json 2>&1 | grep -v Warn | head -12
done
