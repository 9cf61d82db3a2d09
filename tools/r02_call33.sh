#!/usr/bin/env bash
# Round-2 call 33: Dropout2d table, bare-ReLU backward in one pass, ERFNet fused stem, bilinear backward (align_corners both)
set -u
mkdir -p gpurun_out/sweep2
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/r02_tests_call33.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call33.log | head -20
for wl in erfnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048; do
  timeout 300 python tools/graph_timeline.py $wl gpurun_out/timeline_$wl.json 2>&1 | grep -v Warn | head -16
done
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c33.json 2>&1 | grep "^workload\|bilinear"
