#!/usr/bin/env bash
# Round-2 call 27: 16-byte max-pool backward, hat-function bilinear backward, rows wgrad defaults
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "pool_bilinear or conv_backward or resize_pool or (training_matches and (DABNet or FastSCNN))" > gpurun_out/r02_tests_call27.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call27.log | head -20
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c27.json 2>&1 | grep -v Warn | head -28
