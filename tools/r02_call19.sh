#!/usr/bin/env bash
# Round-2 call 19: device timeline of the graph-replayed DABNet training step (CUPTI through torch.profiler), fused / unfused BN
set -u
mkdir -p gpurun_out
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_fused.json > gpurun_out/timeline_fused.txt 2>&1
ESN_FUSED_BN=0 timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_unfused.json > gpurun_out/timeline_unfused.txt 2>&1
head -50 gpurun_out/timeline_fused.txt; head -30 gpurun_out/timeline_unfused.txt
