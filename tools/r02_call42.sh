#!/usr/bin/env bash
# Round-2 call 42: stride-2 depthwise weight gradients on the register-strip kernel
set -u
mkdir -p gpurun_out/sweep
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
for wl in fastscnn_train_bf16_b16_1024x2048 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/r02_bench_$wl.json 2> gpurun_out/sweep/$wl.err
  python tools/show_bench.py gpurun_out/sweep/r02_bench_$wl.json 2>/dev/null | head -1
done
timeout 300 python tools/graph_timeline.py fastscnn_train_bf16_b16_1024x2048 gpurun_out/timeline_fastscnn_train_bf16_b16_1024x2048.json 2>&1 | grep -v Warn | head -16
