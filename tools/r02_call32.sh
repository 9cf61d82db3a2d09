#!/usr/bin/env bash
# Round-2 call 32: training workloads after the one-launch BatchNorm / rows wgrad / side-stream changes + full GPU test-suite
set -u
mkdir -p gpurun_out/sweep2
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests_call32.log 2>&1
tail -3 gpurun_out/r02_tests_call32.log
for wl in erfnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep2/r02_bench_$wl.json 2> gpurun_out/sweep2/$wl.err
  python tools/show_bench.py gpurun_out/sweep2/r02_bench_$wl.json 2>/dev/null | head -1
done
for wl in erfnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048; do
  timeout 300 python tools/graph_timeline.py $wl gpurun_out/timeline_$wl.json 2>&1 | grep -v Warn | head -22
done
