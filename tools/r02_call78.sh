#!/usr/bin/env bash
# Round-2 call 78: Fast-SCNN and ERFNet training on the final tree (vector NHWC bilinear backward, 16-channel K blocks)
set -u
P=gpurun_out
mkdir -p $P/sweep
for wl in fastscnn_train_bf16_b16_1024x2048 erfnet_train_bf16_b8_512x1024; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_$wl.json 2> $P/sweep/$wl.err
  echo "$wl: $(python tools/show_bench.py $P/sweep/r02_bench_$wl.json 2>/dev/null | head -1)"
done
