#!/usr/bin/env bash
# Round-2 call 79: vector bilinear backward only up to an up-sampling factor of 4; op tests, Fast-SCNN training
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "resize_pool or FastSCNN or ESPNet_v2" 2>&1 | tail -2
timeout 300 python bench.py --workload fastscnn_train_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_fastscnn_train_bf16_b16_1024x2048.json 2> $P/sweep/fastscnn_train.err
echo "fastscnn_train: $(python tools/show_bench.py $P/sweep/r02_bench_fastscnn_train_bf16_b16_1024x2048.json 2>/dev/null | head -1)"
