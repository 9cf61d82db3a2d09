#!/usr/bin/env bash
# Round-2 call 73: ESPNetv2 training (147 -> 192 padding) -- the odd-channel 1x1 convs onto the class scores over zero-padded widths on the tensor cores
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "ESPNet_v2" -s 2>&1 | grep -E "ESPNet_v2 torch|autocast|passed|failed|^E  " | cut -c1-260
timeout 600 python bench.py --workload espnetv2_train_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c73_espnetv2_train.json 2> $P/sweep/c73_espnetv2_train.err
python tools/show_bench.py $P/sweep/c73_espnetv2_train.json 2>/dev/null | head -1; tail -2 $P/sweep/c73_espnetv2_train.err | cut -c1-200
