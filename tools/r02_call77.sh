#!/usr/bin/env bash
# Round-2 call 77: 16-byte NHWC bilinear backward, full GPU suite, ESPNetv2 training
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call77.log 2>&1
tail -4 $P/r02_tests_call77.log | cut -c1-200
timeout 600 python bench.py --workload espnetv2_train_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c77_espnetv2_train.json 2> $P/sweep/c77_espnetv2_train.err
python tools/show_bench.py $P/sweep/c77_espnetv2_train.json 2>/dev/null | head -1; tail -1 $P/sweep/c77_espnetv2_train.err | cut -c1-200
