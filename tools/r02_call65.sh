#!/usr/bin/env bash
# Round-2 call 65: fused-close model tests with dropout off (Fast-SCNN / ESPNetv2), ESPNetv2 training A/B
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_bilinear_ce_gpu.py -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --workload espnetv2_train_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs --no-fused-loss > $P/sweep/c65_espnetv2_nofuse.json 2> $P/sweep/c65_espnetv2_nofuse.err
python tools/show_bench.py $P/sweep/c65_espnetv2_nofuse.json 2>/dev/null | head -1; tail -2 $P/sweep/c65_espnetv2_nofuse.err
