#!/usr/bin/env bash
# Round-2 call 67: gradient buckets on one GPU (one multi-tensor copy instead of 61 re-layout copies) A/B, LEDNet odd sizes on the device
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_zz_widening_gpu.py tests/test_train_gpu.py -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c67_default.json 2> $P/sweep/c67_default.err
python tools/show_bench.py $P/sweep/c67_default.json 2>/dev/null | head -1; tail -1 $P/sweep/c67_default.err
ESN_NO_LOCAL_BUCKETS=1 timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c67_default_nobuckets.json 2> $P/sweep/c67_default_nobuckets.err
python tools/show_bench.py $P/sweep/c67_default_nobuckets.json 2>/dev/null | head -1; tail -1 $P/sweep/c67_default_nobuckets.err
timeout 600 python bench.py --workload erfnet_train_bf16_b8_512x1024 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c67_erfnet_train.json 2> $P/sweep/c67_erfnet_train.err
python tools/show_bench.py $P/sweep/c67_erfnet_train.json 2>/dev/null | head -1; tail -1 $P/sweep/c67_erfnet_train.err
