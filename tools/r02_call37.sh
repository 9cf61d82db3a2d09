#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/sweep
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "resize_pool or ESPNet_v2 or graphed" 2>&1 | tail -3
wl=espnetv2_train_bf16_b16_1024x2048
timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/r02_bench_$wl.json 2> gpurun_out/sweep/$wl.err
python tools/show_bench.py gpurun_out/sweep/r02_bench_$wl.json 2>/dev/null | head -1; tail -2 gpurun_out/sweep/$wl.err
timeout 300 python tools/graph_timeline.py $wl gpurun_out/timeline_$wl.json 2>&1 | grep -v Warn | head -26
