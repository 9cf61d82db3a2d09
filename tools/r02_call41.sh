#!/usr/bin/env bash
# Round-2 call 41: bilinear backward with four source rows per CTA; bench roofline = dominant CUDA kernel of the replayed graph
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
for wl in dabnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048; do
timeout 300 python tools/graph_timeline.py $wl gpurun_out/timeline_$wl.json 2>&1 | grep "^workload\|bilinear_bwd_rows"
done
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
python tools/show_bench.py gpurun_out/c41_bench.json 2>/dev/null | head -3 | cut -c1-400
