#!/usr/bin/env bash
# Round-2 call 63: fused-loss test against the fp32 gradients, the two training workloads that failed in the sweep, default line again
set -u
P=gpurun_out
mkdir -p $P/sweep
for i in 1 2 3; do timeout 900 python -m pytest tests/test_bilinear_ce_gpu.py -q -m gpu 2>&1 | tail -3; done
for wl in erfnet_train_bf16_b8_512x1024 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_$wl.json 2> $P/sweep/$wl.err
  python tools/show_bench.py $P/sweep/r02_bench_$wl.json 2>/dev/null | head -1; tail -2 $P/sweep/$wl.err
done
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c63_default.json 2> $P/sweep/c63_default.err
python tools/show_bench.py $P/sweep/c63_default.json 2>/dev/null | head -1; tail -2 $P/sweep/c63_default.err
