#!/usr/bin/env bash
# Round-2 call 62: ESNet odd sizes on the device, Adam with gradient-less parameters, the two training workloads that failed in the sweep
set -u
P=gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_zz_widening_gpu.py tests/test_ops_gpu.py -q -m gpu 2>&1 | tail -6
for wl in erfnet_train_bf16_b8_512x1024 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_$wl.json 2> $P/sweep/$wl.err
  python tools/show_bench.py $P/sweep/r02_bench_$wl.json 2>/dev/null | head -1; tail -2 $P/sweep/$wl.err
done
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call62.log 2>&1
tail -3 $P/r02_tests_call62.log
