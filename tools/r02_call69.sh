#!/usr/bin/env bash
# Round-2 call 69: tcgen05 conv on 32-channel K blocks (Cin = 96 / 160: Fast-SCNN's bottlenecks, ESPNet) -- kernel tests, full suite, benches
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_umma_gpu.py -q -m gpu 2>&1 | tail -6
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call69.log 2>&1
tail -4 $P/r02_tests_call69.log
for wl in fastscnn_train_bf16_b16_1024x2048 fastscnn_infer_bf16_b16_1024x2048 espnet_infer_bf16_b16_1024x2048 contextnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c69_$wl.json 2> $P/sweep/c69_$wl.err
  echo "$wl: $(python tools/show_bench.py $P/sweep/c69_$wl.json 2>/dev/null | head -1)"; tail -1 $P/sweep/c69_$wl.err | cut -c1-200
done
