#!/usr/bin/env bash
# Round-2 call 64: esn_bilinear_ce on any geometry (both align_corners modes), Fast-SCNN / ESPNetv2 fused closes
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_bilinear_ce_gpu.py -q -m gpu 2>&1 | tail -8
timeout 300 python tools/prof_bilinear_ce.py 8 19 64 128 8 2>&1 | tail -1
for wl in fastscnn_train_bf16_b16_1024x2048 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_$wl.json 2> $P/sweep/$wl.err
  python tools/show_bench.py $P/sweep/r02_bench_$wl.json 2>/dev/null | head -1; tail -2 $P/sweep/$wl.err
done
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c64_default.json 2> $P/sweep/c64_default.err
python tools/show_bench.py $P/sweep/c64_default.json 2>/dev/null | head -1; tail -2 $P/sweep/c64_default.err
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call64.log 2>&1
tail -3 $P/r02_tests_call64.log
