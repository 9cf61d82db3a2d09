#!/usr/bin/env bash
# Round-2 call 71: tcgen05 conv on 16-channel K blocks (Cin = 48 / 80 / 112), full suite, affected benches, the default line
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_umma_gpu.py -q -m gpu 2>&1 | tail -4
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call71.log 2>&1
tail -4 $P/r02_tests_call71.log
for wl in fastscnn_train_bf16_b16_1024x2048 fastscnn_infer_bf16_b16_1024x2048 contextnet_infer_bf16_b16_1024x2048 espnetv2_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c71_$wl.json 2> $P/sweep/c71_$wl.err
  echo "$wl: $(python tools/show_bench.py $P/sweep/c71_$wl.json 2>/dev/null | head -1)"; tail -1 $P/sweep/c71_$wl.err | cut -c1-200
done
( time timeout 900 python bench.py > $P/r02_bench_default_n1.json 2> $P/r02_bench_default_n1.err ) 2> $P/r02_bench_default_n1.time
tail -3 $P/r02_bench_default_n1.time | head -1
python tools/show_bench.py $P/r02_bench_default_n1.json 2>/dev/null | head -1
