#!/usr/bin/env bash
# Round-2 call 74: final tree -- full GPU test-suite, smoke(), ESPNet inference (80-channel inputs on 16-channel K blocks), the default
# bench line as the driver runs it
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call74.log 2>&1
tail -3 $P/r02_tests_call74.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in espnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/r02_bench_$wl.json 2> $P/sweep/$wl.err
  echo "$wl: $(python tools/show_bench.py $P/sweep/r02_bench_$wl.json 2>/dev/null | head -1)"
done
( time timeout 900 python bench.py > $P/r02_bench_default_n1.json 2> $P/r02_bench_default_n1.err ) 2> $P/r02_bench_default_n1.time
head -2 $P/r02_bench_default_n1.time | tail -1
python tools/show_bench.py $P/r02_bench_default_n1.json 2>/dev/null | head -1
