#!/usr/bin/env bash
# Round-2 call 49: single-thread tcgen05 issue loops (conv, pair, wgrad) + DABNet's dual / chained epilogues
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_umma_gpu.py -q -m gpu -x 2>&1 | tail -4
for cfg in "32 32 3 3 1 16 512 1024" "64 32 3 3 1 16 256 512" "128 64 3 3 1 16 128 256" "32 64 1 1 1 16 256 512" "128 128 3 1 2 16 128 256" "128 128 1 3 2 16 128 256" "64 64 1 3 1 16 256 512"; do
  timeout 120 python tools/prof_conv.py $cfg 0 20
done
timeout 2400 python -m pytest tests -q -m gpu -x > $P/r02_tests_call49.log 2>&1
tail -3 $P/r02_tests_call49.log
for w in dabnet_infer_bf16_b16_1024x2048 erfnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c49_$w.json 2> $P/sweep/c49_$w.err
  python tools/show_bench.py $P/sweep/c49_$w.json 2>/dev/null | head -1; tail -2 $P/sweep/c49_$w.err
done
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c49_default.json 2> $P/sweep/c49_default.err
python tools/show_bench.py $P/sweep/c49_default.json 2>/dev/null | head -1; tail -2 $P/sweep/c49_default.err
