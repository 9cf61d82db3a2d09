#!/usr/bin/env bash
# Round-2 call 57: esn_bilinear_ce v3 (onehot per label run through shared atomics, reduce-scatter over the four source pixels)
# two CTAs per SM), CGNet.fused_loss
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_bilinear_ce_gpu.py -q -m gpu 2>&1 | tail -8
timeout 300 python tools/prof_bilinear_ce.py 8 19 64 128 8 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c57_default.json 2> $P/sweep/c57_default.err
python tools/show_bench.py $P/sweep/c57_default.json 2>/dev/null | head -1; tail -2 $P/sweep/c57_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bilinear_ce_kernel --launch-skip 2 -c 1 -f -o $P/r02_bilinear_ce_v3 python tools/prof_bilinear_ce.py 8 19 64 128 8 3 > $P/r02_bilinear_ce_v3.log 2>&1
tail -1 $P/r02_bilinear_ce_v3.log
