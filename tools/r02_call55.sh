#!/usr/bin/env bash
# Round-2 call 55: fused training close (esn_bilinear_ce, DABNet.fused_loss through GraphedTrainStep), Adam state_dict fix
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_bilinear_ce_gpu.py tests/test_optim_gpu.py -q -m gpu 2>&1 | tail -15
timeout 300 python tools/prof_bilinear_ce.py 8 19 64 128 8 2>&1 | tail -2
for flag in "" "--no-fused-loss"; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs $flag > $P/sweep/c55_default$flag.json 2> $P/sweep/c55_default$flag.err
  python tools/show_bench.py $P/sweep/c55_default$flag.json 2>/dev/null | head -1; tail -2 $P/sweep/c55_default$flag.err
done
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call55.log 2>&1
tail -5 $P/r02_tests_call55.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bilinear_ce_kernel --launch-skip 2 -c 1 -f -o $P/r02_bilinear_ce python tools/prof_bilinear_ce.py 8 19 64 128 8 3 > $P/r02_bilinear_ce.log 2>&1
tail -1 $P/r02_bilinear_ce.log
