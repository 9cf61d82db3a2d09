#!/usr/bin/env bash
# Round-2 call 54: esn.optim.Adam with double-precision beta complements, ERFNet / ESNet mask head on mma.sync
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_optim_gpu.py -q -m gpu 2>&1 | tail -5
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/prof_head.py 16 512 1024 2>&1 | tail -2
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call54.log 2>&1
tail -5 $P/r02_tests_call54.log
for w in erfnet_infer_bf16_b16_1024x2048 esnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c54_$w.json 2> $P/sweep/c54_$w.err
  python tools/show_bench.py $P/sweep/c54_$w.json 2>/dev/null | head -1; tail -2 $P/sweep/c54_$w.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_convt2x2_mask --launch-skip 2 -c 1 -f -o $P/r02_head_t2_mask python tools/prof_head.py 16 512 1024 3 > $P/r02_head_t2_mask.log 2>&1
tail -1 $P/r02_head_t2_mask.log
