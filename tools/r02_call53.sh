#!/usr/bin/env bash
# Round-2 call 53: esn.optim.Adam (one-launch optimizer step), 16-byte esn_concat_tail, DAB pair back on fp32 stage-1 rows
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -x -k "graph" 2>&1 | tail -5
for flag in "" "--torch-adam"; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs $flag > $P/sweep/c53_default$flag.json 2> $P/sweep/c53_default$flag.err
  python tools/show_bench.py $P/sweep/c53_default$flag.json 2>/dev/null | head -1; tail -2 $P/sweep/c53_default$flag.err
done
timeout 300 python tools/layer_times.py DABNet 16 1024 2048 > $P/r02_layer_times_dabnet_c53.txt 2>&1; grep -v "umma" $P/r02_layer_times_dabnet_c53.txt | tail -25
timeout 600 python bench.py --workload dabnet_infer_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c53_dabnet_infer.json 2> $P/sweep/c53_dabnet_infer.err
python tools/show_bench.py $P/sweep/c53_dabnet_infer.json 2>/dev/null | head -1; tail -2 $P/sweep/c53_dabnet_infer.err
timeout 2400 python -m pytest tests -q -m gpu -x > $P/r02_tests_call53.log 2>&1
tail -3 $P/r02_tests_call53.log
