#!/usr/bin/env bash
# Round-2 call 51: row tiles with tap rows (c128 -> 64 3x3), four staging buffers with a residual, thread issue by default
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_umma_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/conv_ab.py ESN_UMMA_NS=2 ESN_UMMA_NOHROWS=1 2>&1 | tee $P/r02_conv_ab.txt
timeout 300 python tools/layer_times.py DABNet 16 1024 2048 > $P/r02_layer_times_dabnet.txt 2>&1; grep -v "umma\|dab_dw" $P/r02_layer_times_dabnet.txt | tail -25
timeout 2400 python -m pytest tests -q -m gpu -x > $P/r02_tests_call51.log 2>&1
tail -3 $P/r02_tests_call51.log
for w in dabnet_infer_bf16_b16_1024x2048 erfnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c51_$w.json 2> $P/sweep/c51_$w.err
  python tools/show_bench.py $P/sweep/c51_$w.json 2>/dev/null | head -1; tail -2 $P/sweep/c51_$w.err
done
