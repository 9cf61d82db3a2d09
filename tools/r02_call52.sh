#!/usr/bin/env bash
# Round-2 call 52: staging buffers handed back early (residual producer runs ahead), DAB depthwise pair with packed bf16
# stage-1 values, esn_concat_tail (no concat zero fill), ncu of the 3x3 c32 conv on the one-thread issue loop
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python tools/conv_ab.py ESN_UMMA_NS=2 2>&1 | tee $P/r02_conv_ab2.txt
timeout 300 python tools/layer_times.py DABNet 16 1024 2048 > $P/r02_layer_times_dabnet.txt 2>&1; grep -v "umma" $P/r02_layer_times_dabnet.txt | tail -25
timeout 2400 python -m pytest tests -q -m gpu -x > $P/r02_tests_call52.log 2>&1
tail -3 $P/r02_tests_call52.log
for w in dabnet_infer_bf16_b16_1024x2048 erfnet_infer_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c52_$w.json 2> $P/sweep/c52_$w.err
  python tools/show_bench.py $P/sweep/c52_$w.json 2>/dev/null | head -1; tail -2 $P/sweep/c52_$w.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma --launch-skip 2 -c 1 -f -o $P/r02_conv3x3_thread python tools/prof_conv.py 32 32 3 3 1 16 512 1024 0 3 > $P/r02_conv3x3_thread.log 2>&1
tail -1 $P/r02_conv3x3_thread.log
