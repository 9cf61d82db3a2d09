#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -n 5 > gpurun_out/r02_tests_call6.log
python tools/bench_dw.py gpurun_out/r02_bench_dw.json > gpurun_out/r02_bench_dw.log 2>&1
for wl in cgnet_infer_bf16_b32_1024x2048 fastscnn_infer_bf16_b16_1024x2048 espnetv2_infer_bf16_b16_1024x2048; do
  python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_${wl}_dw.json 2> gpurun_out/r02_bench_${wl}_dw.err
done
ncu --set full --clock-control none --import-source on -k regex:dw_strip_kernel -c 1 -o gpurun_out/r02_dw_strip python tools/bench_dw.py > gpurun_out/ncu_dw.log 2>&1
tail -n 3 gpurun_out/r02_tests_call6.log; cat gpurun_out/r02_bench_dw.log
