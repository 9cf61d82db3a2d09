#!/usr/bin/env bash
# Round-2 GPU call 2: GPU suite after the LEDNet / cache / loss changes, the restructured bench.py (default = DABNet training
# with the ERFNet inference leg), secondary workloads with unit rooflines.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | grep -v "^PASSED" | tail -n 80 > gpurun_out/r02_tests_call2.log
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
for wl in dabnet_infer_bf16_b16_1024x2048 cgnet_infer_bf16_b32_1024x2048 fastscnn_infer_bf16_b16_1024x2048 espnetv2_infer_bf16_b16_1024x2048; do
  python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err
done
tail -n 30 gpurun_out/r02_tests_call2.log; tail -n 5 gpurun_out/r02_bench_default.err
