#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 6 > gpurun_out/r02_tests_call17.log
for wl in fastscnn_infer_bf16_b16_1024x2048 dabnet_infer_bf16_b16_1024x2048 cgnet_infer_bf16_b32_1024x2048 contextnet_infer_bf16_b16_1024x2048; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_${wl}_stem.json 2> gpurun_out/r02_bench_${wl}_stem.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${wl}_stem.json')); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
cat gpurun_out/r02_tests_call17.log
