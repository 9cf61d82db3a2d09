#!/usr/bin/env bash
# Every bench workload on the final tree of the round (one JSON line each -> profiles/r02_bench_<workload>.json)
set -u
mkdir -p gpurun_out/sweep
for wl in erfnet_infer_bf16_b16_1024x2048 dabnet_infer_bf16_b16_1024x2048 erfnet_infer_bf16_b16_512x1024 dabnet_infer_bf16_b16_512x1024; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-legs > gpurun_out/sweep/r02_bench_$wl.json 2> gpurun_out/sweep/$wl.err
done
for wl in enet_infer_bf16_b32_1024x2048 cgnet_infer_bf16_b32_1024x2048 fastscnn_infer_bf16_b16_1024x2048 espnet_infer_bf16_b16_1024x2048 \
          espnetv2_infer_bf16_b16_1024x2048 esnet_infer_bf16_b16_1024x2048 contextnet_infer_bf16_b16_1024x2048 edanet_infer_bf16_b16_1024x2048 \
          lednet_infer_bf16_b16_1024x2048 erfnet_train_bf16_b8_512x1024 fastscnn_train_bf16_b16_1024x2048 espnetv2_train_bf16_b16_1024x2048; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/r02_bench_$wl.json 2> gpurun_out/sweep/$wl.err
done
for f in gpurun_out/sweep/*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['config']['workload'], d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('gpu_eager_baseline') or {}).get('speedup_vs_fastest'))
except Exception as e: print('$f', 'ERR', e)
"; done
