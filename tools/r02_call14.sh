#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for g in 0 8 4 2; do
  ESN_ERF_L2_GROUP=$g python bench.py --workload erfnet_infer_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_erf_l2g$g.json 2> gpurun_out/r02_erf_l2g$g.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_erf_l2g$g.json')); print('group $g', d['value'], d['ms_per_step'], d['e2e']['value'], [ (k,v['ms']) for k,v in list(d['units'].items())[:3]])"
done
ESN_ERF_L2_GROUP=4 python -m pytest tests/test_models_gpu.py tests/test_fullsize_parity_gpu.py -q -k "ERFNet" -p no:cacheprovider 2>&1 | tail -3
