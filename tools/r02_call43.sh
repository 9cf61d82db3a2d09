#!/usr/bin/env bash
# Round-2 call 43: three-tap mma.sync weight-gradient kernel (ERFNet's 3x1 / 1x3 convs)
set -u
mkdir -p gpurun_out/sweep
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
for m in 1 2 0; do
ESN_WGRAD_TAPS3=$m timeout 300 python tools/graph_timeline.py erfnet_train_bf16_b8_512x1024 gpurun_out/timeline_erfnet_train_taps3_$m.json 2>&1 | grep "^workload\|wgrad_"
python - <<PY
import json
d=json.load(open('gpurun_out/timeline_erfnet_train_taps3_$m.json'))
for nm in ('wgrad_taps3_kernel<true, 1, 1>','wgrad_taps3_kernel<false, 1, 1>','wgrad_taps3_kernel<true, 2, 2>','wgrad_taps3_kernel<false, 2, 2>','wgrad_taps3_kernel<true, 1, 4>','wgrad_taps3_kernel<false, 1, 4>'):
    ds=[round(r['us'],1) for r in d['first_step_sequence'] if r['name'].startswith(nm)]
    if ds: print('  mode $m', nm, ds)
PY
done
