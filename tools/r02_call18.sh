#!/usr/bin/env bash
# Round-2 call 18: fused one-launch BatchNorm (cooperative grid + barrier) and stride-2 weight gradients by parity on tcgen05:
# tests, then DABNet training with each switch on / off.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "bn_act or conv_backward or training_matches_reference_fp64 or graphed" > gpurun_out/r02_tests_call18.log 2>&1
tail -5 gpurun_out/r02_tests_call18.log
B="python bench.py --workload dabnet_train_bf16_b8_512x1024 --no-cpu-baseline --no-gpu-eager --no-legs"
timeout 300 $B > gpurun_out/c18_fused_s2.json 2> gpurun_out/c18_fused_s2.err
ESN_FUSED_BN=0 timeout 300 $B > gpurun_out/c18_unfused_s2.json 2> gpurun_out/c18_unfused_s2.err
ESN_WGRAD_S2_PARITY=0 timeout 300 $B > gpurun_out/c18_fused_wmma.json 2> gpurun_out/c18_fused_wmma.err
for f in gpurun_out/c18_*.json; do python -c "
import json
try:
    d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
    for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms'])[:14]: print('   ', k, v['launches'], v['ms'])
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/c18_*.err
