#!/usr/bin/env bash
# Round-2 call 26: rows weight-gradient kernel v2 (register double buffering, division-free loader, 16-byte atomics, 3 CTAs/SM)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "conv_backward or wgrad_tcgen05 or (training_matches and DABNet)" > gpurun_out/r02_tests_call26.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call26.log | head -20
for cfg in "64 1" "128 1" "64 2" "32 1"; do
set -- $cfg
ESN_WGRAD_ROWS_TW=$1 ESN_WGRAD_ROWS=$2 timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_c26_$1_$2.json 2>&1 | grep "^workload"
python - <<PY
import json
d=json.load(open('gpurun_out/timeline_c26_$1_$2.json'))
for nm in ('wgrad_rows_kernel<1>','wgrad_rows_kernel<2>','wgrad_umma_kernel'):
    print('  tw=$1 mode=$2', nm, [round(r['us'],1) for r in d['first_step_sequence'] if r['name'].startswith(nm)])
PY
done
