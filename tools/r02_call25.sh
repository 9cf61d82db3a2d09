#!/usr/bin/env bash
# Round-2 call 25: all-taps mma.sync weight-gradient kernel for dense 3x3 convs: parity, timeline, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "conv_backward or wgrad_tcgen05 or (training_matches and DABNet) or graphed" > gpurun_out/r02_tests_call25.log 2>&1
grep -n "^FAILED\|passed\|failed\|^E  " gpurun_out/r02_tests_call25.log | head -20
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c25.json 2>&1 | grep -v Warn | head -16
python - <<'PY'
import json
d=json.load(open('gpurun_out/timeline_dabnet_train_c25.json'))
for nm in ('wgrad_rows_kernel<1>','wgrad_rows_kernel<2>','wgrad_umma_kernel'):
    print(nm, [round(r['us'],1) for r in d['first_step_sequence'] if r['name'].startswith(nm)])
PY
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
python tools/show_bench.py gpurun_out/c25_bench.json 2>/dev/null | head -2
