#!/usr/bin/env bash
# Round-2 call 40: rows weight-gradient kernel with the input-row ring; fills removed from DABNet's training forward
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c40.json 2>&1 | grep -v Warn | head -14
python - <<'PY'
import json
d=json.load(open('gpurun_out/timeline_dabnet_train_c40.json'))
for nm in ('wgrad_rows_kernel<1>','wgrad_rows_kernel<2>','vectorized_elementwise_kernel<8, FillFunctor'):
    print(nm, [round(r['us'],1) for r in d['first_step_sequence'] if r['name'].startswith(nm)])
PY
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.err
python tools/show_bench.py gpurun_out/c40_bench.json 2>/dev/null | head -1
