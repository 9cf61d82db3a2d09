#!/usr/bin/env bash
# Round-2 call 44: ENet head (transposed conv 3x3/s2 + argmax) as one tensor-core launch
set -u
mkdir -p gpurun_out/sweep
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -m gpu -k "convt3x3s2 or ENet" 2>&1 | tail -4
for f in 1 0; do
ESN_ENET_FUSED_HEAD=$f timeout 600 python bench.py --workload enet_infer_bf16_b32_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/enet_fused$f.json 2> gpurun_out/sweep/enet_fused$f.err
python tools/show_bench.py gpurun_out/sweep/enet_fused$f.json 2>/dev/null | head -1; tail -2 gpurun_out/sweep/enet_fused$f.err
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep/enet_fused1.json'))
for k,v in list(d['kernels'].items())[:8]: print(k,v)
PY
