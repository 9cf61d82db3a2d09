#!/usr/bin/env bash
# Round-2 call 45: ENet's 16-channel bottleneck in one launch; smoke()
set -u
mkdir -p gpurun_out/sweep
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_fullsize_parity_gpu.py -q -m gpu -k "bottleneck4 or convt3x3s2 or ENet" 2>&1 | tail -4
timeout 600 python bench.py --workload enet_infer_bf16_b32_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/enet_bneck4.json 2> gpurun_out/sweep/enet_bneck4.err
python tools/show_bench.py gpurun_out/sweep/enet_bneck4.json 2>/dev/null | head -1; tail -2 gpurun_out/sweep/enet_bneck4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep/enet_bneck4.json'))
for k,v in list(d['kernels'].items())[:8]: print(k,v)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
