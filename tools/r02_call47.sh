#!/usr/bin/env bash
# Round-2 call 47: CGNet -- FGlo's global average pool accumulated by the depthwise convs (esn_dwconv_pool)
set -u
mkdir -p gpurun_out/sweep
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_fullsize_parity_gpu.py tests/test_ops_gpu.py -q -m gpu -k "CGNet or dw" 2>&1 | tail -4
for f in 1 0; do
ESN_CGNET_FUSED_POOL=$f timeout 600 python bench.py --workload cgnet_infer_bf16_b32_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > gpurun_out/sweep/cgnet_pool$f.json 2> gpurun_out/sweep/cgnet_pool$f.err
python tools/show_bench.py gpurun_out/sweep/cgnet_pool$f.json 2>/dev/null | head -1; tail -2 gpurun_out/sweep/cgnet_pool$f.err
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep/cgnet_pool1.json'))
for k,v in list(d['kernels'].items())[:8]: print(k,v)
PY
