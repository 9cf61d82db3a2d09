#!/usr/bin/env bash
# Round-2 call 75 (2 GPUs): the default bench line under torchrun on the final tree (uint8 training e2e, deferred all-reduce)
set -u
P=gpurun_out
mkdir -p $P/sweep
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
   > $P/r02_bench_default_n2.json 2> $P/r02_bench_default_n2.err ) 2>&1 | grep real
python tools/show_bench.py $P/r02_bench_default_n2.json 2>/dev/null | head -1; tail -3 $P/r02_bench_default_n2.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default_n2.json").read().strip().splitlines()[-1])
leg=(d.get("legs") or {}).get("erfnet_infer_bf16_b16_1024x2048") or {}
print("n_gpus", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["input"], "leg", leg.get("value"), (leg.get("e2e") or {}).get("value"))
PY
