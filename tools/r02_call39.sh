#!/usr/bin/env bash
# Round-2 call 39 (8 GPUs): the default bench line under torchrun at N = 8 and N = 4 on the final tree
set -u
mkdir -p gpurun_out
for n in 8 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 \
  > gpurun_out/r02_bench_default_n$n.json 2> gpurun_out/r02_bench_default_n$n.err ) 2> gpurun_out/r02_bench_default_n$n.time
python tools/show_bench.py gpurun_out/r02_bench_default_n$n.json 2>/dev/null | head -1
grep -c "destroyed cleanly" gpurun_out/r02_bench_default_n$n.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_default_n$n.json'))
l=d.get('legs',{}).get('erfnet_infer_bf16_b16_1024x2048')
print('n=$n leg', l and (l['value'], l['e2e']['value']))
PY
done
