#!/usr/bin/env bash
# Round-2 call 31 (2 GPUs): all-reduce from inside the backward vs after it, with the cooperative BatchNorm kernels
set -u
mkdir -p gpurun_out
for d in 0 1; do
ESN_DP_DEFER=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$d bench.py --gpus 2 --steps 20 --warmup 5 --no-legs \
  > gpurun_out/c31_bench_n2_defer$d.json 2> gpurun_out/c31_bench_n2_defer$d.err
python tools/show_bench.py gpurun_out/c31_bench_n2_defer$d.json 2>/dev/null | head -1
done
timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'])"
