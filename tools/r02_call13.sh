#!/usr/bin/env bash
# 8 GPUs: the default bench line (DABNet training with the in-graph bucketed all-reduce + ERFNet inference leg)
set -u
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 \
  > gpurun_out/r02_bench_default_n8.json 2> gpurun_out/r02_bench_default_n8.err ) 2> gpurun_out/r02_bench_default_n8.time
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-legs \
  > gpurun_out/r02_bench_default_n4.json 2> gpurun_out/r02_bench_default_n4.err ) 2> gpurun_out/r02_bench_default_n4.time
cat gpurun_out/r02_bench_default_n8.time; grep -c "destroyed cleanly" gpurun_out/r02_bench_default_n8.err; nvidia-smi topo -m | head -12 > gpurun_out/topo.txt
