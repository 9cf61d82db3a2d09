#!/usr/bin/env bash
# Round-2 call 30 (2 GPUs): default bench under torchrun with the cooperative BatchNorm kernels, the side-stream weight
# gradients and the NCCL bucket all-reduces in one captured graph; data-parallel gradient equivalence check; teardown
set -u
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/c30_bench_n2.json 2> gpurun_out/c30_bench_n2.err ) 2> gpurun_out/c30_bench_n2.time
echo "rc=$?" >> gpurun_out/c30_bench_n2.time
cat gpurun_out/c30_bench_n2.time; grep -i "process group\|teardown\|error\|watchdog" gpurun_out/c30_bench_n2.err | tail -5
python tools/show_bench.py gpurun_out/c30_bench_n2.json 2>/dev/null | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/c30_dp_check.log 2>&1; tail -6 gpurun_out/c30_dp_check.log
