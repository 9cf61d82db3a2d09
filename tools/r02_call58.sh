#!/usr/bin/env bash
# Round-2 call 58 (2 GPUs): data-parallel correctness with the fused close, default bench on 2 GPUs (esn.optim.Adam + esn_bilinear_ce)
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > $P/c58_dp_check.log 2>&1; tail -8 $P/c58_dp_check.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
   > $P/sweep/c58_default_n2.json 2> $P/sweep/c58_default_n2.err ) 2>&1 | tail -3
python tools/show_bench.py $P/sweep/c58_default_n2.json 2>/dev/null | head -1; tail -3 $P/sweep/c58_default_n2.err
