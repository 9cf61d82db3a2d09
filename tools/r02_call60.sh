#!/usr/bin/env bash
# Round-2 call 60 (8 GPUs): default bench on 8 GPUs -- bucket all-reduces overlapped with the backward (default), deferred behind it,
# deferred as one bucket
set -u
P=gpurun_out
mkdir -p $P/sweep
run() { # name, flags
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-legs "$@" \
      > $P/sweep/c60_$name.json 2> $P/sweep/c60_$name.err
  echo "$name: $(python tools/show_bench.py $P/sweep/c60_$name.json 2>/dev/null | head -1)"; tail -1 $P/sweep/c60_$name.err
}
run n8_overlap
ESN_DP_DEFER=1 run n8_defer
ESN_DP_DEFER=1 ESN_DP_BUCKET_BYTES=16777216 run n8_defer_one_bucket
