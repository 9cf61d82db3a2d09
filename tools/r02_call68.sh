#!/usr/bin/env bash
# Round-2 call 68 (2 GPUs): gradient buckets under the deferred schedule -- the graph tests with single-GPU buckets, dp_check on 2 GPUs
set -u
P=gpurun_out
mkdir -p $P
for d in 1 0; do
  echo "== ESN_LOCAL_BUCKETS=1 ESN_DP_DEFER=$d"
  ESN_LOCAL_BUCKETS=1 ESN_DP_DEFER=$d timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "graph" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-260 | head -14
done
for d in 1 0; do
  echo "== dp_check ESN_DP_DEFER=$d"
  ESN_DP_DEFER=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$d tools/dp_check.py 2>&1 | grep -E "identical|sharded|Error|error" | cut -c1-200
done
