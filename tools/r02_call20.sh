#!/usr/bin/env bash
# Round-2 call 20: where the fixed cost of the one-launch BatchNorm layer goes (atomics / barrier / cooperative launch)
set -u
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do
  ESN_BN_DBG=$d timeout 200 python tools/bench_bn.py gpurun_out/bench_bn_dbg$d.json 2>&1 | grep -v Warn
done
