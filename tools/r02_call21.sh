#!/usr/bin/env bash
# Round-2 call 21: one-launch BatchNorm, second version (replicated sums, predicated unrolled loads, variants), tests + sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "bn_act or training_matches_reference_fp64 or graphed or eval_after" > gpurun_out/r02_tests_call21.log 2>&1
tail -5 gpurun_out/r02_tests_call21.log
for v in 0 1 2 3; do
  BN_ONLY_FUSED=1 ESN_BN_FWD_VARIANT=$v ESN_BN_BWD_VARIANT=$v timeout 200 python tools/bench_bn.py gpurun_out/bench_bn_v$v.json 2>&1 | grep -v Warn
done
timeout 300 python tools/graph_timeline.py dabnet_train_bf16_b8_512x1024 gpurun_out/timeline_dabnet_train_c21.json 2>&1 | grep -v Warn | head -40
