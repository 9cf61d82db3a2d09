#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for d in fwd bwd; do
echo "== fused only $d"; ESN_FUSED_BN_DIR=$d timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "training_matches_reference_fp64 and FastSCNN and dtype1" 2>&1 | grep -n "FastSCNN torch\|passed\|failed"
done
for v in 1 2; do
echo "== bwd variant $v"; ESN_BN_BWD_VARIANT=$v timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "training_matches_reference_fp64 and FastSCNN and dtype1" 2>&1 | grep -n "FastSCNN torch\|passed\|failed"
done
