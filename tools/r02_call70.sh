#!/usr/bin/env bash
# Round-2 call 70: Fast-SCNN bf16 gradient check at batch 4 against the fp64 oracle; full GPU suite on the final tree
set -u
P=gpurun_out
mkdir -p $P
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu -k "FastSCNN" -s 2>&1 | grep -E "FastSCNN bf16|passed|failed|^E  " | cut -c1-300
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call70.log 2>&1
tail -3 $P/r02_tests_call70.log
