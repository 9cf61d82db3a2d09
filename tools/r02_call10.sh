#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/train_grad_diag.py ENet fp32 > gpurun_out/diag_enet.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_ohem_gpu.py -q -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" > gpurun_out/r02_tests_call10.log
grep "<<<" gpurun_out/diag_enet.log | head -20; grep -n "^FAILED\|passed\|failed" gpurun_out/r02_tests_call10.log | tail
