#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/train_grad_diag.py ESPNet fp32 > gpurun_out/diag_espnet.log 2>&1
python tools/train_grad_diag.py ENet fp32 > gpurun_out/diag_enet.log 2>&1
python -m pytest tests/test_train_gpu.py -q -p no:cacheprovider -k "pool_unpool" 2>&1 | tail -n 20 > gpurun_out/r02_tests_call9.log
grep -c "<<<" gpurun_out/diag_espnet.log gpurun_out/diag_enet.log; tail -5 gpurun_out/r02_tests_call9.log
