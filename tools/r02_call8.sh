#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_train_gpu.py tests/test_ohem_gpu.py -q -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" | tail -n 150 > gpurun_out/r02_tests_call8.log
tail -n 60 gpurun_out/r02_tests_call8.log
