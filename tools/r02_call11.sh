#!/usr/bin/env bash
# Full GPU suite after the train-mode / OHEM / augmentation work, default bench line, ncu launch lists of both legs
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED" | tail -n 120 > gpurun_out/r02_tests_call11.log
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r02_launches_dabnet_train.csv python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu_train.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_erfnet.csv python bench.py --workload erfnet_infer_bf16_b16_1024x2048 --steps 1 --warmup 3 \
  --no-graph --no-cpu-baseline --no-gpu-eager > gpurun_out/ncu_erfnet.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r02_tests_call11.log | tail; tail -c 600 gpurun_out/r02_bench_default.err
