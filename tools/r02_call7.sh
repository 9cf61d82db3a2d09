#!/usr/bin/env bash
# 2 GPUs: the default bench line under torchrun (training + all-reduce in the captured graph, inference leg), teardown check
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ohem_gpu.py tests/test_train_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r02_tests_call7.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/r02_bench_default_n2.json 2> gpurun_out/r02_bench_default_n2.err ) 2> gpurun_out/r02_bench_default_n2.time
echo "rc=$?" >> gpurun_out/r02_bench_default_n2.time
tail -n 6 gpurun_out/r02_tests_call7.log; cat gpurun_out/r02_bench_default_n2.time; grep -i "process group\|teardown\|error" gpurun_out/r02_bench_default_n2.err | tail -5
