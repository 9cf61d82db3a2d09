#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_train_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 8 > gpurun_out/r02_tests_call12.log
python bench.py --no-legs --no-cpu-baseline --no-gpu-eager > gpurun_out/r02_bench_train_arena.json 2> gpurun_out/r02_bench_train_arena.err
tail -n 4 gpurun_out/r02_tests_call12.log; tail -c 300 gpurun_out/r02_bench_train_arena.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_train_arena.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
