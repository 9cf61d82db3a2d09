#!/usr/bin/env bash
# First GPU call of round 2: everything that was built after the round-1 GPU budget was
# spent, in the order "tests, then bench lines, then profiler lists".  Run from the repo root through gpurun:
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy what is to be judged into profiles/ as r02_*).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_pipeline_gpu.py tests/test_zz_widening_gpu.py -q -rxX 2>&1 | tee gpurun_out/r02_pending_tests.log
# default workload incl. the guarded e2e_u8 leg, then the uint8 pipeline as the timed e2e input
python bench.py > gpurun_out/r02_bench_erfnet.json 2> gpurun_out/r02_bench_erfnet.err
python bench.py --e2e-input u8 --no-cpu-baseline > gpurun_out/r02_bench_erfnet_u8.json 2>> gpurun_out/r02_bench_erfnet.err
for wl in esnet edanet lednet contextnet; do
  python bench.py --workload ${wl}_infer_bf16_b16_1024x2048 --no-cpu-baseline \
    > gpurun_out/r02_bench_${wl}_infer_bf16_b16_1024x2048.json 2> gpurun_out/r02_bench_${wl}.err
done
# launch lists (never a bench value): only after the plain runs above exited
for wl in esnet edanet lednet contextnet; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_launches_${wl}.csv python bench.py --workload ${wl}_infer_bf16_b16_1024x2048 \
    --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-u8-leg > /dev/null 2>&1
done
tail -n 5 gpurun_out/r02_pending_tests.log
