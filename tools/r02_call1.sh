#!/usr/bin/env bash
# Round-2 GPU call 1: the whole GPU suite without xfail blankets, LEDNet layer diagnosis, achieved parity of all nets,
# bench lines with the GPU-eager reference leg.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider 2>&1 | tail -n 150 > gpurun_out/r02_tests_call1.log
python tools/diag_layers.py LEDNet 128 256 > gpurun_out/r02_diag_lednet.log 2>&1
python tools/diag_layers.py LEDNet 512 1024 >> gpurun_out/r02_diag_lednet.log 2>&1
python tools/parity_report.py gpurun_out/r02_parity_512x1024.json 512 1024 > gpurun_out/r02_parity.log 2>&1
python bench.py > gpurun_out/r02_bench_erfnet_infer_bf16_b16_1024x2048.json 2> gpurun_out/r02_bench_erfnet.err
python bench.py --workload dabnet_train_bf16_b8_512x1024 > gpurun_out/r02_bench_dabnet_train_bf16_b8_512x1024.json 2> gpurun_out/r02_bench_dabnet_train.err
python bench.py --workload dabnet_infer_bf16_b16_1024x2048 --no-cpu-baseline > gpurun_out/r02_bench_dabnet_infer_bf16_b16_1024x2048.json 2> gpurun_out/r02_bench_dabnet.err
for wl in erfnet_infer_bf16_b16_512x1024 dabnet_infer_bf16_b16_512x1024; do
  python bench.py --impl reference-gpu --workload $wl --steps 20 --warmup 10 > gpurun_out/r02_refgpu_$wl.json 2> gpurun_out/r02_refgpu_$wl.err
done
for wl in esnet edanet lednet contextnet; do
  python bench.py --workload ${wl}_infer_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager \
    > gpurun_out/r02_bench_${wl}_infer_bf16_b16_1024x2048.json 2> gpurun_out/r02_bench_${wl}.err
done
tail -n 40 gpurun_out/r02_tests_call1.log
