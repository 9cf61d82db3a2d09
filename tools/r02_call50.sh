#!/usr/bin/env bash
# Round-2 call 50: A/B of the two tcgen05 issue forms on one box (warp-uniform loop vs one thread), DABNet / ERFNet benches
set -u
P=gpurun_out
mkdir -p $P/sweep
for cfg in "32 32 3 3 1 16 512 1024" "64 32 3 3 1 16 256 512" "128 64 3 3 1 16 128 256" "32 64 1 1 1 16 256 512" "64 128 1 1 1 16 128 256" "128 128 3 1 1 16 128 256" "128 128 3 1 2 16 128 256" "128 128 1 3 1 16 128 256" "128 128 1 3 2 16 128 256" "64 64 1 3 1 16 256 512" "64 64 3 1 1 16 256 512" "16 16 3 1 1 16 512 1024"; do
  for m in warp thread; do
    ESN_UMMA_ISSUE=$m timeout 120 python tools/prof_conv.py $cfg 0 30 | sed "s/^/$m  /"
    ESN_UMMA_ISSUE=$m timeout 120 python tools/prof_conv.py $cfg 1 30 | sed "s/^/$m  /"
  done
done
for w in dabnet_infer_bf16_b16_1024x2048 erfnet_infer_bf16_b16_1024x2048 dabnet_infer_bf16_b16_512x1024; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c50_$w.json 2> $P/sweep/c50_$w.err
  python tools/show_bench.py $P/sweep/c50_$w.json 2>/dev/null | head -1; tail -2 $P/sweep/c50_$w.err
done
ESN_DUAL=0 timeout 600 python bench.py --workload dabnet_infer_bf16_b16_1024x2048 --no-cpu-baseline --no-gpu-eager --no-legs > $P/sweep/c50_dabnet_nodual.json 2> $P/sweep/c50_dabnet_nodual.err
python tools/show_bench.py $P/sweep/c50_dabnet_nodual.json 2>/dev/null | head -1
