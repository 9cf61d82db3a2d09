#!/usr/bin/env bash
# Round-2 call 59 (2 GPUs): where the 2-GPU overhead comes from -- N=1 and N=2 on the same box, new close / optimizer against the old ones
set -u
P=gpurun_out
mkdir -p $P/sweep
run() { # name, nproc, flags
  local name=$1 n=$2; shift 2
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --no-legs "$@" > $P/sweep/c59_$name.json 2> $P/sweep/c59_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-legs "$@" \
      > $P/sweep/c59_$name.json 2> $P/sweep/c59_$name.err
  fi
  echo "$name: $(python tools/show_bench.py $P/sweep/c59_$name.json 2>/dev/null | head -1)"
}
run n1_new 1
run n2_new 2
run n2_old_close 2 --no-fused-loss
run n2_old_adam 2 --torch-adam
run n2_old_both 2 --no-fused-loss --torch-adam
run n1_old_both 1 --no-fused-loss --torch-adam
ESN_DP_DEFER=1 run n2_new_defer 2
