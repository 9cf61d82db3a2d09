#!/usr/bin/env bash
# Round-2 call 34: ncu --set full of the new training kernels (one eager DABNet iteration, profiler range = one step) and the
# launch list of the same command
set -u
mkdir -p gpurun_out
timeout 120 python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/one_step.log 2>&1; tail -1 gpurun_out/one_step.log
K="--profile-from-start off --set full --clock-control none --import-source on"
timeout 400 ncu $K -k regex:bilinear_bwd_rows_kernel -c 1 -o gpurun_out/r02_bilinear_bwd_rows -f python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu1.log 2>&1
timeout 400 ncu $K -k regex:bn_act_bwd_fused_kernel -s 30 -c 1 -o gpurun_out/r02_bn_bwd_fused -f python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu2.log 2>&1
timeout 400 ncu $K -k regex:bn_act_train_fwd_kernel -s 30 -c 1 -o gpurun_out/r02_bn_fwd_fused -f python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu3.log 2>&1
timeout 400 ncu $K -k regex:wgrad_rows_kernel -s 3 -c 1 -o gpurun_out/r02_wgrad_rows_c32 -f python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r02b_launches_dabnet_train.csv python tools/train_one_step.py DABNet 8 512 1024 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu1.log gpurun_out/ncu5.log
