#!/usr/bin/env bash
# Round-2 call 48: where do DABNet's small-channel 3x3 convs (2.9 of 10.2 ms, 0.2-0.3 of their HBM roof) lose their time?
# timing with / without the row ring, then one ncu --set full capture per shape
set -u
P=gpurun_out
for cfg in "32 32 3 3 1 16 512 1024" "64 32 3 3 1 16 256 512" "128 64 3 3 1 16 128 256" "32 64 1 1 1 16 256 512" "64 128 1 1 1 16 128 256"; do
  timeout 120 python tools/prof_conv.py $cfg 0 20
  ESN_UMMA_NOHV=1 timeout 120 python tools/prof_conv.py $cfg 0 20 | sed 's/^/   (one box per tap) /'
done
i=0
for cfg in "32 32 3 3 1 16 512 1024" "64 32 3 3 1 16 256 512" "128 64 3 3 1 16 128 256"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma --launch-skip 2 -c 1 -f -o $P/r02_conv3x3_$i python tools/prof_conv.py $cfg 0 3 > $P/r02_conv3x3_$i.log 2>&1
  tail -1 $P/r02_conv3x3_$i.log
done
