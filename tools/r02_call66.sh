#!/usr/bin/env bash
# Round-2 call 66: final tree -- full GPU test-suite, smoke(), the default bench line as the driver runs it, ncu launch list of one
# eager training iteration (the iteration GraphedTrainStep captures: fused close + one-launch Adam)
set -u
P=gpurun_out
mkdir -p $P/sweep
timeout 2400 python -m pytest tests -q -m gpu > $P/r02_tests_call66.log 2>&1
tail -3 $P/r02_tests_call66.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py > $P/r02_bench_default_n1.json 2> $P/r02_bench_default_n1.err ) 2> $P/r02_bench_default_n1.time
tail -3 $P/r02_bench_default_n1.time
python tools/show_bench.py $P/r02_bench_default_n1.json 2>/dev/null | head -1
timeout 120 python tools/train_one_step.py DABNet 8 512 1024 > $P/one_step.log 2>&1; tail -1 $P/one_step.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $P/r02_launches_dabnet_train.csv python tools/train_one_step.py DABNet 8 512 1024 > $P/ncu_launches.log 2>&1
tail -1 $P/ncu_launches.log; wc -l $P/r02_launches_dabnet_train.csv
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:adam_table_kernel -c 1 -f -o $P/r02_adam_table python tools/train_one_step.py DABNet 8 512 1024 > $P/ncu_adam.log 2>&1
tail -1 $P/ncu_adam.log
