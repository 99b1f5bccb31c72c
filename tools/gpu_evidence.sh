#!/bin/bash
# Evidence session (run under gpurun, ONE GPU): ncu launch list of the bench command, full-set captures of the MLP kernels and
# of K2, summaries for profiles/.  Numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out/evidence
mkdir -p $OUT
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_train_step.csv \
  python bench.py --steps 2 --warmup 3 --profile-steps 2 > $OUT/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 8 -c 4 -o $OUT/mlp python tools/profile_mlp.py > $OUT/ncu_mlp.log 2>&1
python tools/ncu_summary.py $OUT/mlp.ncu-rep > $OUT/ncu_mlp_summary.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:importance -c 2 -o $OUT/k2 python tools/stress_sweep.py --profile > $OUT/ncu_k2.log 2>&1
python tools/ncu_summary.py $OUT/k2.ncu-rep > $OUT/ncu_k2_summary.md 2>&1
ncu -i $OUT/k2.ncu-rep --page raw --csv > $OUT/k2_raw.csv 2>/dev/null
timeout 300 python tools/kernel_timing.py > $OUT/stall_accounting.txt 2>&1
cat $OUT/ncu_mlp_summary.md $OUT/ncu_k2_summary.md
tail -3 $OUT/launches.log
