#!/bin/bash
# Launch lists (training, inference) of the final build of round 2 -> profiles/r2c_launches_*.csv
set -u
mkdir -p gpurun_out
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
INF="python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_train_r2c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/r2c_launches_train.csv $TRN > gpurun_out/ncu_lt_r2c.log 2>&1
$INF > gpurun_out/plain_infer_r2c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches_infer.csv $INF > gpurun_out/ncu_li_r2c.log 2>&1
ls -la gpurun_out/r2c_launches_*.csv
