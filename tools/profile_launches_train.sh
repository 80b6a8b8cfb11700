#!/bin/bash
# Launch list (device time per launch, cold-cache and serialised) of one training step; cheap refresh of
# profiles/<tag>_launches_train.csv after a kernel change.  Run under gpurun from the repo root.
set -u
TAG=${1:-r1e}
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_train_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}_train.csv $TRN > gpurun_out/ncu_lt_$TAG.log 2>&1
ls -la gpurun_out/launches_${TAG}_train.csv
