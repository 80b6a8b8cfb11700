#!/bin/bash
# ncu evidence for profiles/: launch lists (device time per launch) and one --set full capture of every kernel of one
# inference step and one training step.  Run under gpurun from the repo root; outputs land in gpurun_out/.
set -u
TAG=${1:-r1b}
INF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$INF > gpurun_out/plain_infer_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}_infer.csv $INF > gpurun_out/ncu_li_$TAG.log 2>&1
$TRN > gpurun_out/plain_train_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}_train.csv $TRN > gpurun_out/ncu_lt_$TAG.log 2>&1
$INF > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|conv_first|gemm_splitk|head_eval|pearson' -s 36 -c 12 -f -o gpurun_out/infer_${TAG}_full $INF > gpurun_out/ncu_fi_$TAG.log 2>&1
$TRN > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|wgrad|bn_|first_|patch_gram|conv_first|adam|gemm|pack_fc1|feat_transpose' -s 280 -c 95 -f -o gpurun_out/train_${TAG}_full $TRN > gpurun_out/ncu_ft_$TAG.log 2>&1
ls -la gpurun_out | tail -12
