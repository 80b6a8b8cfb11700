#!/bin/bash
# ncu evidence for profiles/: launch lists (device time per launch), one --set full capture of every kernel of one
# inference step, and a memory/compute/occupancy capture of every kernel of one training step.  Run under gpurun from
# the repo root; outputs land in gpurun_out/ (kept under the 64 MiB copy-back limit: big reports are exported to CSV on
# the box and deleted).
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
ncu -i gpurun_out/infer_${TAG}_full.ncu-rep --page raw --csv > gpurun_out/infer_${TAG}_full_raw.csv 2>/dev/null
$TRN > /dev/null 2>&1 &&
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis \
    --clock-control none -k regex:'conv3x3_tc|wgrad|bn_|first_|patch_gram|conv_first|adam|gemm|pack_fc1|feat_transpose' -s 280 -c 95 -f -o /tmp/train_${TAG} $TRN > gpurun_out/ncu_ft_$TAG.log 2>&1
ncu -i /tmp/train_${TAG}.ncu-rep --page raw --csv > gpurun_out/train_${TAG}_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | tail -14
