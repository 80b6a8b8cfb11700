#!/bin/bash
# Round-2 ncu evidence for profiles/ (run under gpurun from the repo root, one GPU): launch lists of one training step and
# one inference step, a --set full capture of the dominant tensor-core kernels of the training step (conv raw+stats / dgrad,
# wgrad) and of the inference conv, and a memory/compute/occupancy capture of every kernel of one training step.
set -u
TAG=${1:-r2}
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
INF="python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_train_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_train.csv $TRN > gpurun_out/ncu_lt_$TAG.log 2>&1
$INF > gpurun_out/plain_infer_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_infer.csv $INF > gpurun_out/ncu_li_$TAG.log 2>&1
# training step: launches 4 x ~135 warm-up/timed steps precede the two instrumented ones; capture one full step's tensor kernels
$TRN > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc_kernel|wgrad_tc_kernel' -s 72 -c 18 -f -o gpurun_out/${TAG}_train_full $TRN > gpurun_out/ncu_ftf_$TAG.log 2>&1
ncu -i gpurun_out/${TAG}_train_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_train_full_raw.csv 2>/dev/null
$INF > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc|conv_first|gemm_splitk|head_eval|pearson' -s 36 -c 12 -f -o gpurun_out/${TAG}_infer_full $INF > gpurun_out/ncu_fi_$TAG.log 2>&1
ncu -i gpurun_out/${TAG}_infer_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_infer_full_raw.csv 2>/dev/null
$TRN > /dev/null 2>&1 &&
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis \
    --clock-control none -k regex:'conv3x3_tc|wgrad|bn_|first_|patch_gram|conv_first|adam|gemm|pack_fc1|feat_transpose|reduce_rows|gram_reduce' -s 420 -c 140 -f -o /tmp/train_${TAG} $TRN > gpurun_out/ncu_ft_$TAG.log 2>&1
ncu -i /tmp/train_${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_train_sections_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_infer_full.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | grep $TAG
