#!/bin/bash
# Source-level (SASS + stall samples) ncu captures of individual training kernels, exported to CSV on the box so that the
# copy-back stays small.  Kernels are picked by base name + launch index inside the 4th training step (3 warm-up steps).
# Usage (under gpurun): bash tools/profile_kernel_sources.sh
set -u
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_src.log 2>&1 || exit 1
for spec in "conv3x3_tc_kernel:44:dgrad_n64" "conv3x3_tc_kernel:36:fwd_stats_n128" "conv3x3_tc_kernel:43:dgrad_n128" "wgrad_tc_kernel:20:wgrad_nci64" "wgrad_tc_kernel:18:wgrad_nci128" "first_wgrad_codes_kernel:6:first_wgrad"; do
  IFS=: read -r name skip tag <<< "$spec"
  ncu --set full --clock-control none --import-source on -k "regex:$name" -s $skip -c 1 -f -o /tmp/src_$tag $TRN > gpurun_out/ncu_src_$tag.log 2>&1
  ncu -i /tmp/src_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$tag.csv 2>/dev/null
  ncu -i /tmp/src_$tag.ncu-rep --page raw --csv > gpurun_out/src_${tag}_raw.csv 2>/dev/null
  head -c 160 gpurun_out/src_$tag.csv | head -1
done
du -sh gpurun_out
