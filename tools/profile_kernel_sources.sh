#!/bin/bash
# Source-level (SASS + stall samples) ncu captures of individual training kernels, exported to CSV on the box so that the
# copy-back stays small.  Usage (under gpurun): bash tools/profile_kernel_sources.sh
set -u
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_src.log 2>&1 || exit 1
i=0
for spec in "conv3x3_tc_kernel<2, 64, 2>:2" "conv3x3_tc_kernel<2, 128, 3>:2" "wgrad_tc_kernel<3>:2" "first_wgrad_codes_kernel:2" "conv3x3_tc_kernel<2, 128, 2>:2" "bn_bwd_apply_kernel:4"; do
  name="${spec%%:*}"; skip="${spec##*:}"
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k "regex:$(echo "$name" | sed 's/[<>]/./g; s/, /.*/g')" -s $skip -c 1 -f -o /tmp/src_$i $TRN > gpurun_out/ncu_src_$i.log 2>&1
  ncu -i /tmp/src_$i.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$i.csv 2>/dev/null
  ncu -i /tmp/src_$i.ncu-rep --page raw --csv > gpurun_out/src_${i}_raw.csv 2>/dev/null
  head -c 200 gpurun_out/src_$i.csv | head -1
done
du -sh gpurun_out
