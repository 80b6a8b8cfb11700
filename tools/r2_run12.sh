#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2k_build.log 2>&1
CMD="python bench.py --mode train --model double --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:first_wgrad_tc -s 4 -c 1 -f -o gpurun_out/r2k_fwtc $CMD > gpurun_out/r2k_ncu.log 2>&1
ncu -i gpurun_out/r2k_fwtc.ncu-rep --page raw --csv > gpurun_out/r2k_fwtc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2k_fwtc.ncu-rep --page source --csv > gpurun_out/r2k_fwtc_source.csv 2>/dev/null
ls -la gpurun_out | grep r2k
