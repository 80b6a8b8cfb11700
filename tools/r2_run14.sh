#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2l_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -x -q -k "first" > gpurun_out/r2l_pytest1.log 2>&1
echo "pytest first-block exit $?"; tail -3 gpurun_out/r2l_pytest1.log
for m in double single; do
    CTK_FIRST_WGRAD=tc timeout 200 python bench.py --mode train --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2l_err_${m}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m tc', round(d['ms_per_step'],3), round(d['value']), 'first_wgrad_codes', d['roofline']['per_call_ms_per_step'].get('ctk_first_wgrad_codes'))"
done
CMD="python bench.py --mode train --model double --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2l_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:first_wgrad_tc -s 4 -c 1 -f -o gpurun_out/r2l_fwtc $CMD > gpurun_out/r2l_ncu.log 2>&1
ncu -i gpurun_out/r2l_fwtc.ncu-rep --page source --csv > gpurun_out/r2l_fwtc_source.csv 2>/dev/null
ncu -i gpurun_out/r2l_fwtc.ncu-rep --page raw --csv > gpurun_out/r2l_fwtc_raw.csv 2>/dev/null
rm -f gpurun_out/r2l_fwtc.ncu-rep
