#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2d_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_models.py -m gpu -q -s -k "warmed or fp32 or split or precision" > gpurun_out/r2d_parity.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2d_parity.log
grep -v "^  " gpurun_out/r2d_parity.log | tail -25
timeout 300 python bench.py --mode infer --precision fp32 --steps 10 --no-cpu-baseline 2>gpurun_out/r2d_fp32.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp32-class infer', d['ms_per_step'], d['value'], d['roofline']['per_call_ms_per_step'])"
tail -c 300 gpurun_out/r2d_fp32.err
bash tools/r2_dp_experiment.sh 2
