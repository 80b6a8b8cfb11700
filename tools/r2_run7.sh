#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2g_build.log 2>&1
timeout 1700 python -m pytest tests/test_gpu_f32_train.py tests/test_gpu_parity_r2.py -m gpu -q -s > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
grep -v "^  \|^   window" gpurun_out/r2g_pytest.log | tail -60
