#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2f_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
bash tools/profile_round2.sh r2
