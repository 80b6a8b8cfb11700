#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2m_build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
grep -v "^  \|^   window" gpurun_out/r2m_pytest.log | grep -i "passed\|failed\|error\|pytest exit\|fp32:\|bf16:\|golden has\|32 fixture\|step:" | tail -30
