#!/bin/bash
# round 2, GPU call 1: full GPU test suite after the deterministic-reduction rewrite, then the stream-overlap A/B.
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2a_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
bash tools/ab_overlap_streams.sh
