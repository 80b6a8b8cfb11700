#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2p_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -k "first_block_gram_path" > gpurun_out/r2p_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2p_pytest.log
for m in double single; do
timeout 200 python bench.py --mode train --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2p_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', round(d['ms_per_step'],3), round(d['value']), 'first_wgrad', d['roofline']['per_call_ms_per_step'].get('ctk_first_wgrad_codes'))"
done
