#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2n_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_f32_train.py -m gpu -q -s -k "first_block_gram_path or bn_finalize or loss_curve_at_batch_64" > gpurun_out/r2n_pytest.log 2>&1
echo "pytest exit $?"; grep -v "^  \|^   window" gpurun_out/r2n_pytest.log | grep -i "passed\|failed\|error\|fp32:\|bf16:\|golden has\|step:\|first 5" | tail -12
timeout 200 python bench.py --mode train --model double --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2n_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('double', round(d['ms_per_step'],3), round(d['value']), {k:v for k,v in d['roofline']['per_call_ms_per_step'].items() if 'bn_' in k})"
