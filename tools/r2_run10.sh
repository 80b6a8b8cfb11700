#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2i_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -m gpu -x -q -k "first" > gpurun_out/r2i_pytest1.log 2>&1
echo "pytest first-block exit $?"; tail -15 gpurun_out/r2i_pytest1.log
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity_r2.py -m gpu -q -k "bit_reproducible or gram_and_stored or forward_loss_and_gradients or batch_64 or edge_cases" > gpurun_out/r2i_pytest2.log 2>&1
echo "pytest e2e exit $?"; tail -8 gpurun_out/r2i_pytest2.log
for m in double single; do
  for k in tc cuda; do
    CTK_FIRST_WGRAD=$k timeout 200 python bench.py --mode train --model $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2i_err_${m}_$k.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m $k', round(d['ms_per_step'],3), round(d['value']), 'first_wgrad_codes', d['roofline']['per_call_ms_per_step'].get('ctk_first_wgrad_codes'))"
  done
done
