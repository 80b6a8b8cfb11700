#!/bin/bash
# Inference tail: 4-images-per-CTA head, one-launch cluster Pearson, ragged split-K FC1 (18 x 8 = 144 CTAs); training defaults
# (gram roles 2, bn_bwd_apply <1,3>).  Kernel tests first, then the inference line and the training line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 --deselect "tests/test_gpu_f32_train.py::test_200_step_loss_curve_at_batch_64[double]" > gpurun_out/r2u_pytest.log 2>&1
echo "pytest exit $?"; tail -14 gpurun_out/r2u_pytest.log
timeout 300 python bench.py --mode infer --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2u_err.log > gpurun_out/r2u_infer.json
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2u_infer.json").read().strip().splitlines()[-1])
    print('infer ms/step', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_call_ms_per_step'], 'conv TF', round(d['roofline']['achieved']))
except Exception as e:
    print('infer FAILED', e); print(open('gpurun_out/r2u_err.log').read()[-2000:])
P
timeout 300 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2u_err2.log > gpurun_out/r2u_train.json
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2u_train.json").read().strip().splitlines()[-1])
    print('train ms/step', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_call_ms_per_step'])
except Exception as e:
    print('train FAILED', e); print(open('gpurun_out/r2u_err2.log').read()[-2000:])
P
