#!/bin/bash
# wgrad_tc_kernel: share of CTAs given to kx group 1 (the load-bound B CTAs); default = 1/3 in (A, A, B) triplets.
set -u
mkdir -p gpurun_out
run() {  # label, env
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2ab_err.log > gpurun_out/r2ab_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2ab_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    print(f"{l:10s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} wgrad {pc['ctk_conv3x3_wgrad_tc']} wgradTF {d['roofline']['wgrad_tc_kernel']['achieved']:.0f} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2ab_err.log').read()[-1500:])
P
}
run default "A=1"
run b360    "CTK_WGRAD_B_PERMILLE=360"
run b400    "CTK_WGRAD_B_PERMILLE=400"
run b300    "CTK_WGRAD_B_PERMILLE=300"
run default2 "A=1"
run b360_2  "CTK_WGRAD_B_PERMILLE=360"
CTK_WGRAD_B_PERMILLE=360 timeout 300 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "wgrad" 2>&1 | tail -2
