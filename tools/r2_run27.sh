#!/bin/bash
# wgrad_tc_kernel with four pipeline stages (working tree) against three (libctk_base.so = HEAD), same box.
set -u
mkdir -p gpurun_out
run() {  # label, env
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2z_err.log > gpurun_out/r2z_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2z_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    print(f"{l:10s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} sum {sum(pc.values()):.3f} wgrad {pc['ctk_conv3x3_wgrad_tc']} wgradTF {d['roofline']['wgrad_tc_kernel']['achieved']:.0f} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2z_err.log').read()[-1500:])
P
}
BASE="CTK_LIB=$PWD/torch-unet_b200/ctk/libctk_base.so"
run base  "$BASE"
run new   "A=1"
run base2 "$BASE"
run new2  "A=1"
run s_base "$BASE" single
run s_new  "A=1" single
timeout 300 python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "wgrad" 2>&1 | tail -3
