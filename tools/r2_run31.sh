#!/bin/bash
# One-launch conv weight packing + wgrad_reduce over four row quarters, against HEAD on the same box; kernel tests first.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_training.py tests/test_gpu_parity_r2.py -m gpu -q -x 2>&1 | tail -3
run() {  # label, env
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2ac_err.log > gpurun_out/r2ac_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2ac_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    packs=sum(v for k,v in pc.items() if 'pack_conv' in k)
    print(f"{l:10s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} sum {sum(pc.values()):.3f} conv-packs {packs:.4f} wgrad {pc['ctk_conv3x3_wgrad_tc']} launches/step {d['gpu_launches']/30:.0f}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2ac_err.log').read()[-1500:])
P
}
BASE="CTK_LIB=$PWD/torch-unet_b200/ctk/libctk_base.so"
run new   "A=1"
run new2  "A=1"
