#!/bin/bash
# head_eval_kernel with 512 threads / six loads in flight against HEAD (256 / four), inference step, same box.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -2
run() {
  env $2 timeout 200 python bench.py --mode infer --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2ad_err.log > gpurun_out/r2ad_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2ad_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    print(f"{l:8s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} head {pc['ctk_head_eval']} fc1 {pc['ctk_gemm_bf16_splitk']} pearson {pc['ctk_pearson_f32']}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2ad_err.log').read()[-1500:])
P
}
BASE="CTK_LIB=$PWD/torch-unet_b200/ctk/libctk_base.so"
run base "$BASE"
run new  "A=1"
run base2 "$BASE"
run new2  "A=1"
