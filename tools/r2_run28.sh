#!/bin/bash
# Small-kernel pass (colstat / bn1d_bwd_reduce with 32 row groups, sgemm K-tile prefetch, first_moments from shared memory)
# against HEAD, same box; training tests first.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_training.py -m gpu -q -x 2>&1 | tail -3
run() {  # label, env
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2aa_err.log > gpurun_out/r2aa_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2aa_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    keys=['ctk_colstat','ctk_sgemm_strided','ctk_bn1d_bwd_reduce','ctk_first_moments']
    print(f"{l:10s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} sum {sum(pc.values()):.3f}", {k[4:]:pc[k] for k in keys}, 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2aa_err.log').read()[-1500:])
P
}
BASE="CTK_LIB=$PWD/torch-unet_b200/ctk/libctk_base.so"
run base  "$BASE"
run new   "A=1"
run base2 "$BASE"
run new2  "A=1"
