#!/bin/bash
# Same-box A/B: libctk at HEAD (base) against the working tree (new), inference and training.
set -u
mkdir -p gpurun_out
run() {  # label, env, mode
  env $2 timeout 200 python bench.py --mode $3 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2v_err.log > gpurun_out/r2v_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2v_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    keys=['ctk_conv_first_eval','ctk_gemm_bf16_splitk','ctk_head_eval','ctk_pearson_f32','ctk_conv3x3_tc_eval','ctk_conv_first_pool_codes','ctk_first_patch_gram','ctk_bn_bwd_apply','ctk_first_wgrad_codes']
    print(f"{l:14s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f}", {k.replace('ctk_',''):pc[k] for k in keys if k in pc}, 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2v_err.log').read()[-1500:])
P
}
BASE="CTK_LIB=$PWD/torch-unet_b200/ctk/libctk_base.so CTK_FC1_SPLITS=pow2"
run infer_base  "$BASE" infer
run infer_new   "A=1" infer
run infer_base2 "$BASE" infer
run infer_new2  "A=1" infer
run train_base  "$BASE CTK_OVERLAP_WGRAD=0" train
run train_new   "A=1" train
run train_base2 "$BASE CTK_OVERLAP_WGRAD=0" train
run train_new2  "A=1" train
