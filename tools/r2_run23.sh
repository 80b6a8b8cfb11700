#!/bin/bash
# Occupancy variants of the first block's Gram kernel (roles per CTA) and of bn_bwd_apply (pixels per iteration / CTAs per SM).
set -u
mkdir -p gpurun_out
run() {  # label, env
  env $2 timeout 200 python bench.py --mode train --model double --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2t_err.log > gpurun_out/r2t_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2t_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    print(f"{l:22s} ms/step {d['ms_per_step']:.3f} gram {pc.get('ctk_first_patch_gram')} apply {pc.get('ctk_bn_bwd_apply')} fwdpool {pc.get('ctk_bn_act_pool_fwd')} reduce {pc.get('ctk_bn_bwd_reduce_guarded')} packfc1 {pc.get('ctk_pack_fc1_weight_bf16')} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2t_err.log').read()[-1500:])
P
}
run gram0_apply0 "CTK_GRAM_VARIANT=0 CTK_BN_APPLY_VARIANT=0"
run gram1_apply1 "CTK_GRAM_VARIANT=1 CTK_BN_APPLY_VARIANT=1"
run gram2_apply2 "CTK_GRAM_VARIANT=2 CTK_BN_APPLY_VARIANT=2"
run gram1_apply3 "CTK_GRAM_VARIANT=1 CTK_BN_APPLY_VARIANT=3"
run gram2_apply0 "CTK_GRAM_VARIANT=2 CTK_BN_APPLY_VARIANT=0"
for v in 1 2; do
CTK_GRAM_VARIANT=$v CTK_BN_APPLY_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_training.py -m gpu -q -x -k "gram or bn_finalize_act_pool or bit_reproducible or forward_loss_and_gradients" > gpurun_out/r2t_pytest_$v.log 2>&1
echo "variant $v pytest exit $?"; tail -3 gpurun_out/r2t_pytest_$v.log
done
