#!/bin/bash
# Which half of the side-stream schedule pays: weight packing beside the first block, deferred weight gradients, stream
# priority, 128-thread BatchNorm CTAs (more of them fit beside a wgrad CTA).  Double-branch training step, 20 steps each.
set -u
mkdir -p gpurun_out
run() {  # label, env, extra args
  env $2 timeout 200 python bench.py --mode train --model ${4:-double} --steps 20 --warmup 5 --no-cpu-baseline $3 2>gpurun_out/r2r_err.log > gpurun_out/r2r_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2r_{l}.json").read().strip().splitlines()[-1])
    r=d['roofline']; pc=r['per_call_ms_per_step']
    print(f"{l:28s} ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} wgrad {pc.get('ctk_conv3x3_wgrad_tc')} apply {pc.get('ctk_bn_bwd_apply')} reduce {pc.get('ctk_bn_bwd_reduce_guarded')} conv {pc.get('ctk_conv3x3_tc_raw')} clk {d["clocks"]["sm_mhz"]} W {d["clocks"].get("power_w")}/{d["clocks"].get("power_limit_w")}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2r_err.log').read()[-1500:])
P
}
run plain            "CTK_OVERLAP_WGRAD=0" ""
run both             "CTK_OVERLAP_WGRAD=1" ""
run pack_only        "CTK_OVERLAP_WGRAD=pack" ""
run wgrad_only       "CTK_OVERLAP_WGRAD=wgrad" ""
run both_noprio      "CTK_OVERLAP_WGRAD=1 CTK_WGRAD_PRIORITY=0" ""
run both_bn128       "CTK_OVERLAP_WGRAD=1 CTK_BN_BLOCK=128" ""
run plain_bn128      "CTK_OVERLAP_WGRAD=0 CTK_BN_BLOCK=128" ""
run plain2           "CTK_OVERLAP_WGRAD=0" ""
run single_pack_only "CTK_OVERLAP_WGRAD=pack" "" single
run single_bn128     "CTK_OVERLAP_WGRAD=1 CTK_BN_BLOCK=128" "" single
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm --format=csv
