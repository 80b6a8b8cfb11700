#!/bin/bash
# Do the BatchNorm passes share SMs with wgrad_tc_kernel once they ask for the same shared-memory carve-out?
set -u
mkdir -p gpurun_out
run() {  # label, env, model
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2s_err.log > gpurun_out/r2s_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2s_{l}.json").read().strip().splitlines()[-1])
    r=d['roofline']; pc=r['per_call_ms_per_step']
    print(f"{l:28s} ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} wgrad {pc.get('ctk_conv3x3_wgrad_tc')} apply {pc.get('ctk_bn_bwd_apply')} reduce {pc.get('ctk_bn_bwd_reduce_guarded')} conv {pc.get('ctk_conv3x3_tc_raw')} fwdpool {pc.get('ctk_bn_act_pool_fwd')} clk {d['clocks']['sm_mhz']} W {d['clocks'].get('power_w')}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2s_err.log').read()[-1500:])
P
}
run plain_default_carveout  "CTK_OVERLAP_WGRAD=0 CTK_BN_CARVEOUT=-1"
run plain_max_carveout      "CTK_OVERLAP_WGRAD=0"
run both_default_carveout   "CTK_OVERLAP_WGRAD=1 CTK_BN_CARVEOUT=-1"
run both_max_carveout       "CTK_OVERLAP_WGRAD=1"
run both_max_noprio         "CTK_OVERLAP_WGRAD=1 CTK_WGRAD_PRIORITY=0"
run both_max_bn128          "CTK_OVERLAP_WGRAD=1 CTK_BN_BLOCK=128"
run single_plain            "CTK_OVERLAP_WGRAD=0" single
run single_both_max         "CTK_OVERLAP_WGRAD=1" single
timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_train_kernels.py -m gpu -q -x -k "stream_overlap or bit_reproducible or bn_finalize_act_pool" > gpurun_out/r2s_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2s_pytest.log
