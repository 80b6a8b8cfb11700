#!/bin/bash
# Lazy weight packing (after the first block is enqueued): plain single stream vs packs on the side stream, same box.
set -u
mkdir -p gpurun_out
run() {  # label, env, model
  env $2 timeout 200 python bench.py --mode train --model ${3:-double} --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2y_err.log > gpurun_out/r2y_$1.json
  python - "$1" <<'P'
import json,sys
l=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2y_{l}.json").read().strip().splitlines()[-1])
    pc=d['roofline']['per_call_ms_per_step']
    print(f"{l:14s} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.4f} sum {sum(pc.values()):.3f} gram {pc['ctk_first_patch_gram']} packfc1 {pc['ctk_pack_fc1_weight_bf16']} clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(l, 'FAILED', e); print(open('gpurun_out/r2y_err.log').read()[-1500:])
P
}
run plain   "CTK_OVERLAP_WGRAD=0"
run pack    "CTK_OVERLAP_WGRAD=pack"
run plain2  "CTK_OVERLAP_WGRAD=0"
run pack2   "CTK_OVERLAP_WGRAD=pack"
run s_plain "CTK_OVERLAP_WGRAD=0" single
run s_pack  "CTK_OVERLAP_WGRAD=pack" single
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "stream_overlap or bit_reproducible or accumulation or eval_after" 2>&1 | tail -3
