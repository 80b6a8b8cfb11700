#!/bin/bash
# Data-parallel training at N GPUs (N = $1): SM reserve for the backward pass x NCCL CTA budget; plus the dp checks and
# the concurrent H2D ceiling.  One line per configuration in gpurun_out/r2_dp_N.txt.
set -u
N=${1:-2}
OUT=gpurun_out/r2_dp_$N.txt
: > $OUT
run() {
  local label="$*"
  local line
  line=$(env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 \
         bench.py --gpus $N --mode train --steps 12 --warmup 4 2>gpurun_out/r2_dp_err.log | grep '^{' | tail -1)
  python - "$label" "$line" >> $OUT <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:44s} {d['ms_per_step']:8.3f} ms/step {d['value']:10.0f} img/s  e2e {d['e2e']['value']:10.0f}  dp {json.dumps(d.get('dp'))}")
except Exception as e:
    print(f"{label:44s} FAILED ({e}): {line[:300]}")
PY
  tail -1 $OUT
}
if [ "${2:-full}" = "short" ]; then
  run CTK_DP_SM_RESERVE=0
  run CTK_DP_SM_RESERVE=16
  run CTK_DP_SM_RESERVE=32
  run CTK_DP_SM_RESERVE=16 NCCL_MAX_CTAS=16
else
  run CTK_DP_SM_RESERVE=0
  run CTK_DP_SM_RESERVE=16
  run CTK_DP_SM_RESERVE=24
  run CTK_DP_SM_RESERVE=32
  run CTK_DP_SM_RESERVE=16 NCCL_MAX_CTAS=16
  run CTK_DP_SM_RESERVE=32 NCCL_MAX_CTAS=32
  run CTK_DP_SM_RESERVE=8 NCCL_MAX_CTAS=8
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 tools/probe_h2d_ranks.py 2>/dev/null | grep '^{' | tee gpurun_out/r2_h2d_ranks_$N.json
