#!/bin/bash
# A/B of the stream-overlap experiment (branch wip/overlap-streams) in ONE gpurun call (~4 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/ab_overlap_streams.sh'
# 1. the overlap test (same step with and without side streams), 2. training bench: plain / overlap / overlap with
# 128-thread BatchNorm CTAs, double- and single-branch.  Read gpurun_out/ab_overlap.txt afterwards.
set -u
OUT=gpurun_out/ab_overlap.txt
: > $OUT
python torch-unet_b200/build.py > gpurun_out/ab_build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_training.py -m gpu -q -s -k "stream_overlap or double_branch_against" > gpurun_out/ab_pytest.log 2>&1
tail -3 gpurun_out/ab_pytest.log >> $OUT
run() {  # label, env, extra args
  local label=$1; shift
  local envs=$1; shift
  local line
  line=$(env $envs timeout 120 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/ab_err_$label.log | tail -1)
  python - "$label" "$line" >> $OUT <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:32s} {d['ms_per_step']:8.3f} ms/step  {d['value']:10.1f} img/s  e2e {d.get('e2e', {}).get('value', 0):10.1f}")
except Exception as e:
    print(f"{label:32s} FAILED ({e}): {line[:200]}")
PY
}
for model in double single; do
  run ${model}_plain            "CTK_BN_BLOCK=256" --model $model
  run ${model}_overlap          "CTK_BN_BLOCK=256" --model $model --overlap-streams
  run ${model}_overlap_bn128    "CTK_BN_BLOCK=128" --model $model --overlap-streams
  run ${model}_plain_bn128      "CTK_BN_BLOCK=128" --model $model
done
cat $OUT
