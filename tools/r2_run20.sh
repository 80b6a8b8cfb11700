#!/bin/bash
# Deferred weight-gradient stream (TrainEngine.overlap_wgrad): schedule test + A/B of the training step, then the whole suite.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "stream_overlap or bit_reproducible or accumulation" > gpurun_out/r2q_pytest_sched.log 2>&1
echo "sched pytest exit $?"; tail -5 gpurun_out/r2q_pytest_sched.log
for m in double single; do
  for v in "" "--no-overlap-wgrad"; do
    timeout 200 python bench.py --mode train --model $m --steps 20 --warmup 5 --no-cpu-baseline $v 2>gpurun_out/r2q_err.log > gpurun_out/r2q_bench_${m}_${v:-overlap}.json
    python - "$m" "${v:-overlap}" <<'P'
import json,sys
m,v=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2q_bench_{m}_{v}.json").read().strip().splitlines()[-1])
    r=d['roofline']
    print(m, v, 'ms/step', round(d['ms_per_step'],3), 'img/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'conv TF', round(r['achieved']), 'dgrad', r['dgrad_launches'], 'wgrad', r['wgrad_tc_kernel'], 'clk', d['clocks'])
except Exception as e:
    print(m, v, 'FAILED', e); print(open('gpurun_out/r2q_err.log').read()[-2000:])
P
  done
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest_all.log 2>&1
echo "full pytest exit $?"; tail -8 gpurun_out/r2q_pytest_all.log
