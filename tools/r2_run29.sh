#!/bin/bash
# Round-end rehearsal on one GPU: the whole GPU suite, smoke(), the default bench line and the reference arm.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python bench.py 2>gpurun_out/r2f_bench_err.log > gpurun_out/r2f_bench.json
echo "bench exit $?"
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print('train', round(d['ms_per_step'],3), 'ms', round(d['value']), 'img/s e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['achieved']), d['roofline']['frac'], 'wgrad', d['roofline']['wgrad_tc_kernel'], 'clk', d['clocks'])
    i=d['infer']; print('infer', round(i['ms_per_step'],4), round(i['value']), 'e2e', round(i['e2e']['value']), 'frac', i['roofline']['frac'])
    s=d['train_single']; print('single', round(s['ms_per_step'],3), round(s['value']), 'e2e', round(s['e2e']['value']))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench FAILED', e); print(open('gpurun_out/r2f_bench_err.log').read()[-2000:])
P
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r2f_ref_err.log | tail -1 | cut -c1-600
