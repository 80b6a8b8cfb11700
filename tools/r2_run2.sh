#!/bin/bash
# round 2, GPU call 2: the new parity tests (with their measured distances printed), the restructured bench at N = 1
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2b_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s > gpurun_out/r2b_parity.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_parity.log
grep -v "^  " gpurun_out/r2b_parity.log | tail -40
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
print("train", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"])
for k, v in sorted(d["roofline"]["per_call_ms_per_step"].items(), key=lambda kv: -kv[1]): print(f"   {k:36s} {v:.4f}")
print("infer", d["infer"]["ms_per_step"], d["infer"]["value"], "e2e", d["infer"]["e2e"]["value"])
print("single", d["train_single"]["ms_per_step"], d["train_single"]["value"])
print("cpu", d["cpu_baseline"])
PY
timeout 300 python bench.py --mode sweep --tiles 20480 --precision bf16 > gpurun_out/r2b_sweep_bf16.json 2> gpurun_out/r2b_sweep.err; tail -c 1200 gpurun_out/r2b_sweep_bf16.json; tail -c 300 gpurun_out/r2b_sweep.err
timeout 300 python bench.py --mode sweep --tiles 20480 --precision fp32 > gpurun_out/r2b_sweep_fp32.json 2>> gpurun_out/r2b_sweep.err; tail -c 1200 gpurun_out/r2b_sweep_fp32.json
