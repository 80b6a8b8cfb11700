#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2c_build.log 2>&1
python tools/probe_accum_bias.py > gpurun_out/r2c_accum_bias.txt 2>&1; cat gpurun_out/r2c_accum_bias.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print("train", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"])
for k, v in sorted(d["roofline"]["per_call_ms_per_step"].items(), key=lambda kv: -kv[1]): print(f"   {k:36s} {v:.4f}")
print("infer", d["infer"]["ms_per_step"], d["infer"]["value"], "e2e", d["infer"]["e2e"]["value"])
for k, v in sorted(d["infer"]["roofline"]["per_call_ms_per_step"].items(), key=lambda kv: -kv[1]): print(f"   {k:36s} {v:.4f}")
print("single", d["train_single"]["ms_per_step"], d["train_single"]["value"])
for k, v in sorted(d["train_single"]["roofline"]["per_call_ms_per_step"].items(), key=lambda kv: -kv[1]): print(f"   {k:36s} {v:.4f}")
print("cpu", d["cpu_baseline"])
PY
timeout 300 python bench.py --mode sweep --tiles 20480 --precision bf16 > gpurun_out/r2c_sweep_bf16.json 2> gpurun_out/r2c_sweep.err; tail -c 900 gpurun_out/r2c_sweep_bf16.json; tail -c 300 gpurun_out/r2c_sweep.err
timeout 300 python bench.py --mode sweep --tiles 20480 --precision fp32 > gpurun_out/r2c_sweep_fp32.json 2>> gpurun_out/r2c_sweep.err; tail -c 900 gpurun_out/r2c_sweep_fp32.json
