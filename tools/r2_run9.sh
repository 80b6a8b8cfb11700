#!/bin/bash
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2h_build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
grep -v "^  \|^   window" gpurun_out/r2h_pytest.log | grep -i "passed\|failed\|error\|pytest exit\|fp32:\|bf16:\|golden has\|32 fixture" | tail -30
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print("train", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
for k, v in sorted(d["roofline"]["per_call_ms_per_step"].items(), key=lambda kv: -kv[1])[:14]: print(f"   {k:36s} {v:.4f}")
print("infer", d["infer"]["ms_per_step"], d["infer"]["value"], "e2e", d["infer"]["e2e"]["value"])
print("single", d["train_single"]["ms_per_step"], d["train_single"]["value"])
print("cpu", d["cpu_baseline"], d["infer"].get("cpu_baseline"))
PY
python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1 && timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r2h_memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/r2h_memcheck.log
