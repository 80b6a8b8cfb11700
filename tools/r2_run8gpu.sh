#!/bin/bash
# final 8-GPU record of the round: the bench headline (training + attached inference + dp checks) and configs[4], the
# 1M-tile bulk sweep in both precisions
set -u
mkdir -p gpurun_out
python torch-unet_b200/build.py > gpurun_out/r2o_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r2o_bench8.err | grep '^{' | tail -1 > gpurun_out/r2o_bench8.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench8.json").read())
print("train8", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "dp", json.dumps(d.get("dp")), "allreduce", d.get("allreduce"))
print("infer8", d["infer"]["ms_per_step"], d["infer"]["value"], "e2e", d["infer"]["e2e"]["value"])
PY
timeout 300 $TR bench.py --gpus 8 --mode sweep --tiles 1000000 --precision bf16 2>gpurun_out/r2o_sweep.err | grep '^{' | tail -1 > gpurun_out/r2o_sweep8_bf16.json
timeout 300 $TR bench.py --gpus 8 --mode sweep --tiles 1000000 --precision fp32 2>>gpurun_out/r2o_sweep.err | grep '^{' | tail -1 > gpurun_out/r2o_sweep8_fp32.json
python - <<'PY'
import json
for p in ("bf16", "fp32"):
    d = json.loads(open(f"gpurun_out/r2o_sweep8_{p}.json").read())
    print("sweep8", p, round(d["value"]), "img/s", round(d["total_ms"], 1), "ms", d["parity_spot_check"])
PY
tail -c 300 gpurun_out/r2o_bench8.err; tail -c 300 gpurun_out/r2o_sweep.err
