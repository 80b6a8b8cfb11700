#!/usr/bin/env python
"""bench.py -- images/sec of the CrosstalkPy hot path on B200 (BASELINE.json metric), one JSON line.

Workload at every N: BASELINE.json configs[1] -- double-branch (-o double) inference of a batch of 256 synthetic
2-channel 256x256 tiles per GPU plus the per-tile Pearson baseline.  One "step" = Pearson + eval forward over
one 256-tile batch.  Tiles are independent, so N GPUs shard tiles with no collective ("weak" scaling).
The metric is "train & infer images/sec": the double-branch training step (batch 256 per GPU, gradients averaged by the
bucketed NCCL all-reduce at N > 1) is measured in the same run and attached under "train".

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm (inference headline + "train")
    python bench.py --mode train [--model single|double] [--sync-bn]   # the training step as its own line
    python bench.py --precision fp32                               # fp32-class inference path
    python bench.py --impl reference [...]                         # the reference's CPU path (oracle port)

Under torchrun (N > 1) each rank drives one GPU; rank 0 prints the line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "torch-unet_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "infer images/sec (double-branch, 2ch 256x256, batch 256/GPU, + Pearson)"
BATCH = 256
GFLOP_PER_IMG = 14.92          # double-branch forward


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "bf16_tflops_burst": p["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1400.0, "src": "fallback"}


NOMINAL_BF16_TFLOPS = 2250.0     # B200 dense bf16 (B200_PROFILING.md)


def ncu_dram_traffic(kernel_substr, csv_name="r1d_infer_full_raw.csv"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, averaged over the launches of the committed
    `ncu --set full` capture under profiles/ (None if the file is absent)."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None, None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for r in rows[2:]:
        if kernel_substr in r[idx["Kernel Name"]]:
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[key]].replace(",", "")) * scale.get(units[idx[key]], 1.0)
            n += 1
    return (tot / n if n else None), os.path.join("profiles", csv_name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(n_tiles, steps, warmup):
    """The reference's CPU path (oracle port: same ATen ops the reference modules call) on the host cores."""
    import crosstalk_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    x, _ = orc.synthetic_batch(n_tiles, seed=1234)
    sd = orc.init_double_state_dict(0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.pearson_batch(x, f64=False)                       # scipy-style loop, test-cross-talk-model.py:58-64
            orc.double_forward(sd, x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_tiles / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_tiles = 16
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    rate, sec, cores = cpu_reference_rate(n_tiles, steps, warmup)
    sample = f"{n_tiles} of the {BATCH} tiles of one step (pearson loop + double-branch eval forward), fp32, torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "double-branch inference, batch 256 synthetic 2ch 256x256 tiles per GPU + Pearson "
                                   "(BASELINE.json configs[1])",
                       "per_gpu_batch": BATCH, "global_batch": BATCH, "parallelism": "host CPU cores, rank 0 only",
                       "bounded_sample_tiles": n_tiles,
                       "weights": "seed-0 random init"},
            "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_train_rate(kind, n_tiles, steps, warmup):
    """Reference training step (forward + MSE + backward + Adam) on the host cores: oracle port of train_model.py:419-424."""
    import crosstalk_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = orc.synthetic_batch(n_tiles, seed=1234)
    tr = orc.OracleTrainer(kind, orc.INIT[kind](0), lr=5e-4, weight_decay=1e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tr.step(x, y)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_tiles / sec, sec, torch.get_num_threads()


def run_train(args, shared_pg=False):
    """Training throughput: zero_grad -> forward -> MSELoss -> backward -> Adam.step -> loss.item() (train_model.py:419-426),
    per-GPU batch fixed (weak scaling), gradients averaged across ranks by the bucketed NCCL all-reduce.
    Prints its own JSON line (--mode train) or, with shared_pg=True, returns it for the "train" key of the default line."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    kind, batch = args.model, args.batch
    if args.impl == "reference":
        if rank == 0:
            rate, sec, cores = cpu_train_rate(kind, 8, max(1, min(args.steps, 3)), 1)
            print(json.dumps({"impl": "reference", "metric": f"train images/sec ({kind}-branch, 2ch 256x256)", "value": rate,
                              "unit": "images/sec", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 1,
                              "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "f32", "data": "synthetic",
                              "config": {"workload": f"{kind}-branch training step", "bounded_sample_tiles": 8},
                              "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                                               "sample": "8-tile batches, fwd+MSE+bwd+Adam, fp32, torch CPU"},
                              "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
                  flush=True)
        return
    import ctk
    from ctk import _lib
    import crosstalk_oracle as orc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ctk hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not shared_pg:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = (min(args.steps, 10), 3) if shared_pg else (args.steps, max(3, args.warmup))
    torch.manual_seed(0)
    model = (ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64) if kind == "double"
             else ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)).to(dev).train()
    ctk.models.get_train_engine(model).overlap_streams = bool(getattr(args, "overlap_streams", False))
    sync = None
    if world > 1:
        ctk.parallel.broadcast_parameters(model)
        sync = ctk.parallel.attach(model, sync_bn=args.sync_bn)
    opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    crit = torch.nn.MSELoss()
    base_x, base_y = orc.synthetic_batch(32, seed=1234 + rank)
    reps = (batch + 31) // 32
    host = [(base_x.roll(i, 0).repeat(reps, 1, 1, 1)[:batch].contiguous().pin_memory(),
             base_y.roll(i, 0).repeat(reps, 1)[:batch].contiguous().pin_memory()) for i in range(2)]
    devb = [(a.to(dev), b.to(dev)) for a, b in host]

    def step(xb, yb):
        opt.zero_grad()
        out = model(xb)
        loss = crit(out, yb)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(*devb[i % 2]).item()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count
    # inside the timed region only the tensor-core kernels the roofline is about carry CUDA events; the full per-call
    # breakdown comes from two extra instrumented steps afterwards (events around all ~100 calls of a step cost time)
    timeline = _lib.start_timeline(only=("ctk_conv3x3_tc_raw", "ctk_conv3x3_wgrad_tc"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = float("nan")
    for i in range(steps):
        last = step(*devb[i % 2]).item()            # loss.item() every step, like train_model.py:426
    e1.record()
    barrier()
    _lib.stop_timeline()
    launches = _lib.launch_count - launches0
    detail = _lib.start_timeline()
    for i in range(2):
        step(*devb[i % 2]).item()
    barrier()
    _lib.stop_timeline()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    # end to end: the batch starts in pinned host memory every step (the DataLoader's pin_memory=True path, :607-614)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(4, steps)
    pre = ctk.DevicePrefetcher(device=str(dev))
    for xb, yb in pre.iterate(host[i % 2] for i in range(2)):                 # staging buffers allocated, copies warm
        step(xb, yb).item()
    barrier()
    t0 = time.perf_counter()
    for xb, yb in pre.iterate(host[i % 2] for i in range(e2e_steps)):
        step(xb, yb).item()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        per = {}
        for name, a, b, meta in timeline:
            d = per.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
            d["flops"] += (meta or {}).get("flops", 0.0)
        per_all = {}
        for name, a, b, meta in detail:
            d = per_all.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
            d["flops"] += (meta or {}).get("flops", 0.0)
        tc = {k: per[k] for k in ("ctk_conv3x3_tc_raw", "ctk_conv3x3_wgrad_tc") if k in per}
        tc_flops = sum(v["flops"] for v in tc.values())
        tc_ms = sum(v["ms"] for v in tc.values())
        tf = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        gflop_img = {"double": 44.6, "single": 77.6}[kind]
        line = {"metric": f"train images/sec ({kind}-branch, 2ch 256x256, batch {batch}/GPU)", "value": world * batch * steps / (ms / 1e3),
                "unit": "images/sec", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{kind}-branch training step (fwd + MSE + bwd + Adam lr 5e-4 wd 1e-4), batch {batch} per GPU",
                           "per_gpu_batch": batch, "global_batch": batch * world,
                           "parallelism": (f"dp{world}: bucketed NCCL all-reduce (AVG) overlapped with backward"
                                           + (", SyncBN" if args.sync_bn else ", per-rank BatchNorm statistics")) if world > 1 else "single GPU",
                           "l2_policy": "inputs larger than L2 (134 MB per batch), 2 distinct batches rotated"},
                "whole_net_tflops": world * batch * steps / (ms / 1e3) * gflop_img / 1e3,
                "roofline": {"kernel": "conv3x3_tc_kernel (fwd+dgrad) + wgrad_tc_kernel", "bound": "tensor", "achieved": tf,
                             "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"], "traffic": None,
                             "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
                             "frac_of_burst_cublas": tf / pk["bf16_tflops_burst"], "frac_of_nominal_dense": tf / NOMINAL_BF16_TFLOPS,
                             "share_of_step": tc_ms / ms if ms > 0 else None,
                             "per_call_ms_per_step": {k: round(v["ms"] / 2, 4) for k, v in sorted(per_all.items())},
                             "per_call_tflops": {k: round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)
                                                 for k, v in sorted(per_all.items()) if v["ms"] > 0 and v["flops"] > 0},
                             "per_call_source": "two fully instrumented steps after the timed region"},
                "clocks": clocks, "gpu_launches": launches, "last_loss": last,
                "e2e": {"value": world * batch * e2e_steps / (e2e_ms / 1e3), "unit": "images/sec",
                        "h2d_bytes_per_step": batch * (2 * 256 * 256 + 1) * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                        "ms_per_step": e2e_ms / e2e_steps,
                        "api": "for x, y in ctk.DevicePrefetcher().iterate(pinned host batches): model(x); MSELoss; backward; "
                               "ctk.Adam.step; loss.item()  (every step's H2D and loss read-back inside the timed region)"}}
        if sync is not None:
            line["allreduce"] = {"collectives_per_step": sync.collectives / (warmup + steps + e2e_steps),
                                 "bytes_per_step": sync.bytes_reduced / (warmup + steps + e2e_steps)}
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, cores = cpu_train_rate(kind, 8, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                                    "sample": "8-tile batches, fwd+MSE+bwd+Adam, oracle port (same ATen CPU ops as the reference), fp32"}
        if shared_pg:
            return line
        print(json.dumps(line), flush=True)
    del model, opt, devb
    torch.cuda.empty_cache()
    if world > 1 and not shared_pg:
        dist.destroy_process_group()
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ctk", choices=["ctk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="default mode: skip the attached training measurement")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="inference arithmetic: bf16 operands (default, the headline) or the fp32-class split-bf16 path")
    ap.add_argument("--sync-bn", action="store_true", help="train mode, N > 1: BatchNorm statistics over the global batch")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] (default, the headline line); train = configs[2]/[3] training step")
    ap.add_argument("--model", default="double", choices=["double", "single"])
    ap.add_argument("--overlap-streams", action="store_true",
                    help="train mode, EXPERIMENTAL: branches and weight gradients on side streams (TrainEngine.overlap_streams)")
    ap.add_argument("--batch", type=int, default=BATCH, help="per-GPU batch (train mode)")
    args = ap.parse_args()
    if args.mode == "train":
        return run_train(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import ctk
    from ctk import _lib
    import crosstalk_oracle as orc          # cpu_baseline leg + synthetic generator only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ctk hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(3, args.warmup)

    # ---- model (random init of the reference architecture) and synthetic tiles (SURVEY 8d generator)
    torch.manual_seed(0)
    model = ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)
    model.load_state_dict(orc.randomize_bn(model.state_dict(), seed=7))
    model = model.to(dev).eval()
    if args.precision != "bf16":
        ctk.set_precision(model, args.precision)
    base, _ = orc.synthetic_batch(32, seed=1234 + rank)
    n_rot = 3                                   # rotate distinct 134 MB input batches (> 126 MB L2 each)
    host_batches = [base.roll(shifts=i, dims=0).repeat(BATCH // 32, 1, 1, 1).contiguous().pin_memory() for i in range(n_rot)]
    dev_batches = [b.to(dev) for b in host_batches]
    scores = torch.empty(BATCH, 1, device=dev)
    r_out = torch.empty(BATCH, device=dev, dtype=torch.float64)
    engine = ctk.models.get_engine(model)

    def step(i):
        x = dev_batches[i % n_rot]
        ctk.pearson_per_image(x, out=r_out)
        engine.forward(x, out=scores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_count
        timeline = _lib.start_timeline(only=("ctk_conv3x3_tc_eval", "ctk_conv3x3_tc_eval_split"))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        barrier()
        _lib.stop_timeline()
        launches = _lib.launch_count - launches0
        detail = _lib.start_timeline()
        for i in range(3):
            step(i)
        barrier()
        _lib.stop_timeline()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        # ---- end to end: host buffers in, host results out, through the public HostScorer API
        scorer = ctk.HostScorer(model, slice_tiles=64, device=str(dev))
        for _ in scorer.score_stream(host_batches[i % n_rot] for i in range(3)):
            pass
        barrier()
        e2e_steps = max(5, steps)
        t0 = time.perf_counter()
        n_out = 0
        for s_host, r_host in scorer.score_stream(host_batches[i % n_rot] for i in range(e2e_steps)):
            n_out += s_host.numel()             # host results of every step are consumed inside the timed region
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert n_out == e2e_steps * BATCH
        # the one-shot call (no look-ahead across batches), for reference
        barrier()
        t0 = time.perf_counter()
        for i in range(5):
            scorer.score(host_batches[i % n_rot])
        torch.cuda.synchronize()
        oneshot_ms = (time.perf_counter() - t0) * 1e3 / 5
        barrier()

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    value = world * BATCH * steps / (ms / 1e3)
    e2e_value = world * BATCH * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        pk = peaks()
        per = {}
        for name, a, b, meta in timeline:
            d = per.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
            d["flops"] += (meta or {}).get("flops", 0.0)
        per_all = {}
        for name, a, b, meta in detail:
            per_all[name] = per_all.get(name, 0.0) + a.elapsed_time(b) / 3
        conv = per.get("ctk_conv3x3_tc_eval") or per.get("ctk_conv3x3_tc_eval_split") or {"ms": 0.0, "n": 1, "flops": 0.0}
        conv_tf = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        traffic, traffic_src = ncu_dram_traffic("conv3x3_tc_kernel")
        roof = {"kernel": "conv3x3_tc_kernel", "bound": "tensor", "achieved": conv_tf, "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s", "frac": conv_tf / pk["bf16_tflops"], "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch (ncu --set full, mean of the 6 launches of a step)",
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": (537e6 + 268e6 + 268e6 + 134e6 + 134e6 + 67e6) / 3,
                "algorithmic_flop_per_launch": 618.5e9,
                "peak_source": pk["src"] + " (sustained cuBLAS bf16: the kernel is timed inside a long step)",
                "frac_of_burst_cublas": conv_tf / pk["bf16_tflops_burst"], "frac_of_nominal_dense": conv_tf / NOMINAL_BF16_TFLOPS,
                "avg_launch_ms": conv["ms"] / max(1, conv["n"]), "launches": conv["n"],
                "share_of_step": conv["ms"] / ms if ms > 0 else None,
                "per_call_ms_per_step": per_all,
                "per_call_source": "three fully instrumented steps after the timed region"}
        line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class: hi/lo operand pairs, fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": "double-branch inference, batch 256 synthetic 2ch 256x256 tiles per GPU + Pearson "
                                       "(BASELINE.json configs[1])",
                           "per_gpu_batch": BATCH, "global_batch": BATCH * world, "parallelism": f"tiles sharded x{world}, no collective",
                           "l2_policy": "inputs larger than L2 (134 MB per batch), 3 distinct batches rotated",
                           "weights": "seed-0 random init, randomised BN stats"},
                "whole_net_tflops": value * GFLOP_PER_IMG / 1e3,
                "roofline": roof, "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": scorer.h2d_bytes,
                        "d2h_bytes_per_step": scorer.d2h_bytes, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                        "api": "ctk.HostScorer.score_stream(pinned host batches) -> host scores + Pearson r per batch "
                               "(one batch of look-ahead; every step's H2D and D2H inside the timed region)",
                        "oneshot_score_call_ms": oneshot_ms}}
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, cores = cpu_reference_rate(16, 3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                                    "sample": "16 of the 256 tiles of one step (pearson loop + double-branch eval forward), "
                                              "oracle port = the same ATen CPU ops the reference modules call, fp32"}
    # the metric is "train & infer images/sec": the double-branch training step (BASELINE configs[2]/[3] shape, batch 256 per
    # GPU) is measured in the same run and attached under "train" (its own value / e2e / roofline / clocks)
    train_line = None
    if not args.no_train:
        del model, engine, dev_batches, scorer
        torch.cuda.empty_cache()
        targs = argparse.Namespace(**vars(args))
        targs.no_cpu_baseline = True
        train_line = run_train(targs, shared_pg=True)
    if rank == 0:
        if train_line is not None:
            line["train"] = {k: train_line[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config",
                                                        "whole_net_tflops", "roofline", "clocks", "gpu_launches", "e2e")}
            if "allreduce" in train_line:
                line["train"]["allreduce"] = train_line["allreduce"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
