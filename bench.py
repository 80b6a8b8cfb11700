#!/usr/bin/env python
"""bench.py -- images/sec of the CrosstalkPy hot path on B200 (BASELINE.json metric), one JSON line.

Headline (every N): the double-branch TRAINING step of BASELINE.json configs[3] -- zero_grad / forward / MSE / backward /
Adam.step / loss.item() (train_model.py:419-426) on 256 synthetic 2-channel 256x256 tiles per GPU; at N > 1 the gradients
are averaged by the bucketed NCCL all-reduce (global batch 2048 at N = 8), so the driver's scaling record measures the
path that has a collective.  Attached to the same line:
  "infer"        configs[1]: double-branch inference of 256 tiles per GPU + per-tile Pearson r (tiles sharded, no collective)
  "train_single" configs[2]: single-branch training, batch 256, lr 5e-4, ctk.CosineWarmupLR (N = 1 only)
  "dp"           N > 1: parameter checksums agree across ranks after the timed steps; SyncBN shards == whole-batch step

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # headline + attachments
    python bench.py --mode train --model single|double [--sync-bn]        # one training line
    python bench.py --mode infer [--precision bf16|fp32]                  # the inference line alone
    python bench.py --mode sweep --tiles 1000000 --precision bf16|fp32    # configs[4]: bulk sweep, tiles sharded over ranks
    python bench.py --impl reference [...]                                # the reference's CPU path for the same workload

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
_PKG = os.path.join(ROOT, "torch-unet_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402

BATCH = 256
GFLOP_INFER = {"double": 14.92, "single": 25.98}          # SURVEY 8d, per image
GFLOP_TRAIN = {"double": 44.6, "single": 77.6}
NOMINAL_BF16_TFLOPS = 2250.0                              # B200 dense bf16 (B200_PROFILING.md)
CONV_GFLOP_PER_LAUNCH = 618.5                             # 2 N H W Cout 9 Cin, the same for every tensor-core conv of the double model
TRAIN_WORKLOAD = ("{kind}-branch training step (zero_grad, forward, MSE, backward, Adam lr 5e-4 wd 1e-4, loss.item()), "
                  "batch {batch} synthetic 2ch 256x256 tiles per GPU (BASELINE.json configs[{cfg}])")
INFER_WORKLOAD = ("double-branch inference, batch 256 synthetic 2ch 256x256 tiles per GPU + Pearson "
                  "(BASELINE.json configs[1])")


# ------------------------------------------------------------------------------------------------ environment
def env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "bf16_tflops_burst": p["bf16_tflops"], "src": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1400.0, "src": "B200_PROFILING.md fallback"}


def ncu_dram_traffic(kernel_substr, csv_names):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, averaged over the launches of the newest committed
    `ncu --set full` capture under profiles/ that holds it (None if there is none)."""
    import csv
    for csv_name in csv_names:
        path = os.path.join(ROOT, "profiles", csv_name)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        if "dram__bytes_read.sum" not in idx or "Kernel Name" not in idx:
            continue
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot, n = 0.0, 0
        for r in rows[2:]:
            if kernel_substr in r[idx["Kernel Name"]]:
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[idx[key]].replace(",", "")) * scale.get(units[idx[key]], 1.0)
                n += 1
        if n:
            return tot / n, os.path.join("profiles", csv_name)
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")

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
        sm, mx, reasons, watts, limit = [], [], set(), [], []
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
            try:
                watts.append(float(parts[6]))
                limit.append(float(parts[7]))
            except (IndexError, ValueError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(watts) if watts else None, "power_limit_w": max(limit) if limit else None}


# ------------------------------------------------------------------------------------------------ the reference's CPU path
def reference_modules():
    """The UNMODIFIED reference model classes from baseline/_ref (git-ignored copy of the reference's own .py files made by
    __graft_entry__.build() when /root/reference is present; it travels to the GPU box with the snapshot).  None when that
    copy is absent: the CPU arm then runs the oracle port -- the same ATen CPU ops over a state_dict."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not (os.path.exists(os.path.join(ref_dir, "regression_model.py")) and
            os.path.exists(os.path.join(ref_dir, "two_branch_regression.py"))):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from regression_model import AdvancedRegressionModel
    from two_branch_regression import SimplifiedTwoBranchRegressionModel
    return {"single": lambda: AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6),     # train_model.py:537
            "double": lambda: SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)}   # train_model.py:535


def _oracle():
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import crosstalk_oracle as orc          # CPU legs only (cpu_baseline / --impl reference); never on the GPU arm
    return orc


def cpu_train_rate(kind, n_tiles, steps, warmup):
    """Reference training step (train_model.py:419-426) on the host cores: images/sec, seconds per step, threads, kind."""
    from ctk import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    x, y = synthetic.synthetic_batch(n_tiles, seed=1234)
    mods = reference_modules()
    if mods is not None:
        torch.manual_seed(0)
        model = mods[kind]().train()
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)      # train_model.py:637
        crit = torch.nn.MSELoss()                                                    # train_model.py:636

        def step():
            opt.zero_grad()
            loss = crit(model(x), y)
            loss.backward()
            opt.step()
            return loss.item()
        impl_kind = "reference"
    else:
        orc = _oracle()
        tr = orc.OracleTrainer(kind, orc.INIT[kind](0), lr=5e-4, weight_decay=1e-4)
        step = lambda: tr.step(x, y)        # noqa: E731
        impl_kind = "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_tiles / sec, sec, torch.get_num_threads(), impl_kind


def cpu_infer_rate(n_tiles, steps, warmup):
    """Reference inference + per-tile Pearson (test-cross-talk-model.py:44-64) on the host cores."""
    from ctk import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    x, _ = synthetic.synthetic_batch(n_tiles, seed=1234)
    mods = reference_modules()
    if mods is not None:
        from scipy.stats import pearsonr
        import numpy as np
        torch.manual_seed(0)
        model = mods["double"]().eval()

        def step():
            with torch.no_grad():
                model(x)
            images = x.numpy()
            for i in range(images.shape[0]):                                        # test-cross-talk-model.py:58-64
                a, b = images[i, 0].flatten(), images[i, 1].flatten()
                if np.std(a) == 0 or np.std(b) == 0:
                    continue
                pearsonr(a, b)
        impl_kind = "reference"
    else:
        orc = _oracle()
        sd = orc.init_double_state_dict(0)

        def step():
            with torch.no_grad():
                orc.pearson_batch(x, f64=False)
                orc.double_forward(sd, x)
        impl_kind = "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n_tiles / sec, sec, torch.get_num_threads(), impl_kind


def _kind_text(impl_kind):
    return ("the unmodified reference modules from baseline/_ref" if impl_kind == "reference"
            else "oracle port = the same ATen CPU ops the reference modules call")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the workload of the selected mode, rank 0 only."""
    _, rank, _ = env()
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    if args.mode in ("infer", "sweep"):
        n_tiles = 16
        rate, sec, cores, impl_kind = cpu_infer_rate(n_tiles, steps, warmup)
        metric = "infer images/sec (double-branch, 2ch 256x256, batch 256/GPU, + Pearson)"
        workload = INFER_WORKLOAD
        sample = f"{n_tiles} of the {BATCH} tiles of one step (pearson loop + double-branch eval forward), fp32, torch CPU; " + _kind_text(impl_kind)
    else:
        kind = args.model
        n_tiles = 8
        rate, sec, cores, impl_kind = cpu_train_rate(kind, n_tiles, steps, warmup)
        metric = f"train images/sec ({kind}-branch, 2ch 256x256, batch {args.batch}/GPU)"
        workload = TRAIN_WORKLOAD.format(kind=kind, batch=args.batch, cfg=3 if kind == "double" else 2)
        sample = f"{n_tiles}-tile batches of the same step (forward + MSELoss + backward + Adam), fp32, torch CPU; " + _kind_text(impl_kind)
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": "images/sec", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "per_gpu_batch": args.batch, "global_batch": args.batch,
                       "parallelism": "host CPU cores, rank 0 only", "bounded_sample_tiles": n_tiles,
                       "weights": "seed-0 random init"},
            "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": impl_kind, "sample": sample},
            "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers for the GPU arm
def _barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(vals, dev, world):
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def _timeline_table(timeline, denom, by_role=False):
    """Per entry point (by_role: per (entry point, meta['role'])): summed CUDA-event time, launches, algorithmic FLOP."""
    per = {}
    for name, a, b, meta in timeline:
        key = f"{name}:{(meta or {}).get('role', '')}" if by_role else name
        d = per.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0})
        d["ms"] += a.elapsed_time(b)
        d["n"] += 1
        d["flops"] += (meta or {}).get("flops", 0.0)
    for d in per.values():
        d["ms_per_step"] = d["ms"] / denom
    return per


def build_model(kind, dev, randomized_bn=False):
    import ctk
    from ctk import synthetic
    torch.manual_seed(0)
    model = (ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64) if kind == "double"
             else ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6))
    if randomized_bn:
        model.load_state_dict(synthetic.randomize_bn(model.state_dict(), seed=7))
    return model.to(dev)


# ------------------------------------------------------------------------------------------------ training measurement
def measure_train(args, kind, steps, warmup, with_scheduler=False):
    """One training line: K steps timed with CUDA events between barriers (max over ranks), the tensor-core entry points
    carrying their own events inside the timed region; then the same loop end to end from pinned host batches."""
    import ctk
    from ctk import _lib, synthetic
    world, rank, local = env()
    dev = torch.device("cuda", local)
    batch = args.batch
    model = build_model(kind, dev).train()
    eng = ctk.models.get_train_engine(model)
    eng.overlap_streams = bool(getattr(args, "overlap_streams", False))
    if getattr(args, "overlap_wgrad", False):
        eng.overlap_wgrad = eng.overlap_pack = True
    concurrent_backward = eng.overlap_wgrad or eng.overlap_streams
    eng_overlap_wgrad = eng.overlap_wgrad and not eng.overlap_streams
    sync = None
    if world > 1:
        ctk.parallel.broadcast_parameters(model)
        sync = ctk.parallel.attach(model, sync_bn=args.sync_bn)
    opt = ctk.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    # the schedule the reference configures for -r cosine_warmup (train_model.py:356-365), stepped once per EPOCH there:
    # within the timed steps of epoch 0 it holds lr = max_lr / warmup_epochs
    sched = ctk.CosineWarmupLR(opt, warmup_epochs=5, max_lr=5e-4, final_lr=1e-7, total_epochs=50) if with_scheduler else None
    crit = ctk.MSELoss()
    base_x, base_y = synthetic.synthetic_batch(32, seed=1234 + rank)
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

    for i in range(warmup):
        step(*devb[i % 2]).item()
    _barrier(world)
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
    _barrier(world)
    _lib.stop_timeline()
    launches = _lib.launch_count - launches0
    detail = _lib.start_timeline()
    for i in range(2):
        step(*devb[i % 2]).item()
    _barrier(world)
    _lib.stop_timeline()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    # end to end: the batch starts in pinned host memory every step (the DataLoader's pin_memory=True path, :607-614)
    e2e_steps = max(4, steps)
    pre = ctk.DevicePrefetcher(device=str(dev))
    for xb, yb in pre.iterate(host[i % 2] for i in range(2)):                 # staging buffers allocated, copies warm
        step(xb, yb).item()
    _barrier(world)
    t0 = time.perf_counter()
    for xb, yb in pre.iterate(host[i % 2] for i in range(e2e_steps)):
        step(xb, yb).item()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    _barrier(world)
    ms, e2e_ms = _max_over_ranks([ms, e2e_ms], dev, world)
    dp = dp_checks(model, kind, args) if world > 1 else None
    line = None
    if rank == 0:
        pk = peaks()
        per = _timeline_table(timeline, steps)
        per_all = _timeline_table(detail, 2)
        zero = {"ms": 0.0, "n": 0, "flops": 0.0}
        roles = _timeline_table(timeline, steps, by_role=True)
        conv_fwd, conv_dg = roles.get("ctk_conv3x3_tc_raw:fwd", zero), roles.get("ctk_conv3x3_tc_raw:dgrad", zero)
        # With the weight gradients on their own stream the dgrad launches share the GPU with them: CUDA events around such a
        # launch also count the time its CTAs wait for SMs the previous wgrad still holds.  The kernel's roofline is then taken
        # from the launches that own the GPU while they run -- the train-mode forward ones (same kernel template, the heavier
        # raw+stats epilogue); the dgrad launches are reported beside it as measured.
        conv = conv_fwd if concurrent_backward else per.get("ctk_conv3x3_tc_raw", zero)
        wg = per.get("ctk_conv3x3_wgrad_tc", zero)
        conv_tf = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        dg_tf = conv_dg["flops"] / (conv_dg["ms"] / 1e3) / 1e12 if conv_dg["ms"] > 0 else 0.0
        wg_tf = wg["flops"] / (wg["ms"] / 1e3) / 1e12 if wg["ms"] > 0 else 0.0
        traffic, traffic_src = ncu_dram_traffic("conv3x3_tc_kernel", ("r2b_train_full_raw.csv", "r2_train_full_raw.csv"))
        value = world * batch * steps / (ms / 1e3)
        cfg_idx = 3 if kind == "double" else 2
        line = {
            "metric": f"train images/sec ({kind}-branch, 2ch 256x256, batch {batch}/GPU)", "value": value, "unit": "images/sec",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD.format(kind=kind, batch=batch, cfg=cfg_idx),
                       "per_gpu_batch": batch, "global_batch": batch * world,
                       "parallelism": (f"dp{world}: bucketed NCCL all-reduce (AVG) overlapped with backward"
                                       + (", SyncBN" if args.sync_bn else ", per-rank BatchNorm statistics")) if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (134 MB per batch), 2 distinct batches rotated",
                       "lr_schedule": f"ctk.CosineWarmupLR, epoch 0 (lr {opt.param_groups[0]['lr']:.1e})" if sched is not None else "constant 5e-4 (what the reference's cosine_warmup amounts to, SURVEY D6)",
                       "weights": "seed-0 random init", "dropout": "libctk Philox keep-masks", "loss": "ctk.MSELoss"},
            "whole_net_tflops": value * GFLOP_TRAIN[kind] / 1e3,
            "roofline": {"kernel": ("conv3x3_tc_kernel (raw+stats forward launches; the dgrad launches run beside the weight "
                                    "gradients and are listed under dgrad_launches)") if concurrent_backward
                                   else "conv3x3_tc_kernel (raw+stats forward and dgrad launches)", "bound": "tensor",
                         "achieved": conv_tf, "peak": pk["bf16_tflops_burst"], "unit": "TFLOP/s",
                         "frac": conv_tf / pk["bf16_tflops_burst"], "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (ncu --set full, mean over the launches of a step)",
                         "traffic_source": traffic_src,
                         "peak_source": pk["src"] + " (burst cuBLAS bf16: the timed region is tens of milliseconds)",
                         "frac_of_sustained_cublas": conv_tf / pk["bf16_tflops"],
                         "frac_of_nominal_dense": conv_tf / NOMINAL_BF16_TFLOPS,
                         "algorithmic_flop_per_launch": conv["flops"] / max(1, conv["n"]),
                         "avg_launch_ms": conv["ms"] / max(1, conv["n"]), "launches": conv["n"],
                         "share_of_step": conv["ms"] / ms if ms > 0 else None,
                         "dgrad_launches": {"achieved": dg_tf, "avg_launch_ms": conv_dg["ms"] / max(1, conv_dg["n"]),
                                            "launches": conv_dg["n"],
                                            "concurrent_with": "wgrad_tc_kernel on the weight-gradient stream" if concurrent_backward else None},
                         "backward_schedule": ("weight gradients on a high-priority side stream behind each layer's dgrad, "
                                               "beside the next layer's BatchNorm-backward passes (TrainEngine.overlap_wgrad)")
                                              if eng_overlap_wgrad else "single stream",
                         "wgrad_tc_kernel": {"achieved": wg_tf, "frac": wg_tf / pk["bf16_tflops_burst"],
                                             "avg_launch_ms": wg["ms"] / max(1, wg["n"]), "launches": wg["n"],
                                             "share_of_step": wg["ms"] / ms if ms > 0 else None},
                         "per_call_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_all.items())},
                         "per_call_tflops": {k: round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)
                                             for k, v in sorted(per_all.items()) if v["ms"] > 0 and v["flops"] > 0},
                         "per_call_source": "two fully instrumented steps after the timed region"},
            "clocks": clocks, "gpu_launches": launches, "last_loss": last,
            "e2e": {"value": world * batch * e2e_steps / (e2e_ms / 1e3), "unit": "images/sec",
                    "h2d_bytes_per_step": batch * (2 * 256 * 256 + 1) * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "for x, y in ctk.DevicePrefetcher().iterate(pinned host batches): model(x); ctk.MSELoss; backward; "
                           "ctk.Adam.step; loss.item()  (every step's H2D and loss read-back inside the timed region)"}}
        if sync is not None:
            n_steps = warmup + steps + 2 + 2 + e2e_steps
            line["allreduce"] = {"collectives_per_step": sync.collectives / n_steps, "bytes_per_step": sync.bytes_reduced / n_steps}
        if dp is not None:
            line["dp"] = dp
    del model, opt, devb, eng
    torch.cuda.empty_cache()
    return line


def dp_checks(model, kind, args):
    """N > 1, after the timed steps.  (1) every rank must hold bit-identical parameters and buffers-free state: all-reduce
    MIN and MAX of per-tensor checksums.  (2) SyncBN: world x (16 / world) tiles with global BatchNorm statistics against the
    single-process step on all 16 tiles (what the reference's one-process loop computes, train_model.py:408,420)."""
    import torch.distributed as dist
    import ctk
    from ctk import synthetic
    world, rank, local = env()
    dev = torch.device("cuda", local)
    sums = torch.stack([p.detach().double().sum() for p in model.parameters()] +
                       [p.detach().double().abs().sum() for p in model.parameters()])
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ranks_agree = bool(torch.equal(lo, hi))
    out = {"ranks_agree": ranks_agree, "checksum_tensors": int(sums.numel())}
    if args.sync_bn:
        return out            # BN buffers also agree only under SyncBN; the per-rank-statistics run checks parameters alone
    total = 16
    if total % world:
        return out
    per = total // world
    x, y = synthetic.synthetic_batch(total, seed=77)
    g = torch.Generator().manual_seed(5)
    p_drop = 0.1 if kind == "single" else 0.5
    m1 = (torch.rand(total, 512, generator=g) >= p_drop).float()
    m2 = (torch.rand(total, 128, generator=g) >= p_drop).float()
    x, y, m1, m2 = x.to(dev), y.to(dev), m1.to(dev), m2.to(dev)

    def one_step(m, xs, ys, masks):
        eng = ctk.models.get_train_engine(m)
        eng.forced_masks = masks
        for p in m.parameters():
            p.grad = None
        loss = ctk.MSELoss()(m(xs), ys)
        loss.backward()
        return loss.detach(), [p.grad.detach().clone() for p in m.parameters()]

    full = build_model(kind, dev).train()
    loss_full, g_full = one_step(full, x, y, (m1, m2))
    dp = build_model(kind, dev).train()
    ctk.parallel.broadcast_parameters(dp)
    ctk.parallel.attach(dp, sync_bn=True)
    sl = slice(rank * per, (rank + 1) * per)
    loss_loc, g_dp = one_step(dp, x[sl].contiguous(), y[sl].contiguous(), (m1[sl].contiguous(), m2[sl].contiguous()))
    loss_dp = loss_loc.clone()
    dist.all_reduce(loss_dp, op=dist.ReduceOp.AVG)
    num = sum(((a.double() - b.double()) ** 2).sum() for a, b in zip(g_dp, g_full))
    den = sum((b.double() ** 2).sum() for b in g_full)
    sd_f, sd_d = full.state_dict(), dp.state_dict()
    stat_err = max(((sd_d[k].double() - sd_f[k].double()).abs().max() / (sd_f[k].double().abs().max() + 1e-12)).item()
                   for k in sd_f if "running_" in k)
    out["syncbn_check"] = {"tiles": total, "per_rank": per,
                           "loss_rel_err": abs(loss_dp.item() - loss_full.item()) / abs(loss_full.item()),
                           "whole_gradient_rel_l2": float((num / den).sqrt().item()),
                           "running_stat_rel_err": stat_err,
                           "what": f"{world} x {per} tiles with sync_bn=True vs one process on {total} tiles, bf16 operands"}
    del full, dp
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ inference measurement
def measure_infer(args, steps, warmup):
    import ctk
    from ctk import _lib, synthetic
    world, rank, local = env()
    dev = torch.device("cuda", local)
    model = build_model("double", dev, randomized_bn=True).eval()
    if args.precision != "bf16":
        ctk.set_precision(model, args.precision)
    base, _ = synthetic.synthetic_batch(32, seed=1234 + rank)
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

    with torch.no_grad():
        for i in range(warmup):
            step(i)
        _barrier(world)
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
        _barrier(world)
        _lib.stop_timeline()
        launches = _lib.launch_count - launches0
        detail = _lib.start_timeline()
        for i in range(3):
            step(i)
        _barrier(world)
        _lib.stop_timeline()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        # ---- end to end: host buffers in, host results out, through the public HostScorer API
        scorer = ctk.HostScorer(model, slice_tiles=64, device=str(dev))
        for _ in scorer.score_stream(host_batches[i % n_rot] for i in range(3)):
            pass
        _barrier(world)
        e2e_steps = max(5, steps)
        t0 = time.perf_counter()
        n_out = 0
        for s_host, r_host in scorer.score_stream(host_batches[i % n_rot] for i in range(e2e_steps)):
            n_out += s_host.numel()             # host results of every step are consumed inside the timed region
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert n_out == e2e_steps * BATCH
        _barrier(world)
    ms, e2e_ms = _max_over_ranks([ms, e2e_ms], dev, world)
    line = None
    if rank == 0:
        pk = peaks()
        per = _timeline_table(timeline, steps)
        per_all = _timeline_table(detail, 3)
        conv = per.get("ctk_conv3x3_tc_eval") or per.get("ctk_conv3x3_tc_eval_split") or {"ms": 0.0, "n": 1, "flops": 0.0}
        conv_tf = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        traffic, traffic_src = ncu_dram_traffic("conv3x3_tc_kernel", ("r2b_infer_full_raw.csv", "r2_infer_full_raw.csv", "r1d_infer_full_raw.csv"))
        value = world * BATCH * steps / (ms / 1e3)
        line = {"metric": "infer images/sec (double-branch, 2ch 256x256, batch 256/GPU, + Pearson)", "value": value,
                "unit": "images/sec", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class: hi/lo operand pairs, fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": INFER_WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world,
                           "parallelism": f"tiles sharded x{world}, no collective",
                           "l2_policy": "inputs larger than L2 (134 MB per batch), 3 distinct batches rotated",
                           "weights": "seed-0 random init, randomised BN stats"},
                "whole_net_tflops": value * GFLOP_INFER["double"] / 1e3,
                "roofline": {"kernel": "conv3x3_tc_kernel", "bound": "tensor", "achieved": conv_tf,
                             "peak": pk["bf16_tflops_burst"], "unit": "TFLOP/s", "frac": conv_tf / pk["bf16_tflops_burst"],
                             "traffic": traffic,
                             "traffic_unit": "DRAM bytes per launch (ncu --set full, mean of the 6 launches of a step)",
                             "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": (537e6 + 268e6 + 268e6 + 134e6 + 134e6 + 67e6) / 3,
                             "algorithmic_flop_per_launch": CONV_GFLOP_PER_LAUNCH * 1e9,
                             "peak_source": pk["src"] + " (burst cuBLAS bf16: the timed region is tens of milliseconds)",
                             "frac_of_sustained_cublas": conv_tf / pk["bf16_tflops"],
                             "frac_of_nominal_dense": conv_tf / NOMINAL_BF16_TFLOPS,
                             "avg_launch_ms": conv["ms"] / max(1, conv["n"]), "launches": conv["n"],
                             "share_of_step": conv["ms"] / ms if ms > 0 else None,
                             "per_call_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_all.items())},
                             "per_call_source": "three fully instrumented steps after the timed region"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": world * BATCH * e2e_steps / (e2e_ms / 1e3), "unit": "images/sec",
                        "h2d_bytes_per_step": scorer.h2d_bytes, "d2h_bytes_per_step": scorer.d2h_bytes, "steps": e2e_steps,
                        "ms_per_step": e2e_ms / e2e_steps,
                        "api": "ctk.HostScorer.score_stream(pinned host batches) -> host scores + Pearson r per batch "
                               "(one batch of look-ahead; every step's H2D and D2H inside the timed region)"}}
        # the host's measured pinned-H2D ceiling with `world` ranks copying at once (tools/probe_h2d_ranks.py), if committed
        ceil_path = os.path.join(ROOT, "profiles", f"r2_h2d_ranks_{world}.json")
        if os.path.exists(ceil_path):
            ceiling = json.load(open(ceil_path)).get("tiles_per_s_ceiling_together")
            if ceiling:
                line["e2e"]["host_h2d_ceiling_tiles_per_s"] = ceiling
                line["e2e"]["fraction_of_host_h2d_ceiling"] = line["e2e"]["value"] / ceiling
                line["e2e"]["host_h2d_ceiling_source"] = os.path.join("profiles", f"r2_h2d_ranks_{world}.json")
    del model, engine, dev_batches, scorer
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------ bulk sweep (configs[4])
def measure_sweep(args):
    """BASELINE.json configs[4]: `--tiles` synthetic tiles sharded over the ranks (parallel.shard_range, no collective),
    generated on the device in 256-tile batches, scored by the double-branch model + Pearson in the chosen precision.
    64 tiles of rank 0's shard are re-scored by the CPU oracle in the same run.  Replaces the batch-1 loop of
    test-cross-talk-model.py:44-51,302-308."""
    import ctk
    from ctk import _lib, synthetic
    world, rank, local = env()
    dev = torch.device("cuda", local)
    model = build_model("double", dev, randomized_bn=True).eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ctk.set_precision(model, args.precision)
    engine = ctk.models.get_engine(model)
    begin, end = ctk.parallel.shard_range(args.tiles, rank, world)
    n_local = end - begin
    # a ring of device-generated 256-tile batches (distinct seeds per rank and slot, each 134 MB > L2); generation itself is
    # ATen RNG and excluded from the timed region ("generated on-device so H2D is excluded")
    ring = [synthetic.synthetic_batch(BATCH, seed=10_000 + rank * 16 + i, device=dev)[0] for i in range(4)]
    scores = torch.empty(n_local, 1, device=dev)
    pear = torch.empty(n_local, device=dev, dtype=torch.float64)
    with torch.no_grad():
        for i in range(3):
            ctk.pearson_per_image(ring[i % 4], out=pear[:BATCH])
            engine.forward(ring[i % 4], out=scores[:BATCH])
        _barrier(world)
        launches0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0.record()
        done, i = 0, 0
        while done < n_local:
            nb = min(BATCH, n_local - done)
            x = ring[i % 4][:nb]
            ctk.pearson_per_image(x, out=pear[done:done + nb])
            engine.forward(x, out=scores[done:done + nb])
            done += nb
            i += 1
        e1.record()
        _barrier(world)
        clocks = sampler.stop() if rank == 0 else None
        launches = _lib.launch_count - launches0
        ms = e0.elapsed_time(e1)
    (ms,) = _max_over_ranks([ms], dev, world)
    line = None
    if rank == 0:
        # parity spot check on the exact sweep path: 64 tiles of the first batches against the CPU oracle
        orc = _oracle()
        idx = torch.arange(0, min(n_local, 4 * BATCH), max(1, min(n_local, 4 * BATCH) // 64))[:64]
        xs = torch.stack([ring[(int(j) // BATCH) % 4][int(j) % BATCH] for j in idx]).cpu()
        with torch.no_grad():
            ref = orc.double_forward(sd, xs).flatten()
        got = scores[idx.to(dev)].flatten().cpu()
        r_ref = torch.from_numpy(orc.pearson_batch(xs))
        r_got = pear[idx.to(dev)].cpu()
        value = args.tiles / (ms / 1e3)
        line = {"metric": f"bulk-sweep images/sec (double-branch inference + Pearson, {args.tiles} tiles)", "value": value,
                "unit": "images/sec", "n_gpus": world, "steps": (n_local + BATCH - 1) // BATCH, "warmup": 3,
                "ms_per_step": ms / max(1, (n_local + BATCH - 1) // BATCH), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class)", "data": "synthetic",
                "config": {"workload": f"bulk inference sweep: {args.tiles} synthetic 2ch 256x256 tiles sharded over {world} GPU(s), "
                                       f"{args.precision} (BASELINE.json configs[4])",
                           "tiles": args.tiles, "tiles_per_rank": n_local, "per_gpu_batch": BATCH,
                           "parallelism": f"parallel.shard_range x{world}, no collective",
                           "l2_policy": "ring of 4 distinct device-generated batches (134 MB each, larger than L2)"},
                "whole_net_tflops": value * GFLOP_INFER["double"] / 1e3, "total_ms": ms, "clocks": clocks,
                "gpu_launches": launches,
                "parity_spot_check": {"tiles": int(idx.numel()), "score_max_abs_err": float((got - ref).abs().max()),
                                      "pearson_max_abs_err": float((r_got - r_ref).abs().max()),
                                      "bound": 1e-3 if args.precision == "bf16" else 1e-5}}
    del model, engine, ring
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ctk", choices=["ctk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-attach", action="store_true", help="headline mode: skip the attached infer / train_single lines")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="inference arithmetic: bf16 operands (default) or the fp32-class split-bf16 path")
    ap.add_argument("--sync-bn", action="store_true", help="training, N > 1: BatchNorm statistics over the global batch")
    ap.add_argument("--mode", default="headline", choices=["headline", "train", "infer", "sweep"],
                    help="headline = double-branch training (configs[3]) + attached infer / train_single; train / infer = one "
                         "line; sweep = configs[4]")
    ap.add_argument("--model", default="double", choices=["double", "single"])
    ap.add_argument("--overlap-streams", action="store_true",
                    help="training, EXPERIMENTAL: branches and weight gradients on side streams (measured: no gain)")
    ap.add_argument("--overlap-wgrad", action="store_true",
                    help="training, experiment: weight packing and deferred weight gradients on side streams (measured: no gain)")
    ap.add_argument("--batch", type=int, default=BATCH, help="per-GPU batch (training)")
    ap.add_argument("--tiles", type=int, default=1_000_000, help="sweep mode: total tiles over all ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world, rank, local = env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ctk hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(3, args.warmup)
    line = None
    if args.mode == "sweep":
        line = measure_sweep(args)
    elif args.mode == "infer":
        line = measure_infer(args, steps, warmup)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, sec, cores, impl_kind = cpu_infer_rate(16, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": impl_kind,
                                    "sample": "16 of the 256 tiles of one step (pearson loop + double-branch eval forward), fp32; "
                                              + _kind_text(impl_kind)}
    else:
        kind = args.model if args.mode == "train" else "double"
        line = measure_train(args, kind, steps, warmup)
        if args.mode == "headline" and not args.no_attach:
            infer = measure_infer(args, steps, warmup)
            single = None
            if world == 1:
                sargs = argparse.Namespace(**vars(args))
                single = measure_train(sargs, "single", min(steps, 10), 3, with_scheduler=True)
            if rank == 0:
                keep = ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "dtype", "config", "whole_net_tflops",
                        "roofline", "clocks", "gpu_launches", "e2e")
                line["infer"] = {k: infer[k] for k in keep}
                if single is not None:
                    line["train_single"] = {k: single[k] for k in keep}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, sec, cores, impl_kind = cpu_train_rate(kind, 8, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": impl_kind,
                                    "sample": "8-tile batches of the same step (forward + MSELoss + backward + Adam), fp32; "
                                              + _kind_text(impl_kind)}
            if "infer" in line:
                rate, sec, cores, impl_kind = cpu_infer_rate(16, 2, 1)
                line["infer"]["cpu_baseline"] = {"value": rate, "unit": "images/sec", "cores": cores, "kind": impl_kind,
                                                 "sample": "16 of the 256 tiles of one step (pearson loop + eval forward), fp32"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
