"""Secondary comparison asked for by SURVEY 8(d): the same two module graphs executed by PyTorch's *library* kernels
(ATen / cuDNN / cuBLAS) on the B200 -- eager fp32, TF32, bf16 autocast, bf16 autocast + channels_last -- next to the
libctk path, on bench.py's workload (256 synthetic 2x256x256 tiles per step, inputs resident on the device).

This is a measurement tool, not a product path and not a fallback: nothing in ctk imports it.  The eager forward below
strings the mirrored modules' own nn.Sequential children together exactly like the reference's forward()
(regression_model.py:58-61, two_branch_regression.py:85-100).

    python tools/library_path_timing.py [--batch 256] [--steps 5] [--out gpurun_out/library_path.json]
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))


def eager_forward(model, x):
    if hasattr(model, "conv_layers"):                                   # regression_model.py:58-61
        return model.fc_layers(model.conv_layers(x))
    fb = model.bleed_branch.conv_blocks(x[:, 0:1])                      # two_branch_regression.py:88-93
    fs = model.source_branch.conv_blocks(x[:, 1:2])
    return model.regression_head.fc_layers(torch.cat((fb, fs), dim=1)) * 0.5


def build(kind):
    import ctk
    torch.manual_seed(0)
    if kind == "single":
        return ctk.AdvancedRegressionModel(initial_filters=128, num_conv_blocks=6)          # train_model.py:537
    return ctk.SimplifiedTwoBranchRegressionModel(initial_filters_per_branch=64)            # train_model.py:535


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--models", default="double,single")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "library_path.json"))
    a = ap.parse_args()

    import ctk
    from ctk import _lib
    torch.backends.cudnn.benchmark = True                               # give the library path its best algorithms
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(a.batch, 2, 256, 256, generator=g).cuda()
    y = (torch.rand(a.batch, 1, generator=g) * 0.49 + 0.01).cuda()
    rows = []
    variants = [("fp32", False, None, False), ("tf32", True, None, False),
                ("bf16_autocast", True, torch.bfloat16, False), ("bf16_autocast_channels_last", True, torch.bfloat16, True)]
    for kind in a.models.split(","):
        for name, tf32, amp, cl in variants:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ctx = (lambda: torch.autocast("cuda", dtype=amp)) if amp else contextlib.nullcontext
            model = build(kind).cuda()
            if cl:
                model = model.to(memory_format=torch.channels_last)
            xin = x.contiguous(memory_format=torch.channels_last) if cl else x
            row = {"model": kind, "path": "torch_eager_" + name, "batch": a.batch}
            try:
                model.eval()

                def infer():
                    with torch.no_grad(), ctx():
                        return eager_forward(model, xin)
                ms = timed(infer, a.steps, a.warmup)
                row["infer_ms_per_step"], row["infer_images_per_sec"] = ms, a.batch / ms * 1e3
                model.train()
                opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)   # train_model.py:637

                def train():
                    opt.zero_grad()
                    with ctx():
                        out = eager_forward(model, xin)
                    loss = torch.nn.functional.mse_loss(out.float(), y)
                    loss.backward()
                    opt.step()
                ms = timed(train, a.steps, a.warmup)
                row["train_ms_per_step"], row["train_images_per_sec"] = ms, a.batch / ms * 1e3
                row["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
            except Exception as e:                                      # e.g. out of memory: report, keep going
                row["error"] = repr(e)[:200]
            rows.append(row)
            print(json.dumps(row), flush=True)
            del model
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
        # the libctk path on the same inputs, same timing loop
        _lib.load()
        model = build(kind).cuda().eval()
        row = {"model": kind, "path": "libctk_bf16", "batch": a.batch}

        def infer_ctk():
            with torch.no_grad():
                return model(x)
        ms = timed(infer_ctk, a.steps, a.warmup)
        row["infer_ms_per_step"], row["infer_images_per_sec"] = ms, a.batch / ms * 1e3
        model.train()
        opt = ctk.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)

        def train_ctk():
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(model(x), y)
            loss.backward()
            opt.step()
        ms = timed(train_ctk, a.steps, a.warmup)
        row["train_ms_per_step"], row["train_images_per_sec"] = ms, a.batch / ms * 1e3
        rows.append(row)
        print(json.dumps(row), flush=True)
        del model, opt
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "torch": torch.__version__,
                   "cudnn": torch.backends.cudnn.version(), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
