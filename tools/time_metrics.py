"""Device time of the per-tile comparison metrics on one 256-tile batch (CUDA events, 20 repetitions after 3 warm-ups,
three distinct 134 MB batches rotated so the input does not sit in L2).  python tools/time_metrics.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))
import ctk  # noqa: E402

g = torch.Generator().manual_seed(1234)
batches = [torch.rand(256, 2, 256, 256, generator=g).cuda() for _ in range(3)]
out = {}
for name, fn in (("pearson", ctk.pearson_per_image), ("tile_metrics (pearson+rmse+hist+hist_corr)", ctk.tile_metrics),
                 ("nmi", ctk.nmi_per_image), ("ssim", ctk.ssim_per_image)):
    for i in range(3):
        fn(batches[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        fn(batches[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[name] = {"ms_per_256_tiles": ms, "tiles_per_sec": 256 / ms * 1e3, "input_GB_per_s": 256 * 524288 / ms / 1e6}
print(json.dumps(out, indent=1))
