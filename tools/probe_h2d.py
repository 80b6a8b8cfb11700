"""Pinned host->device copy behaviour on this box (what bounds bench.py's e2e number)."""
import os
import time
import torch


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("cpus", os.cpu_count(), "numa nodes", [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
for mb in (8.4, 33.5, 134.2, 268.4, 536.9):
    n = int(mb * 1e6) // 4
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    print(f"H2D pinned {mb:7.1f} MB x1 stream: {ms:.3f} ms  {n * 4 / ms / 1e6:.1f} GB/s")
    ms = timed(lambda: h.copy_(d, non_blocking=True))
    print(f"D2H pinned {mb:7.1f} MB x1 stream: {ms:.3f} ms  {n * 4 / ms / 1e6:.1f} GB/s")

n = int(134.2e6) // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for parts in (2, 4, 8):
    hs, ds = h.chunk(parts), d.chunk(parts)
    ms = timed(lambda: [b.copy_(a, non_blocking=True) for a, b in zip(hs, ds)])
    print(f"134 MB as {parts} back-to-back copies, one stream: {ms:.3f} ms {n * 4 / ms / 1e6:.1f} GB/s")
    streams = [torch.cuda.Stream() for _ in range(parts)]

    def multi():
        cur = torch.cuda.current_stream()
        for a, b, s in zip(hs, ds, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
    ms = timed(multi)
    print(f"134 MB as {parts} copies on {parts} streams: {ms:.3f} ms {n * 4 / ms / 1e6:.1f} GB/s")

# copy while the SMs are busy
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
side = torch.cuda.Stream()


def overlapped():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        d.copy_(h, non_blocking=True)
    for _ in range(4):
        a @ a
    cur.wait_stream(side)
ms_mm = timed(lambda: [a @ a for _ in range(4)])
ms = timed(overlapped)
print(f"4 matmuls alone {ms_mm:.3f} ms; with a concurrent 134 MB H2D {ms:.3f} ms")
