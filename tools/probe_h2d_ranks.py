"""Aggregate pinned host->device bandwidth of this box when every rank copies at once (what bounds the end-to-end inference
number at 8 GPUs).  Run under torchrun with N ranks; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 tools/probe_h2d_ranks.py

Per rank: a 134 MB pinned batch (one bench input batch) copied to its GPU in four 33.5 MB pieces, 20 times, (a) alone --
the other ranks idle -- and (b) all ranks together between barriers.  Also prints the NUMA node of each GPU and the CPU
affinity of the rank, which is what decides whose memory controller a pinned buffer lands on."""
import json
import os

import torch
import torch.distributed as dist


def copy_rate(h, d, reps=20):
    hs, ds = h.chunk(4), d.chunk(4)
    for _ in range(3):
        for a, b in zip(hs, ds):
            b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for a, b in zip(hs, ds):
            b.copy_(a, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return h.numel() * 4 * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 256 * 2 * 256 * 256
    h = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    alone = torch.zeros(world, device="cuda", dtype=torch.float64)
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone[r] = copy_rate(h, d)
        torch.cuda.synchronize()
    dist.all_reduce(alone)
    dist.barrier()
    torch.cuda.synchronize()
    together = torch.zeros(world, device="cuda", dtype=torch.float64)
    together[rank] = copy_rate(h, d)
    dist.all_reduce(together)
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        numa = open(f"/sys/bus/pci/devices/0000:{bus:02x}:00.0/numa_node").read().strip()
    except Exception:
        numa = "?"
    info = [None] * world
    dist.all_gather_object(info, {"rank": rank, "gpu_numa_node": numa, "cpu_affinity": len(os.sched_getaffinity(0))})
    if rank == 0:
        a, t = alone.tolist(), together.tolist()
        print(json.dumps({"ranks": world, "h2d_gbs_alone": [round(v, 1) for v in a], "h2d_gbs_together": [round(v, 1) for v in t],
                          "aggregate_alone_sum": round(sum(a), 1), "aggregate_together": round(sum(t), 1),
                          "tiles_per_s_ceiling_together": round(sum(t) * 1e9 / (2 * 256 * 256 * 4)),
                          "cpus": os.cpu_count(), "ranks_info": info}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
