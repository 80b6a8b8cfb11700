"""world_size-2 CPU (gloo) tests of the data-parallel host logic: bucketing, big-tensor path, averaging, ordering."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctk.parallel import GradSynchronizer, broadcast_parameters
        torch.manual_seed(100 + rank)
        shapes = [(3,), (128, 64), (1,), (512, 700), (17, 5, 3, 3), (64,)]
        grads = {i: torch.randn(s) for i, s in enumerate(shapes)}
        keep = {k: v.clone() for k, v in grads.items()}
        sync = GradSynchronizer(bucket_bytes=40_000, big_tensor_bytes=1_000_000)
        for k in sorted(grads):                                   # backward order = insertion order
            sync.on_grad_ready(k, grads[k])
        out = sync.finalize(grads)
        # expected mean over ranks, recomputed independently with blocking collectives
        ok = True
        for k in sorted(keep):
            ref = keep[k].clone()
            dist.all_reduce(ref)
            ref /= world
            ok &= out[k].shape == keep[k].shape and torch.allclose(out[k], ref, atol=1e-6)
        # broadcast_parameters: every rank ends with rank 0's values
        lin = torch.nn.Linear(4, 3)
        broadcast_parameters(lin, src=0)
        w = lin.weight.detach().clone()
        dist.broadcast(w, src=0)
        ok &= torch.equal(w, lin.weight.detach())
        # shard helpers: contiguous, disjoint, covering; mask rows follow the same split as the batch rows
        from ctk.parallel import shard, shard_dropout_masks, shard_range
        g = torch.Generator().manual_seed(7)
        xb = torch.rand(9, 2, 4, 4, generator=g)                 # 9 images over 2 ranks: 5 + 4
        masks = (torch.rand(9, 512, generator=g) > 0.5, torch.rand(9, 128, generator=g) > 0.5)
        b, e = shard_range(9)
        mine = shard(xb)
        m1, m2 = shard_dropout_masks(masks)
        ok &= (b, e) == ((0, 5) if rank == 0 else (5, 9)) and torch.equal(mine, xb[b:e])
        ok &= torch.equal(m1, masks[0][b:e]) and torch.equal(m2, masks[1][b:e])
        sizes = torch.tensor([e - b])
        dist.all_reduce(sizes)
        ok &= int(sizes) == 9
        q.put((rank, bool(ok), sync.collectives, sync.bytes_reduced))
    finally:
        dist.destroy_process_group()


def test_grad_synchronizer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ncoll, nbytes in res:
        assert ok, f"rank {rank} mismatch"
        # one big tensor (512x700 fp32 = 1.4 MB) reduced in place + buckets for the rest
        assert ncoll >= 2
        assert nbytes == sum(n * 4 for n in (3, 128 * 64, 1, 512 * 700, 17 * 5 * 9, 64))


def test_shard_range_covers_every_unit_once():
    sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))
    from ctk.parallel import shard_range
    for n in (0, 1, 7, 8, 255, 256, 2048, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
