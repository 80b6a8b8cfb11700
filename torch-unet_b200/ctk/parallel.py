"""Data-parallel gradient exchange for the ctk training path.

One process per GPU; the global batch is split evenly and the only data-path exchange is the gradient all-reduce.
``GradSynchronizer`` hooks ``TrainEngine.on_grad_ready``: gradients arrive in backward order (head first, then the
conv blocks last-to-first), large tensors (FC1's 537 MB weight gradient is the first big one out) are reduced in
place as soon as their wgrad kernel has been enqueued, small ones are packed into ~25 MB buckets.  Every collective is
launched asynchronously -- NCCL runs it on its own stream behind the producing kernel -- so the exchange overlaps the
remaining dgrad / wgrad work, and the compute stream only waits for it at the end of ``backward``.
The result is the MEAN over ranks (matching ``MSELoss(reduction='mean')`` over the global batch); with NCCL that is a
single ``ReduceOp.AVG``.  torch.distributed is plumbing here: communicator setup, stream ordering, the collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.distributed as dist

from ._lib import CtkError

BIG_TENSOR_BYTES = 8 << 20
BUCKET_BYTES = 4 << 20            # small gradients: 12.7 MB per step in all, flushed as they accumulate, not at the end
# SMs the backward pass's persistent tensor-core kernels leave free while the exchange runs (TrainEngine.backward_sm_reserve
# -> ctk_set_persistent_sm_reserve).  Measured on 8 x B200 (tools/r2_dp_experiment.sh, gpurun_out/r2_dp_8.txt): reserve
# 0 / 16 / 32 SMs -> 16.36 / 16.41 / 16.85 ms per step, and 15.90 / 16.25 / 16.39 ms at 2 GPUs: NCCL's kernel finds its
# SMs at kernel boundaries without help, and the reserve only slows the convolutions.  Hence off by default.
BACKWARD_SM_RESERVE = int(os.environ.get("CTK_DP_SM_RESERVE", "0"))


class GradSynchronizer:
    def __init__(self, process_group=None, bucket_bytes: int = BUCKET_BYTES, big_tensor_bytes: int = BIG_TENSOR_BYTES):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_bytes = bucket_bytes
        self.big_bytes = big_tensor_bytes
        self.use_avg = dist.get_backend(process_group) == "nccl"
        self._pending: List[Tuple[object, torch.Tensor]] = []      # (work, tensor to divide when not AVG)
        self._bucket: List[Tuple[object, torch.Tensor]] = []       # (key, grad)
        self._bucket_bytes = 0
        self._replaced: Dict[object, torch.Tensor] = {}
        self.collectives = 0
        self.bytes_reduced = 0

    # ------------------------------------------------------------------ called by TrainEngine.backward
    def on_grad_ready(self, key, grad: torch.Tensor) -> None:
        nbytes = grad.numel() * grad.element_size()
        if nbytes >= self.big_bytes and grad.is_contiguous():
            self._launch(grad)
            return
        self._bucket.append((key, grad))
        self._bucket_bytes += nbytes
        if self._bucket_bytes >= self.bucket_bytes:
            self._flush()

    def finalize(self, grads: Dict[object, torch.Tensor]) -> Dict[object, torch.Tensor]:
        """Flush the last bucket, make the current stream wait for every collective, hand back the reduced gradients."""
        self._flush()
        for work, t in self._pending:
            work.wait()
            if not self.use_avg:
                t.div_(self.world)
        self._pending.clear()
        out = dict(grads)
        out.update(self._replaced)
        self._replaced = {}
        return out

    # ------------------------------------------------------------------ internals
    def _launch(self, t: torch.Tensor) -> None:
        op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM
        work = dist.all_reduce(t, op=op, group=self.pg, async_op=True)
        self._pending.append((work, t))
        self.collectives += 1
        self.bytes_reduced += t.numel() * t.element_size()

    def _flush(self) -> None:
        if not self._bucket:
            return
        flat = torch.cat([g.reshape(-1) for _, g in self._bucket])
        self._launch(flat)
        off = 0
        for key, g in self._bucket:
            n = g.numel()
            self._replaced[key] = flat[off:off + n].view(g.shape)     # gradients become views of the reduced bucket
            off += n
        self._bucket = []
        self._bucket_bytes = 0


def attach(model: torch.nn.Module, process_group=None, sync_bn: bool = False, **kw) -> GradSynchronizer:
    """Make ``loss.backward()`` of a ctk model all-reduce (average) its gradients across the process group.

    ``sync_bn=True`` additionally sums every BatchNorm's batch statistics (and the matching backward reductions) over the
    ranks, so that normalisation uses the GLOBAL batch like the single-process reference does (train_model.py runs one
    process; its BatchNorm sees all of ``-b``).  These are a few KB per layer: ~20 small blocking all-reduces per step.
    Without it the statistics are per rank, which is what torch's DistributedDataParallel does by default."""
    from .models import get_train_engine
    sync = GradSynchronizer(process_group, **kw)
    eng = get_train_engine(model)
    eng.on_grad_ready = sync.on_grad_ready
    eng.finalize_grads = sync.finalize
    if sync.world > 1 and sync.use_avg:
        eng.backward_sm_reserve = max(0, min(64, BACKWARD_SM_RESERVE))
    if sync_bn and sync.world > 1:
        pg = process_group
        eng.stat_allreduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
        eng.stat_world = sync.world

        def check_equal_batch(n: int, dev) -> None:
            # global statistics are formed with count = n * world: ragged shards (shard_range on a batch that does not
            # divide, a ragged last DataLoader batch) would be normalised with the wrong count -- refuse them
            t = torch.tensor([float(n), -float(n)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
            hi, neg_lo = t.tolist()
            if hi != -neg_lo:
                raise CtkError(f"sync_bn=True needs the same number of tiles on every rank (this rank {n}, "
                               f"min {int(-neg_lo)}, max {int(hi)}): pad or drop the ragged batch")

        eng.stat_check_equal_batch = check_equal_batch
    return sync


def broadcast_parameters(model: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """All ranks start from rank ``src``'s parameters and buffers (what DDP does at construction)."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)


def shard_range(n: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """[begin, end) of this rank's contiguous share of ``n`` units (tiles of an inference sweep, images of a global
    training batch).  Ranks differ by at most one unit; the first ``n % world`` ranks get the extra one."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if not (0 <= rank < world) or n < 0:
        raise ValueError("need 0 <= rank < world and n >= 0")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's rows of a tensor laid out over the GLOBAL batch (inputs, labels)."""
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e]


def shard_dropout_masks(masks, rank: Optional[int] = None, world: Optional[int] = None):
    """Parity mode: every rank draws (or is handed) the Dropout keep-masks of the GLOBAL batch and uses its
    own rows, so that N ranks x B/N images reproduce the single-process step on B images mask for mask.
    ``masks``: the (m1 [B,512], m2 [B,128]) pair ``TrainEngine.forced_masks`` takes."""
    return tuple(shard(m, rank, world) for m in masks)
