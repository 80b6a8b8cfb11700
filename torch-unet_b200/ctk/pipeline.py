"""Host-buffer entry points: score tiles that live in (pinned) host memory.

This is the call a user of test-cross-talk-model.py's loop (:44-64) makes once the path is swapped in: hand
over host batches, get back the predicted crosstalk score and the Pearson r per tile.

* ``HostScorer.score(batch)`` -- one batch, synchronous.  The batch is cut into slices; slice i+1 is copied
  host->device on a copy stream while slice i is computed.
* ``HostScorer.score_stream(batches)`` -- a generator over many batches (what a DataLoader loop is).  Batch k+1 is copied
  into the second staging buffer while batch k is computed as ONE launch sequence, and batch k's results are read
  back while batch k+1 computes, so in steady state a step costs max(PCIe time, kernel time).

Host->device copies are issued in ``copy_tiles`` pieces (default 64 tiles = 33.5 MB): on the B200 boxes measured, a
single 134 MB cudaMemcpyAsync from pinned memory runs at 30-38 GB/s while the same bytes as four back-to-back copies
reach 54.8 GB/s (tools/probe_h2d.py, gpurun_out/probe_h2d.log).
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch

from . import _lib
from .metrics import nmi_per_image, pearson_per_image, ssim_per_image, tile_metrics
from .models import get_engine


class HostScorer:
    def __init__(self, model: torch.nn.Module, slice_tiles: int = 64, device: str = "cuda", metrics: str = "pearson"):
        """``metrics``: "pearson" -> results are (scores, r);  "all" -> (scores, {"pearson", "rmse", "ssim", "hist_corr",
        "nmi"}), every comparison metric of test-cross-talk-model.py:59-85 for every tile."""
        if metrics not in ("pearson", "all"):
            raise _lib.CtkError("metrics must be 'pearson' or 'all'")
        self.metrics = metrics
        self.model = model
        self.slice = slice_tiles
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._slots = [None, None]           # double-buffered staging: device input, device/host results, reuse event
        self._turn = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ------------------------------------------------------------------ staging
    def _slot(self, tiles_host):
        i = self._turn
        self._turn ^= 1
        n = tiles_host.shape[0]
        key = (n, tuple(tiles_host.shape[1:]), tiles_host.dtype)
        sl = self._slots[i]
        if sl is None or sl["key"] != key:
            if sl is not None and sl["free"] is not None:
                sl["free"].synchronize()
            sl = {"key": key,
                  "x": torch.empty((n, *tiles_host.shape[1:]), device=self.dev, dtype=tiles_host.dtype),
                  "scores": torch.empty(n, 1, device=self.dev, dtype=torch.float32),
                  "r": torch.empty(n, device=self.dev, dtype=torch.float64),
                  "extra": torch.empty(4, n, device=self.dev, dtype=torch.float64),      # rmse, hist_corr, nmi, ssim
                  "scores_h": torch.empty(n, dtype=torch.float32, pin_memory=True),
                  "r_h": torch.empty(n, dtype=torch.float64, pin_memory=True),
                  "extra_h": torch.empty(4, n, dtype=torch.float64, pin_memory=True),
                  "free": None, "done": None}
            self._slots[i] = sl
        return sl

    def _check(self, tiles_host):
        if tiles_host.is_cuda:
            raise _lib.CtkError("HostScorer takes host tensors; call the model directly for device tensors")
        if self.model.training:
            raise _lib.CtkError("HostScorer scores with the eval-mode path: call model.eval() first")

    def _submit(self, tiles_host, compute_tiles):
        """Enqueue copies + kernels + result read-back of one batch; returns the slot (wait on slot['done'])."""
        self._check(tiles_host)
        engine = get_engine(self.model)
        sl = self._slot(tiles_host)
        n = tiles_host.shape[0]
        main = torch.cuda.current_stream(self.dev)
        if sl["free"] is not None:
            self.copy_stream.wait_event(sl["free"])     # the kernels that last read this staging buffer are done
        events = []
        for s in range(0, n, self.slice):
            e = min(n, s + self.slice)
            with torch.cuda.stream(self.copy_stream):
                sl["x"][s:e].copy_(tiles_host[s:e], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            events.append((e, ev))
        done_to = 0
        for e, ev in events:
            if e - done_to < compute_tiles and e < n:
                continue
            main.wait_event(ev)
            part = sl["x"][done_to:e]
            if self.metrics == "all":
                m = tile_metrics(part)
                sl["r"][done_to:e] = m["pearson"]
                sl["extra"][0, done_to:e] = m["rmse"].double()
                sl["extra"][1, done_to:e] = m["hist_corr"]
                sl["extra"][2, done_to:e] = nmi_per_image(part)
                sl["extra"][3, done_to:e] = ssim_per_image(part)
            else:
                pearson_per_image(part, out=sl["r"][done_to:e])
            engine.forward(part, out=sl["scores"][done_to:e])
            done_to = e
        sl["free"] = torch.cuda.Event()
        sl["free"].record(main)
        sl["scores_h"].copy_(sl["scores"].flatten(), non_blocking=True)
        sl["r_h"].copy_(sl["r"], non_blocking=True)
        if self.metrics == "all":
            sl["extra_h"].copy_(sl["extra"], non_blocking=True)
        sl["all"] = self.metrics == "all"
        sl["done"] = torch.cuda.Event()
        sl["done"].record(main)
        self.h2d_bytes = tiles_host.numel() * tiles_host.element_size()
        self.d2h_bytes = n * 4 + n * 8 + (4 * n * 8 if self.metrics == "all" else 0)
        return sl

    @staticmethod
    def _finish(sl):
        sl["done"].synchronize()
        if sl["all"]:
            ex = sl["extra_h"].clone()
            return sl["scores_h"].clone(), {"pearson": sl["r_h"].clone(), "rmse": ex[0].float(), "hist_corr": ex[1], "nmi": ex[2], "ssim": ex[3]}
        return sl["scores_h"].clone(), sl["r_h"].clone()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def score(self, tiles_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """tiles_host: [N,2,H,W] float32 on the host (pinned for full speed).  Returns (scores[N] f32, r[N] f64) on the host."""
        return self._finish(self._submit(tiles_host, self.slice))

    @torch.no_grad()
    def score_stream(self, batches: Iterable[torch.Tensor]) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yield (scores, r) for every host batch of ``batches``, in order, keeping one batch of look-ahead in flight."""
        pending = None
        for tiles_host in batches:
            job = self._submit(tiles_host, tiles_host.shape[0])
            if pending is not None:
                yield self._finish(pending)
            pending = job
        if pending is not None:
            yield self._finish(pending)


class DevicePrefetcher:
    """Host batches -> device with the copy of batch k+1 running on a side stream while the caller's kernels work on
    batch k.  The training-loop counterpart of ``HostScorer.score_stream``; staging buffers persist across ``iterate``
    calls (epochs).  Batches are a tensor or a tuple of tensors, ideally pinned (the reference's DataLoader uses
    ``pin_memory=True``, train_model.py:607-614):

        pre = ctk.DevicePrefetcher()
        for inputs, labels in pre.iterate(train_loader):        # train_model.py:415-417 without the .to(device)
            ...
    """

    def __init__(self, device: str = "cuda", depth: int = 2):
        self.dev = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = [None] * depth        # device tensors, ready event, released event

    def _issue(self, batch, i):
        single = torch.is_tensor(batch)
        items = (batch,) if single else tuple(batch)
        slot = self.slots[i]
        if slot is None or len(slot["t"]) != len(items) or any(a.shape != b.shape or a.dtype != b.dtype
                                                               for a, b in zip(slot["t"], items)):
            if slot is not None and slot["released"] is not None:
                slot["released"].synchronize()
            slot = {"t": [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in items], "released": None}
            self.slots[i] = slot
        if slot["released"] is not None:
            self.copy_stream.wait_event(slot["released"])      # the kernels that read this slot's previous batch are enqueued
        with torch.cuda.stream(self.copy_stream):
            for d, h in zip(slot["t"], items):
                n = h.shape[0] if h.dim() > 0 else 0
                if h.dim() > 0 and h.numel() * h.element_size() > (64 << 20) and n >= 4:
                    # large copies in four pieces: a single 134 MB cudaMemcpyAsync runs at 30-38 GB/s on these boxes,
                    # the same bytes in four back-to-back copies at 54.8 GB/s (tools/probe_h2d.py)
                    step = (n + 3) // 4
                    for a in range(0, n, step):
                        d[a:a + step].copy_(h[a:a + step], non_blocking=True)
                else:
                    d.copy_(h, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self.copy_stream)
        slot["single"] = single

    def _hand_out(self, i):
        slot = self.slots[i]
        torch.cuda.current_stream(self.dev).wait_event(slot["ready"])
        return slot, (slot["t"][0] if slot["single"] else tuple(slot["t"]))

    def iterate(self, batches: Iterable):
        pending = []                       # slot indices in flight, oldest first
        turn = 0
        last = None

        def release(slot):
            # the generator was resumed: the consumer has enqueued all its work on the batch handed out last, so an
            # event recorded now on its stream marks the point after which that slot may be overwritten
            if slot is not None:
                slot["released"] = torch.cuda.Event()
                slot["released"].record(torch.cuda.current_stream(self.dev))

        for batch in batches:
            release(last)
            last = None
            self._issue(batch, turn)
            pending.append(turn)
            turn = (turn + 1) % self.depth
            if len(pending) == self.depth:
                last, out = self._hand_out(pending.pop(0))
                yield out
        while pending:
            release(last)
            last, out = self._hand_out(pending.pop(0))
            yield out
        release(last)


def prefetch_to_device(batches: Iterable, device: str = "cuda", depth: int = 2):
    """One-off form of ``DevicePrefetcher(device, depth).iterate(batches)``."""
    return DevicePrefetcher(device, depth).iterate(batches)
