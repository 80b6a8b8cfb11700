"""Host-buffer entry point: score tiles that live in (pinned) host memory.

This is the call a user of test-cross-talk-model.py's loop (:44-64) makes once the path is swapped in: hand
over a host batch, get back the predicted crosstalk score and the Pearson r per tile.  The batch is cut into
slices; slice i+1 is copied host->device on a copy stream while slice i is computed, so PCIe time hides
behind the kernels.  Results come back with one small device->host copy per call.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from .metrics import pearson_per_image
from .models import get_engine


class HostScorer:
    def __init__(self, model: torch.nn.Module, slice_tiles: int = 64, device: str = "cuda"):
        self.model = model
        self.slice = slice_tiles
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._stage = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _staging(self, n, shape, dtype):
        key = (n, tuple(shape), dtype)
        if self._stage is None or self._stage[0] != key:
            self._stage = (key, torch.empty((n, *shape), device=self.dev, dtype=dtype))
        return self._stage[1]

    @torch.no_grad()
    def score(self, tiles_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """tiles_host: [N,2,H,W] float32 on the host (pinned for full speed).  Returns (scores[N] f32, r[N] f64) on the host."""
        if tiles_host.is_cuda:
            raise _lib.CtkError("HostScorer.score takes host tensors; call the model directly for device tensors")
        if self.model.training:
            raise _lib.CtkError("HostScorer scores with the eval-mode path: call model.eval() first")
        engine = get_engine(self.model)
        n = tiles_host.shape[0]
        dev_in = self._staging(n, tiles_host.shape[1:], tiles_host.dtype)
        main = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(main)          # staging buffer reuse: previous call's kernels must be done
        events = []
        for s in range(0, n, self.slice):
            e = min(n, s + self.slice)
            with torch.cuda.stream(self.copy_stream):
                dev_in[s:e].copy_(tiles_host[s:e], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            events.append((s, e, ev))
        scores = torch.empty(n, 1, device=self.dev, dtype=torch.float32)
        r = torch.empty(n, device=self.dev, dtype=torch.float64)
        for s, e, ev in events:
            main.wait_event(ev)
            pearson_per_image(dev_in[s:e], out=r[s:e])
            engine.forward(dev_in[s:e], out=scores[s:e])
        out_scores = scores.flatten().to("cpu", non_blocking=False)
        out_r = r.to("cpu", non_blocking=False)
        self.h2d_bytes = tiles_host.numel() * tiles_host.element_size()
        self.d2h_bytes = out_scores.numel() * 4 + out_r.numel() * 8
        return out_scores, out_r
