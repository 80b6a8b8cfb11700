"""Synthetic workloads for benchmarks and sweeps: tiles shaped and normalised like the reference's inputs.

The reference trains on 2-channel 256x256 float32 tiles, each plane min-max normalised (train_model.py:211-216), with the
label alpha in [0.01, 0.5] (the crosstalk fraction in the file names, train_model.py:105).  ``synthetic_batch`` draws a
uniform "source" plane and mixes it into the first channel with that alpha, so Pearson r and the regression target are
non-trivial.  This module is product code (bench.py's GPU arm and the bulk sweep use it); the test oracle keeps its own
copy and ``tests/test_io_host.py`` checks the two agree bit for bit.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def _normalise(t: torch.Tensor) -> torch.Tensor:
    lo = t.amin(dim=(1, 2), keepdim=True)
    hi = t.amax(dim=(1, 2), keepdim=True)
    return (t - lo) / (hi - lo)


def synthetic_batch(n: int, seed: int = 1234, size: int = 256, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(tiles [n,2,size,size] float32, labels [n,1] float32).  ``device=None``: host tensors from the CPU generator (the
    seeded set tests and benchmarks share); a CUDA device: drawn and normalised on that device (bulk sweeps)."""
    dev = torch.device("cpu") if device is None else torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    src = torch.rand(n, size, size, generator=g, device=dev)
    noise = torch.rand(n, size, size, generator=g, device=dev)
    alpha = 0.01 + 0.49 * torch.rand(n, generator=g, device=dev)
    mixed = alpha[:, None, None] * src + (1.0 - alpha[:, None, None]) * noise
    x = torch.stack([_normalise(mixed), _normalise(src)], dim=1).contiguous()
    return x, alpha[:, None].contiguous()


def randomize_bn(sd: Dict[str, torch.Tensor], seed: int = 7) -> Dict[str, torch.Tensor]:
    """A state_dict whose BatchNorm layers have non-trivial gamma (a quarter negative), beta and running statistics:
    random-init eval outputs are nearly constant otherwise, which would make throughput runs see a degenerate head."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out.keys()):
        if k.endswith("running_mean"):
            base = k[: -len("running_mean")]
            n = out[k].numel()
            gamma = 0.5 + torch.rand(n, generator=g)
            sign = torch.where(torch.rand(n, generator=g) < 0.25, -1.0, 1.0)
            out[base + "weight"] = gamma * sign
            out[base + "bias"] = 0.2 * torch.randn(n, generator=g)
            out[base + "running_mean"] = out[base + "running_mean"] + 0.05 * torch.randn(n, generator=g)
            out[base + "running_var"] = out[base + "running_var"] * (0.5 + torch.rand(n, generator=g))
    return out
