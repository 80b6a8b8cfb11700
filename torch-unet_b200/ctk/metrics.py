"""Per-image comparison metrics on the device (replaces the host loop at test-cross-talk-model.py:52-64)."""
from __future__ import annotations

from ctypes import c_int, c_size_t

import torch

from . import _lib
from ._lib import call, ptr, stream


def pearson_per_image(inputs: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Pearson r between channel 0 and channel 1 of every [2,H,W] float32 tile of ``inputs`` ([N,2,H,W], CUDA).

    Returns float64 [N]; NaN where either plane is constant (the ``np.std(...) == 0`` guard of
    test-cross-talk-model.py:61-62); clipped to [-1, 1] like scipy.stats.pearsonr.
    """
    _lib.require_device(inputs, torch.float32, "inputs")
    if inputs.dim() != 4 or inputs.shape[1] != 2:
        raise _lib.CtkError(f"inputs must be [N,2,H,W], got {tuple(inputs.shape)}")
    n = inputs.shape[0]
    plane = inputs.shape[2] * inputs.shape[3]
    if out is None:
        out = torch.empty(n, device=inputs.device, dtype=torch.float64)
    else:
        _lib.require_device(out, torch.float64, "out")
        if out.numel() != n:
            raise _lib.CtkError("out must hold one float64 per tile")
    if n == 0:
        return out
    lib = _lib.load()
    ws_bytes = lib.ctk_pearson_workspace_bytes(c_int(n))
    ws = torch.empty(ws_bytes // 8, device=inputs.device, dtype=torch.float64)
    call("ctk_pearson_f32", ptr(inputs), c_int(n), c_int(plane), ptr(out), ptr(ws), c_size_t(ws_bytes), stream())
    return out


def nmi_per_image(inputs: torch.Tensor) -> torch.Tensor:
    """Normalised mutual information of channel 0 vs channel 1 of every tile after np.digitize into 256 levels
    (test-cross-talk-model.py:71-74,84).  ``inputs``: [N,2,H,W] float32 CUDA; returns float64 [N]."""
    _lib.require_device(inputs, torch.float32, "inputs")
    if inputs.dim() != 4 or inputs.shape[1] != 2:
        raise _lib.CtkError(f"inputs must be [N,2,H,W], got {tuple(inputs.shape)}")
    n = inputs.shape[0]
    out = torch.empty(n, device=inputs.device, dtype=torch.float64)
    if n == 0:
        return out
    lib = _lib.load()
    ws_bytes = lib.ctk_tile_nmi_workspace_bytes(c_int(n))
    ws = torch.empty((ws_bytes + 7) // 8, device=inputs.device, dtype=torch.float64)
    call("ctk_tile_nmi_f32", ptr(inputs), c_int(n), c_int(inputs.shape[2] * inputs.shape[3]), ptr(out), ptr(ws),
         c_size_t(ws_bytes), stream())
    return out


def ssim_per_image(inputs: torch.Tensor) -> torch.Tensor:
    """Mean structural similarity of channel 0 vs channel 1 of every tile, as the reference's evaluation loop calls
    scikit-image (test-cross-talk-model.py:80-82: default 7x7 uniform window, ``data_range`` = max - min over both
    planes).  ``inputs``: [N,2,H,W] float32 CUDA, H, W >= 7; returns float64 [N]."""
    _lib.require_device(inputs, torch.float32, "inputs")
    if inputs.dim() != 4 or inputs.shape[1] != 2:
        raise _lib.CtkError(f"inputs must be [N,2,H,W], got {tuple(inputs.shape)}")
    if not inputs.is_contiguous():
        raise _lib.CtkError("inputs must be contiguous")
    n, _, h, w = inputs.shape
    out = torch.empty(n, device=inputs.device, dtype=torch.float64)
    if n == 0:
        return out
    lib = _lib.load()
    chunk = 65535                                   # one grid row per tile: the entry point takes at most 65535 tiles
    ws_bytes = lib.ctk_tile_ssim_workspace_bytes(c_int(min(n, chunk)))
    ws = torch.empty((ws_bytes + 7) // 8, device=inputs.device, dtype=torch.float64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        call("ctk_tile_ssim_f32", ptr(inputs[s:e]), c_int(e - s), c_int(h), c_int(w), ptr(out[s:e]), ptr(ws),
             c_size_t(ws_bytes), stream())
    return out


def tile_metrics(inputs: torch.Tensor, histograms: bool = True) -> dict:
    """Pearson r, RMSE and 256-bin histogram correlation of channel 0 vs channel 1 for every tile of ``inputs``
    ([N,2,H,W] float32, CUDA) in one fused pass -- test-cross-talk-model.py:59-70,79 without the device->host copy of the
    inputs (:52).  Returns {"pearson": f64[N], "rmse": f32[N], "hist_corr": f64[N], "hist": i32[N,2,256]} on the device
    (the last two only with ``histograms=True``); histogram counts equal ``np.histogram(plane, bins=256)[0]`` exactly.
    """
    _lib.require_device(inputs, torch.float32, "inputs")
    if inputs.dim() != 4 or inputs.shape[1] != 2:
        raise _lib.CtkError(f"inputs must be [N,2,H,W], got {tuple(inputs.shape)}")
    n = inputs.shape[0]
    plane = inputs.shape[2] * inputs.shape[3]
    dev = inputs.device
    out = {"pearson": torch.empty(n, device=dev, dtype=torch.float64), "rmse": torch.empty(n, device=dev, dtype=torch.float32)}
    if histograms:
        out["hist_corr"] = torch.empty(n, device=dev, dtype=torch.float64)
        out["hist"] = torch.empty(n, 2, 256, device=dev, dtype=torch.int32)
    if n == 0:
        return out
    lib = _lib.load()
    ws_bytes = lib.ctk_tile_metrics_workspace_bytes(c_int(n))
    ws = torch.empty((ws_bytes + 7) // 8, device=dev, dtype=torch.float64)
    call("ctk_tile_metrics_f32", ptr(inputs), c_int(n), c_int(plane), ptr(out["pearson"]), ptr(out["rmse"]),
         ptr(out.get("hist_corr")), ptr(out.get("hist")), ptr(ws), c_size_t(ws_bytes), stream())
    return out
