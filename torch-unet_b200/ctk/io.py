"""Input side of the hot path: TIFF pixel payloads -> device -> normalised [N,2,H,W] float32 batches.

The reference reads each plane with ``imageio.imread(path).astype(np.float32)`` and min-max normalises it on the host
(train_model.py:166-167, 211-216).  Its fixtures are uncompressed single-image TIFFs (float64, one strip at a fixed
offset), so the pixel payload can be sliced out of the file bytes without decoding anything and shipped to the GPU as
is; ``prepare_tiles`` then does the cast, the normalisation and the augmentation flips in one kernel.
"""
from __future__ import annotations

import struct
from ctypes import c_int
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 16: "Q"}


def tiff_payload_info(buf: bytes) -> Tuple[int, int, int, np.dtype]:
    """(offset, height, width, dtype) of the pixel payload of an uncompressed, contiguous, single-sample classic TIFF."""
    if buf[:2] == b"II":
        e = "<"
    elif buf[:2] == b"MM":
        e = ">"
    else:
        raise _lib.CtkError("not a TIFF file")
    if struct.unpack(e + "H", buf[2:4])[0] != 42:
        raise _lib.CtkError("only classic (non-Big) TIFF is supported")
    ifd = struct.unpack(e + "I", buf[4:8])[0]
    n = struct.unpack(e + "H", buf[ifd:ifd + 2])[0]
    tags = {}
    for i in range(n):
        ent = buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(e + "HHI", ent[:8])
        size = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 16: 8}.get(typ)
        if size is None:
            continue
        raw = ent[8:12] if size * cnt <= 4 else buf[struct.unpack(e + "I", ent[8:12])[0]:][:size * cnt]
        if typ in (3, 4, 16):
            tags[tag] = list(struct.unpack(e + _TYPES[typ] * cnt, raw[:size * cnt]))
    w, h = tags[256][0], tags[257][0]
    bits = tags.get(258, [1])[0]
    fmt = tags.get(339, [1])[0]                       # 1 unsigned, 2 signed, 3 IEEE float
    if tags.get(259, [1])[0] != 1 or tags.get(277, [1])[0] != 1:
        raise _lib.CtkError("only uncompressed single-sample TIFFs are on the fast path")
    offs, counts = tags[273], tags.get(279)
    if counts is not None and any(o + c != o2 for o, c, o2 in zip(offs, counts, offs[1:])):
        raise _lib.CtkError("TIFF strips are not contiguous")
    kind = {(3, 64): "f8", (3, 32): "f4", (1, 16): "u2", (1, 8): "u1"}.get((fmt, bits))
    if kind is None:
        raise _lib.CtkError(f"unsupported TIFF sample format {fmt}/{bits} bits")
    return offs[0], h, w, np.dtype(e + kind)


def read_tiff_plane(path: str) -> np.ndarray:
    """The image of a fixture-style TIFF as a NumPy array in its stored dtype (no imageio needed)."""
    buf = open(path, "rb").read()
    off, h, w, dt = tiff_payload_info(buf)
    return np.frombuffer(buf, dtype=dt, count=h * w, offset=off).reshape(h, w)


def prepare_tiles(raw: torch.Tensor, flips: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``raw``: [N,2,H,W] float64 or float32 on the device (pixel payloads as stored).  Returns the model input batch:
    float32, every plane min-max normalised (constant planes unchanged), sample i flipped horizontally if
    ``flips[i] & 1`` and vertically if ``flips[i] & 2`` (both planes alike) -- train_model.py:166-167, 211-232."""
    if raw.dtype not in (torch.float64, torch.float32):
        raise _lib.CtkError("raw tiles must be float64 or float32")
    _lib.require_device(raw, raw.dtype, "raw tiles")
    if raw.dim() != 4 or raw.shape[1] != 2:
        raise _lib.CtkError(f"raw tiles must be [N,2,H,W], got {tuple(raw.shape)}")
    n, _, h, w = raw.shape
    if flips is not None:
        _lib.require_device(flips, torch.uint8, "flips")
        if flips.numel() != n:
            raise _lib.CtkError("flips must hold one byte per sample")
    if out is None:
        out = torch.empty((n, 2, h, w), device=raw.device, dtype=torch.float32)
    else:
        _lib.require_device(out, torch.float32, "out")
        if tuple(out.shape) != (n, 2, h, w):
            raise _lib.CtkError("out has the wrong shape")
    if n:
        call("ctk_prepare_tiles", ptr(raw), c_int(1 if raw.dtype == torch.float64 else 0), ptr(flips), c_int(n), c_int(h),
             c_int(w), ptr(out), stream())
    return out
