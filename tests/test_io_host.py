"""CPU tests of the host side of the input pipeline: the minimal TIFF payload locator and the oracle restatement."""
import glob
import os
import struct

import numpy as np
import pytest

import crosstalk_oracle as orc


def _tiff_bytes(img: np.ndarray, big_endian=False, pad=0) -> bytes:
    """A classic single-strip uncompressed TIFF like the ones tifffile wrote for the reference's Training_Data."""
    e = ">" if big_endian else "<"
    h, w = img.shape
    fmt, bits = {"f": (3, img.itemsize * 8), "u": (1, img.itemsize * 8)}[img.dtype.kind]
    entries = [(256, 4, w), (257, 4, h), (258, 3, bits), (259, 3, 1), (262, 3, 1), (277, 3, 1), (278, 4, h), (339, 3, fmt)]
    n = len(entries) + 2
    data_off = 8 + 2 + 12 * n + 4 + pad
    entries += [(273, 4, data_off), (279, 4, img.nbytes)]
    entries.sort()
    out = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, 8) + struct.pack(e + "H", n)
    for tag, typ, val in entries:
        out += struct.pack(e + "HHI", tag, typ, 1) + (struct.pack(e + "H", val) + b"\0\0" if typ == 3 else struct.pack(e + "I", val))
    out += struct.pack(e + "I", 0) + b"\0" * pad
    return out + img.astype(img.dtype.newbyteorder(e)).tobytes()


@pytest.mark.parametrize("dtype,big", [("f8", False), ("f4", False), ("u2", False), ("f8", True)])
def test_tiff_payload_locator(tmp_path, dtype, big):
    from ctk import io
    rng = np.random.default_rng(0)
    img = (rng.random((24, 40)) * 1000).astype(dtype)
    p = tmp_path / "t.tif"
    p.write_bytes(_tiff_bytes(img, big_endian=big, pad=6))
    got = io.read_tiff_plane(str(p))
    assert got.shape == img.shape and np.array_equal(got.astype(dtype), img)


def test_tiff_locator_on_reference_fixtures(golden):
    from ctk import io
    files = sorted(glob.glob("/root/reference/Training_Data/Mixed/*.tif"))
    if not files:
        pytest.skip("reference fixtures are only present in the build container")
    f = [x for x in files if "image_14144_" in x][0]
    off, h, w, dt = io.tiff_payload_info(open(f, "rb").read())
    assert (off, h, w, dt) == (272, 256, 256, np.dtype("<f8"))
    # train_model.py:166: imread(...).astype(np.float32) == the tile make_golden.py stored
    assert np.array_equal(io.read_tiff_plane(f).astype(np.float32), golden["tiles"][0, 0])


def test_prepare_tiles_restatement(golden):
    tiles = golden["tiles"]
    raw = tiles.astype(np.float64)
    flips = np.array([0, 1, 2, 3, 0], dtype=np.uint8)
    out = orc.prepare_tiles(raw, flips)
    for i in range(5):
        for c in range(2):
            ref = orc.normalize_image(tiles[i, c])
            if flips[i] & 1:
                ref = np.flip(ref, axis=-1)
            if flips[i] & 2:
                ref = np.flip(ref, axis=-2)
            assert np.array_equal(out[i, c], ref)
            assert out[i, c].min() == 0.0 and out[i, c].max() == 1.0


def test_product_synthetic_generator_matches_the_oracle_copy():
    """bench.py's GPU arm draws its tiles from ctk.synthetic (product code, no oracle import); the oracle keeps its own copy
    of the SURVEY 8d generator for the tests.  The two must stay bit-identical, and so must randomize_bn."""
    import torch
    import crosstalk_oracle as orc
    from ctk import synthetic
    xa, ya = synthetic.synthetic_batch(5, seed=1234)
    xb, yb = orc.synthetic_batch(5, seed=1234)
    assert torch.equal(xa, xb) and torch.equal(ya, yb)
    assert xa.shape == (5, 2, 256, 256) and float(xa.amin()) == 0.0 and float(xa.amax()) == 1.0
    sd = orc.init_double_state_dict(0)
    ra, rb = synthetic.randomize_bn(sd, seed=7), orc.randomize_bn(sd, seed=7)
    assert ra.keys() == rb.keys() and all(torch.equal(ra[k], rb[k]) for k in ra)
