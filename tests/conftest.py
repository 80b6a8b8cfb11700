"""pytest config: registers the ``gpu`` marker and puts the product package and the oracle on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-unet_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
    z = np.load(os.path.join(ROOT, "tests", "golden", "tiles.npz"))
    g["tiles"] = z["tiles"]
    g["labels_np"] = z["labels"]
    return g


@pytest.fixture(scope="session")
def golden32():
    """32 of the reference's Training_Data pairs + what the unmodified reference computes for them
    (tests/golden/make_fixtures32.py).  ``x`` is the normalised input batch the reference's dataset would yield."""
    import json
    import numpy as np
    import torch
    import crosstalk_oracle as orc
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden32.json")))
    tiles = np.load(os.path.join(ROOT, "tests", "golden", "tiles32.npz"))["tiles"]
    g["tiles"] = tiles
    g["x"] = torch.from_numpy(np.stack([np.stack([orc.normalize_image(t[0]), orc.normalize_image(t[1])]) for t in tiles]))
    return g
